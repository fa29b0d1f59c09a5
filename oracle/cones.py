"""Oracle (test infrastructure, CPU): cone projections pi and their derivative Dpi.

Restates ``DiffOpt.π`` / ``DiffOpt.Dπ`` (``src/diff_opt.jl:491-519``), which
loop over the constraints of a ``ProductOfSets`` (``src/product_of_sets.jl``)
and call MathOptSetDistances' ``projection_on_set`` /
``projection_gradient_on_set`` on ``MOI.dual_set(set)``.

MathOptSetDistances (compat "0.2.9", ``Project.toml:23``) is NOT vendored, so
the per-cone arithmetic restates its published algorithm; the conventions are
pinned by reproducing the reference's golden values in
``tests/test_oracle_kat.py`` (``test/conic_program.jl:107-111, 134-210,
581-647, 801-844``):

* ``Zeros``  -> dual ``Reals``:  pi = id, Dpi = I
* ``Nonnegatives``: pi = max(v, 0), Dpi = diag((sign(v)+1)/2)   (0.5 at v == 0)
* ``SecondOrderCone`` v = (t, x): |x| <= t -> id/I ; |x| <= -t -> 0 ;
  else pi = (|x|+t)/2 (1, x/|x|),
  Dpi = 1/(2|x|) [ |x|  x' ; x  (|x|+t) I - (t/|x|^2) x x' ]
* ``PositiveSemidefiniteConeTriangle(d)``: UNSCALED column-wise upper triangle
  (X11, X12, X22, X13, ...).  pi = vec(U max(L,0) U').  Dpi: all eig >= 0 -> I;
  else k = #{eig < 1e-4}, B_ij = 1 (i,j > k), 0 (i,j <= k),
  eig+_i/(eig-_j + eig+_i) (i > k >= j, symmetric), and ROW idx of Dpi is
  vec(U (B o (U' unvec(e_idx) U)) U')  -- i.e. the TRANSPOSE of the Jacobian in
  unscaled coordinates, a non-symmetric matrix.

Cone type codes are shared with ``include/diffopt_b200.h``.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

ZERO, NONNEG, SOC, PSD = 0, 1, 2, 3   # DIFFOPT_CONE_* in include/diffopt_b200.h
PSD_EIG_THRESHOLD = 1e-4


def psd_side(k):
    d = int((np.sqrt(8 * k + 1) - 1) // 2)
    assert d * (d + 1) // 2 == k, "PSD triangle length must be d(d+1)/2"
    return d


def unvec_symm(v, d):
    X = np.zeros((d, d))
    iu = np.triu_indices(d)
    # column-wise upper triangle: (0,0),(0,1),(1,1),(0,2),... == row-wise lower
    idx = np.lexsort((iu[0], iu[1]))
    r, c = iu[0][idx], iu[1][idx]
    X[r, c] = v
    X[c, r] = v
    return X


def vec_symm(X):
    d = X.shape[0]
    iu = np.triu_indices(d)
    idx = np.lexsort((iu[0], iu[1]))
    return X[iu[0][idx], iu[1][idx]].copy()


def _psd_B(lam):
    d = lam.size
    k = int(np.sum(lam < PSD_EIG_THRESHOLD))
    lp = np.maximum(lam, 0.0)
    lm = -np.minimum(lam, 0.0)
    B = np.zeros((d, d))
    B[k:, k:] = 1.0
    if 0 < k < d:
        blk = lp[k:, None] / (lm[None, :k] + lp[k:, None])
        B[k:, :k] = blk
        B[:k, k:] = blk.T
    return B


def project(v, ctype):
    """``MOSD.projection_on_set(DefaultDistance(), v, dual_set(S))``."""
    v = np.asarray(v, float)
    if ctype == ZERO:
        return v.copy()
    if ctype == NONNEG:
        return np.maximum(v, 0.0)
    if ctype == SOC:
        t, x = v[0], v[1:]
        nx = np.linalg.norm(x)
        if nx <= t:
            return v.copy()
        if nx <= -t:
            return np.zeros_like(v)
        out = np.empty_like(v)
        out[0] = 1.0
        out[1:] = x / nx
        return out * ((nx + t) / 2.0)
    if ctype == PSD:
        d = psd_side(v.size)
        lam, U = np.linalg.eigh(unvec_symm(v, d))
        return vec_symm((U * np.maximum(lam, 0.0)) @ U.T)
    raise ValueError(ctype)


def project_gradient(v, ctype):
    """Dense ``MOSD.projection_gradient_on_set(DefaultDistance(), v, dual_set(S))``."""
    v = np.asarray(v, float)
    k = v.size
    if ctype == ZERO:
        return np.eye(k)
    if ctype == NONNEG:
        return np.diag((np.sign(v) + 1.0) / 2.0)
    if ctype == SOC:
        t, x = v[0], v[1:]
        nx = np.linalg.norm(x)
        if nx <= t:
            return np.eye(k)
        if nx <= -t:
            return np.zeros((k, k))
        D = np.empty((k, k))
        D[0, 0] = nx
        D[0, 1:] = x
        D[1:, 0] = x
        D[1:, 1:] = (nx + t) * np.eye(k - 1) - (t / nx**2) * np.outer(x, x)
        return D / (2.0 * nx)
    if ctype == PSD:
        d = psd_side(k)
        lam, U = np.linalg.eigh(unvec_symm(v, d))
        if np.all(lam >= 0):
            return np.eye(k)
        B = _psd_B(lam)
        D = np.empty((k, k))
        e = np.zeros(k)
        for idx in range(k):
            e[:] = 0.0
            e[idx] = 1.0
            Xt = unvec_symm(e, d)
            D[idx, :] = vec_symm(U @ (B * (U.T @ Xt @ U)) @ U.T)
        return D
    raise ValueError(ctype)


def _psd_F(lam, U, X):
    return U @ (_psd_B(lam) * (U.T @ X @ U)) @ U.T


def apply_gradient(v, ctype, y, transpose=False):
    """Operator form of ``project_gradient(v, ctype) @ y`` (or its transpose) that never
    forms the dense block -- the algorithm the CUDA path implements; verified against
    ``project_gradient`` in the tests."""
    v = np.asarray(v, float)
    y = np.asarray(y, float)
    if ctype == ZERO:
        return y.copy()
    if ctype == NONNEG:
        return (np.sign(v) + 1.0) / 2.0 * y
    if ctype == SOC:  # symmetric block
        t, x = v[0], v[1:]
        nx = np.linalg.norm(x)
        if nx <= t:
            return y.copy()
        if nx <= -t:
            return np.zeros_like(y)
        out = np.empty_like(y)
        xy = x @ y[1:]
        out[0] = nx * y[0] + xy
        out[1:] = x * y[0] + (nx + t) * y[1:] - (t / nx**2) * xy * x
        return out / (2.0 * nx)
    if ctype == PSD:
        d = psd_side(v.size)
        lam, U = np.linalg.eigh(unvec_symm(v, d))
        if np.all(lam >= 0):
            return y.copy()
        if transpose:      # true Jacobian in unscaled coordinates: vec(F(unvec(y)))
            return vec_symm(_psd_F(lam, U, unvec_symm(y, d)))
        # reference's matrix (rows = vec(F(unvec(e_idx)))): S' F T' y
        Y = unvec_symm(y, d)
        Y[np.diag_indices(d)] *= 2.0
        R = _psd_F(lam, U, Y)
        R[np.diag_indices(d)] *= 0.5
        return vec_symm(R)
    raise ValueError(ctype)


def cone_offsets(cone_dims):
    off = np.zeros(len(cone_dims) + 1, dtype=np.int64)
    np.cumsum(np.asarray(cone_dims, dtype=np.int64), out=off[1:])
    return off


def pi(v, cone_types, cone_dims):
    """``DiffOpt.π`` (diff_opt.jl:491-499) over a product of cones."""
    off = cone_offsets(cone_dims)
    out = np.empty(off[-1])
    for c, t in enumerate(cone_types):
        out[off[c]:off[c + 1]] = project(v[off[c]:off[c + 1]], t)
    return out


def Dpi_blocks(v, cone_types, cone_dims):
    """``DiffOpt.Dπ`` (diff_opt.jl:509-519): list of dense blocks (BlockDiagonal)."""
    off = cone_offsets(cone_dims)
    return [project_gradient(v[off[c]:off[c + 1]], t) for c, t in enumerate(cone_types)]


def Dpi_dense(v, cone_types, cone_dims):
    return sla.block_diag(*Dpi_blocks(v, cone_types, cone_dims))


def Dpi_apply(v, cone_types, cone_dims, y, transpose=False):
    off = cone_offsets(cone_dims)
    out = np.empty(off[-1])
    for c, t in enumerate(cone_types):
        s = slice(off[c], off[c + 1])
        out[s] = apply_gradient(v[s], t, y[s], transpose)
    return out


class DpiOperator:
    """``Dpi_apply`` with the per-cone work that does not depend on the argument done once (eigendecompositions,
    triangle index maps): the form an iterative solve needs on large PSD cones.  Same arithmetic as
    ``apply_gradient``; ``tests/test_oracle_kat.py`` checks the two against each other."""

    def __init__(self, v, cone_types, cone_dims):
        v = np.asarray(v, float)
        self.off = cone_offsets(cone_dims)
        self.types = list(cone_types)
        self.v = v
        self.psd = {}
        for c, t in enumerate(self.types):
            if t != PSD:
                continue
            vs = v[self.off[c]:self.off[c + 1]]
            d = psd_side(vs.size)
            iu = np.triu_indices(d)
            idx = np.lexsort((iu[0], iu[1]))
            r, cc = iu[0][idx], iu[1][idx]
            X = np.zeros((d, d))
            X[r, cc] = vs
            X[cc, r] = vs
            lam, U = np.linalg.eigh(X)
            self.psd[c] = (d, r, cc, U, None if np.all(lam >= 0) else _psd_B(lam))

    def apply(self, y, transpose=False):
        y = np.asarray(y, float)
        out = np.empty(self.off[-1])
        for c, t in enumerate(self.types):
            s = slice(self.off[c], self.off[c + 1])
            if t != PSD:
                out[s] = apply_gradient(self.v[s], t, y[s], transpose)
                continue
            d, r, cc, U, B = self.psd[c]
            if B is None:
                out[s] = y[s]
                continue
            Y = np.zeros((d, d))
            Y[r, cc] = y[s]
            Y[cc, r] = y[s]
            if not transpose:
                Y[np.diag_indices(d)] *= 2.0
            R = U @ (B * (U.T @ Y @ U)) @ U.T
            if not transpose:
                R[np.diag_indices(d)] *= 0.5
            out[s] = R[r, cc]
        return out
