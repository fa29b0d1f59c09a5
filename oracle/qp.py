"""Oracle (test infrastructure, CPU): QuadraticProgram backend arithmetic.

Restates, array-level only, ``src/QuadraticProgram/QuadraticProgram.jl``:

* ``create_lhs``        <- ``create_LHS_matrix``            (:256-282)
* ``solve_system``      <- ``solve_system``                 (:486-496)
* ``reverse``           <- ``reverse_differentiate!``       (:316-351)
* ``forward``           <- ``forward_differentiate!``       (:357-446, RHS at :429-433)
* ``reverse_param_grads`` <- the lazy getters               (:307-314, :448-473)

Conventions: block order (z, lambda, nu); the cached
matrix is K^T of OptNet eq. (6); reverse solves with LHS, forward with LHS';
``lam``/``nu`` are the NEGATED MOI duals (:156-180); LSQR is used iff
``norm(Q) == 0`` (:333, :436).  Substitution named per ``oracle/__init__.py``:
Julia's sparse ``\\`` (UMFPACK) -> LAPACK dense LU with partial pivoting, or
SuperLU (``splu``) when ``sparse=True``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .lsqr import lsqr


def create_lhs(z, lam, Q, G, h, A):
    """``[Q G'diag(lam) A'; G diag(Gz-h) 0; A 0 0]`` (QuadraticProgram.jl:256-282)."""
    z = np.asarray(z, float)
    n = z.size
    Q = np.asarray(Q, float).reshape(n, n)
    G = np.zeros((0, n)) if G is None else np.asarray(G, float).reshape(-1, n)
    A = np.zeros((0, n)) if A is None else np.asarray(A, float).reshape(-1, n)
    m, p = G.shape[0], A.shape[0]
    lam = np.asarray(lam, float).reshape(m)
    h = np.asarray(h, float).reshape(m)
    N = n + m + p
    K = np.zeros((N, N))
    K[:n, :n] = Q
    if m:
        K[:n, n:n + m] = G.T * lam[None, :]
        K[n:n + m, :n] = G
        K[n:n + m, n:n + m] = np.diag(G @ z - h)
    if p:
        K[:n, n + m:] = A.T
        K[n + m:, :n] = A
    return K


def is_iterative(Q):
    """``norm(Q) ≈ 0`` with default isapprox tolerances == exact zero (:333)."""
    return float(np.linalg.norm(np.asarray(Q, float))) == 0.0


def solve_system(LHS, RHS, iterative, sparse=False, **lsqr_kw):
    """``iterative ? lsqr(LHS, RHS) : LHS \\ RHS`` (QuadraticProgram.jl:486-492)."""
    if iterative:
        return lsqr(LHS, RHS, **lsqr_kw)
    if sparse or sp.issparse(LHS):
        return spla.splu(sp.csc_matrix(LHS)).solve(np.asarray(RHS, float))
    return np.linalg.solve(LHS, RHS)  # raises LinAlgError on exact singularity


def reverse(Q, G, h, A, z, lam, nu, dl_dz, **kw):
    """(dz, dlam, dnu) = -LHS^{-1} [dl_dz; 0; 0]   (QuadraticProgram.jl:316-351)."""
    K = create_lhs(z, lam, Q, G, h, A)
    n, m = len(z), len(lam)
    rhs = np.zeros(K.shape[0])
    rhs[:n] = dl_dz
    x = -solve_system(K, rhs, is_iterative(Q), **kw)
    return x[:n], x[n:n + m], x[n + m:]


def forward_rhs(z, lam, nu, dQ, dq, dG, dh, dA, db):
    """RHS of :429-433: [dQ z + dq + dG'lam + dA'nu ; lam.(dG z) - lam.dh ; dA z - db]."""
    z = np.asarray(z, float)
    n = z.size
    lam = np.asarray(lam, float)
    nu = np.asarray(nu, float)
    m, p = lam.size, nu.size
    dG = np.asarray(dG, float).reshape(m, n)
    dA = np.asarray(dA, float).reshape(p, n)
    r1 = np.asarray(dQ, float).reshape(n, n) @ z + np.asarray(dq, float) + dG.T @ lam + dA.T @ nu
    r2 = lam * (dG @ z) - lam * np.asarray(dh, float).reshape(m)
    r3 = dA @ z - np.asarray(db, float).reshape(p)
    return np.concatenate([r1, r2, r3])


def forward(Q, G, h, A, z, lam, nu, dQ, dq, dG, dh, dA, db, **kw):
    """(dz, dlam, dnu) = -(LHS')^{-1} RHS          (QuadraticProgram.jl:357-446)."""
    K = create_lhs(z, lam, Q, G, h, A)
    n, m = len(z), len(lam)
    rhs = forward_rhs(z, lam, nu, dQ, dq, dG, dh, dA, db)
    x = -solve_system(K.T, rhs, is_iterative(Q), **kw)
    return x[:n], x[n:n + m], x[n + m:]


def reverse_param_grads(z, lam, nu, dz, dlam, dnu):
    """Gradients w.r.t. problem data from a reverse solve.

    Getter arithmetic of :307-314 and :448-473, returned in the sign
    convention the reference's test-suite reads them in (``test/utils.jl``
    :178-233: ``dhb = -constant``, ``dbb = -constant``, coefficients as is):

        dQ = (dz z' + z dz')/2        dq = dz
        dG = diag(lam)(dlam z' + 1 dz')   i.e. row i: lam_i*dlam_i*z + lam_i*dz
        dh = -lam .* dlam
        dA = dnu z' + nu dz'          db = -dnu
    """
    z, dz = np.asarray(z, float), np.asarray(dz, float)
    lam, dlam = np.asarray(lam, float), np.asarray(dlam, float)
    nu, dnu = np.asarray(nu, float), np.asarray(dnu, float)
    dQ = 0.5 * (np.outer(dz, z) + np.outer(z, dz))
    dq = dz.copy()
    dG = np.outer(lam * dlam, z) + np.outer(lam, dz)
    dh = -lam * dlam
    dA = np.outer(dnu, z) + np.outer(nu, dz)
    db = -dnu
    return dQ, dq, dG, dh, dA, db


# ---- batched convenience wrappers (instance-major, used by parity tests/bench) ----

def batch_forward_reverse(Q, G, A, h, z, lam, nu, seed, dQ, dq, dG, dh, dA, db):
    """Loop of ``forward`` + ``reverse`` over a batch; arrays are (B, rows, cols) row-major
    logical matrices (numpy views), returns (fwd[B,N], rev[B,N])."""
    B = z.shape[0]
    n, m, p = z.shape[1], lam.shape[1], nu.shape[1]
    N = n + m + p
    fwd = np.empty((B, N))
    rev = np.empty((B, N))
    for b in range(B):
        K = create_lhs(z[b], lam[b], Q[b], G[b], h[b], A[b])
        rf = forward_rhs(z[b], lam[b], nu[b], dQ[b], dq[b], dG[b], dh[b], dA[b], db[b])
        rb = np.zeros(N)
        rb[:n] = seed[b]
        fwd[b] = -np.linalg.solve(K.T, rf)
        rev[b] = -np.linalg.solve(K, rb)
    return fwd, rev
