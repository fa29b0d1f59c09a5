"""Oracle (test infrastructure, CPU): ConicProgram backend arithmetic.

Restates, array-level only, ``src/ConicProgram/ConicProgram.jl``:

* ``gradient_cache``  <- ``_gradient_cache``             (:172-255; M at :243-247)
* ``forward``         <- ``forward_differentiate!``      (:257-334) + getter :403-412
* ``reverse``         <- ``reverse_differentiate!``      (:336-394)
* ``reverse_param_grads`` <- getters                     (:396-401, :414-443)

Inputs are the arrays AFTER the reference's own sign handling: ``A`` here is the
reference's ``A = -coefficients`` (:179-183), ``b`` the constants, ``c`` already
negated for MAX sense (:206-208); geometric form ``A x + s = b, s in K``.
Reference behaviours kept on purpose: w == 1, v = y - s,
forward dA is used as packed (un-negated), forward zero test 1e-400 (== 0.0),
reverse zero test 1e-4, reverse solves with M (not M'), only dx seeds are used.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from . import cones
from .lsqr import lsqr


@dataclass
class ConicCache:
    M: sp.csc_matrix
    vp: np.ndarray
    Dpi: np.ndarray  # dense block-diagonal (small cases) or sparse
    A: sp.csc_matrix
    b: np.ndarray
    c: np.ndarray
    x: np.ndarray
    s: np.ndarray
    y: np.ndarray


def gradient_cache(A, b, c, x, s, y, cone_types, cone_dims):
    A = sp.csc_matrix(A, dtype=np.float64)
    m, n = A.shape
    b = np.asarray(b, float)
    c = np.asarray(c, float)
    x, s, y = (np.asarray(t, float) for t in (x, s, y))
    v = y - s                                                        # :222
    blocks = cones.Dpi_blocks(v, cone_types, cone_dims)              # :225
    Dpi = sp.block_diag([sp.csc_matrix(B) for B in blocks], format="csc")
    cc = sp.csc_matrix(c.reshape(n, 1))
    bb = sp.csc_matrix(b.reshape(m, 1))
    M = sp.bmat([                                                    # :243-247
        [sp.csc_matrix((n, n)), A.T @ Dpi, cc],
        [-A, sp.identity(m, format="csc") - Dpi, bb],
        [-cc.T, -(bb.T @ Dpi), sp.csc_matrix((1, 1))],
    ], format="csc")
    vp = cones.pi(v, cone_types, cone_dims)                          # :249
    return ConicCache(M, vp, Dpi, A, b, c, x, s, y)


def forward(cache, dA, db, dc, **lsqr_kw):
    """Returns (dx, dz) with dx_i = -(du_i - x_i dw)   (:314-326, :403-412)."""
    n = cache.x.size
    m = cache.b.size
    dA = sp.csc_matrix(dA, shape=(m, n)) if not sp.issparse(dA) else dA
    db = np.asarray(db, float)
    dc = np.asarray(dc, float)
    g = np.concatenate([
        dA.T @ cache.vp + dc,
        -(dA @ cache.x) + db,
        [-(dc @ cache.x) - db @ cache.vp],
    ])
    if np.linalg.norm(g) <= 0.0:           # `<= 1e-400` is `<= 0.0` in Float64 (:320)
        dz = np.zeros_like(g)
    else:
        dz = lsqr(cache.M, g, **lsqr_kw)
    du, dw = dz[:n], dz[n + m]
    return -(du - cache.x * dw), dz


def reverse(cache, dx_seed, **lsqr_kw):
    """Returns g (length n+m+1)   (:349-373)."""
    n = cache.x.size
    m = cache.b.size
    dx = np.asarray(dx_seed, float)
    dz = np.concatenate([dx, np.zeros(m), [-(cache.x @ dx)]])
    if np.linalg.norm(dz) <= 1e-4:                                   # :369
        return np.zeros_like(dz)
    return lsqr(cache.M, dz, **lsqr_kw)


def reverse_param_grads(cache, g, dense_dA=True):
    """dc = g[1:n] - g[N] x ; db = g[n+I] - g[N] vp[I] ; dA = g[n+I] x' - vp[I] g[1:n]'
    (:396-401, :414-428, :430-443).  These are the values of the reference's
    ``ReverseObjectiveFunction`` / ``ReverseConstraintFunction`` getters as is."""
    n = cache.x.size
    gN = g[-1]
    dc = g[:n] - gN * cache.x
    db = g[n:-1] - gN * cache.vp
    dA = np.outer(g[n:-1], cache.x) - np.outer(cache.vp, g[:n]) if dense_dA else None
    return dA, db, dc


def M_apply(A, b, c, v, cone_types, cone_dims, t, transpose=False, dpi=None):
    """Matrix-free M t / M' t -- the operator the CUDA LSQR uses; checked against the
    explicit ``gradient_cache().M`` in the tests.  ``dpi``: optional ``cones.DpiOperator(v, ...)``
    (cached eigendecompositions) for repeated applications."""
    A = sp.csr_matrix(A)
    m, n = A.shape
    t1, t2, t3 = t[:n], t[n:n + m], t[n + m]
    Dpi = (lambda y, tr=False: cones.Dpi_apply(v, cone_types, cone_dims, y, tr)) if dpi is None else dpi.apply
    if not transpose:
        w = Dpi(t2)
        return np.concatenate([A.T @ w + c * t3, -(A @ t1) + t2 - w + b * t3,
                               [-(c @ t1) - b @ w]])
    r = A @ t1 - t2 - b * t3
    return np.concatenate([-(A.T @ t2) - c * t3, Dpi(r, True) + t2, [c @ t1 + b @ t2]])


def matrix_free_ops(A, b, c, x, s, y, cone_types, cone_dims):
    """(matvec, rmatvec, shape) of M for ``lsqr`` without forming M (large PSD cones: the reference's dense
    Dpi block of a 200 x 200 cone alone is 3.2 GB)."""
    A = sp.csr_matrix(A)
    m, n = A.shape
    v = np.asarray(y, float) - np.asarray(s, float)
    op = cones.DpiOperator(v, cone_types, cone_dims)
    N = n + m + 1
    return ((lambda t: M_apply(A, b, c, v, cone_types, cone_dims, t, False, op)),
            (lambda t: M_apply(A, b, c, v, cone_types, cone_dims, t, True, op)), (N, N))


def reverse_matrix_free(A, b, c, x, s, y, cone_types, cone_dims, dx_seed, **lsqr_kw):
    """``reverse`` (:349-373) on the matrix-free operator."""
    A = sp.csr_matrix(A)
    m, n = A.shape
    x = np.asarray(x, float)
    dx = np.asarray(dx_seed, float)
    dz = np.concatenate([dx, np.zeros(m), [-(x @ dx)]])
    if np.linalg.norm(dz) <= 1e-4:
        return np.zeros_like(dz)
    return lsqr(matrix_free_ops(A, b, c, x, s, y, cone_types, cone_dims), dz, **lsqr_kw)
