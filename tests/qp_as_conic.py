"""Test infrastructure: a solved QP restated as a solved conic program, the way the reference's MOI bridges hand a QP to
``ConicProgram.Model`` in its cross-backend check (``test/utils.jl:369-377``: every ``qp_test`` also runs with
``DiffOpt.ConicProgram.Model``).

    min 1/2 z'Qz + q'z  s.t.  G z <= h,  A z = b          (solution z, lam = -dual(<=), nu = -dual(==))

becomes, with variables (z, t) and Q = U'U:

    min t   s.t.   A z - b                        in Zeros(p)
                   h - G z                        in Nonnegatives(m)
                   ((1 + t - q'z)/sqrt2, (1 - t + q'z)/sqrt2, U z)  in SecondOrderCone(n + 2)   [only when Q != 0]

(the epigraph / QuadtoSOC / RSOCtoSOC chain of ``src/bridges.jl:246-323`` and MOI's own bridges).  For Q == 0 the
objective stays q'z and there is no t.  The conic solution follows from the QP's: slacks are the function values, duals
are y_eq = -nu, y_le = lam, y_soc = (s0, -s1, -U z) (complementary to s on the cone boundary, scaled by the stationarity
of t).  ``A``/``b``/``c`` are returned in the reference's conic convention A = -coefficients, b = constants
(ConicProgram.jl:179-183).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from oracle import cones as oc


def qp_as_conic(Q, q, G, h, A, b, z, lam, nu):
    Q = np.atleast_2d(np.asarray(Q, float))
    n = Q.shape[0]
    q, z = np.asarray(q, float).reshape(n), np.asarray(z, float).reshape(n)
    G = np.asarray(G, float).reshape(-1, n)
    A = np.asarray(A, float).reshape(-1, n)
    h, lam = np.asarray(h, float).reshape(-1), np.asarray(lam, float).reshape(-1)
    b, nu = np.asarray(b, float).reshape(-1), np.asarray(nu, float).reshape(-1)
    m, p = G.shape[0], A.shape[0]
    quad = bool(np.any(Q != 0.0))
    nv = n + (1 if quad else 0)
    coef, const, s, y, types, dims = [], [], [], [], [], []
    if p:
        coef.append(np.hstack([A, np.zeros((p, nv - n))])); const.append(-b)
        s.append(A @ z - b); y.append(-nu); types.append(oc.ZERO); dims.append(p)
    if m:
        coef.append(np.hstack([-G, np.zeros((m, nv - n))])); const.append(h)
        s.append(h - G @ z); y.append(lam); types.append(oc.NONNEG); dims.append(m)
    if quad:
        w, V = np.linalg.eigh(Q)
        U = (V * np.sqrt(np.maximum(w, 0.0))).T          # Q = U'U (any square root serves)
        t = 0.5 * z @ Q @ z + q @ z
        r2 = np.sqrt(2.0)
        C = np.zeros((n + 2, nv))
        C[0, :n], C[0, n] = -q / r2, 1 / r2
        C[1, :n], C[1, n] = q / r2, -1 / r2
        C[2:, :n] = U
        k = np.concatenate([[1 / r2, 1 / r2], np.zeros(n)])
        x = np.concatenate([z, [t]])
        ss = C @ x + k
        coef.append(C); const.append(k); s.append(ss)
        y.append(np.concatenate([[ss[0], -ss[1]], -ss[2:]]))
        types.append(oc.SOC); dims.append(n + 2)
        c = np.concatenate([np.zeros(n), [1.0]])
    else:
        x = z.copy()
        c = q.copy()
    coef = np.vstack(coef)
    return dict(A=sp.csc_matrix(-coef), b=np.concatenate(const), c=c, x=x, s=np.concatenate(s), y=np.concatenate(y),
                cone_types=types, cone_dims=dims, n=n, m=m, p=p, quad=quad)


def conic_forward_direction(cp, dq=None, dG=None, dh=None, dA=None, db=None):
    """QP forward direction (dq, dG, dh, dA, db in the reference's packed convention: constraint functions
    dG z - dh, dA z - db) as the (dA, db, dc) arguments of the conic ``forward_differentiate!``: ``dA`` = perturbation
    of the constraint COEFFICIENTS as the bridges pass it on (the reference uses it un-negated, ConicProgram.jl:296-305),
    ``db`` = perturbation of the constants, ``dc`` of the objective.  dQ has no linear image in this form."""
    n, m, p = cp["n"], cp["m"], cp["p"]
    nv = cp["x"].size
    zero = lambda shape: np.zeros(shape)
    dq = zero(n) if dq is None else np.asarray(dq, float)
    dG = zero((m, n)) if dG is None else np.asarray(dG, float).reshape(m, n)
    dA = zero((p, n)) if dA is None else np.asarray(dA, float).reshape(p, n)
    dh = zero(m) if dh is None else np.asarray(dh, float)
    db = zero(p) if db is None else np.asarray(db, float)
    pad = lambda M: np.hstack([M, np.zeros((M.shape[0], nv - n))])
    rows_c, rows_k = [], []
    if p:
        rows_c.append(pad(dA)); rows_k.append(-db)        # A z - b
    if m:
        rows_c.append(pad(-dG)); rows_k.append(dh)        # h - G z
    if cp["quad"]:
        r2 = np.sqrt(2.0)
        C = np.zeros((n + 2, nv))
        C[0, :n], C[1, :n] = -dq / r2, dq / r2
        rows_c.append(C); rows_k.append(np.zeros(n + 2))
        dc = np.zeros(nv)
    else:
        dc = dq
    return sp.csc_matrix(np.vstack(rows_c)), np.concatenate(rows_k), dc
