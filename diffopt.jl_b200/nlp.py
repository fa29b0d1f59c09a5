"""Host-side mirror of the NonLinearProgram backend's linear algebra (src/NonLinearProgram/NonLinearProgram.jl:356-435,
nlp_utilities.jl:436-444) on top of the C ABI: the sparse LU of the KKT Jacobian ``M`` with the reference's inertia
correction, and ``ds = -(K \\ N)`` for all parameter columns against that one factorisation.  Assembling ``M`` and ``N``
(derivative evaluation through MOI.Nonlinear) stays host code, as SURVEY.md 8(f) says."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import HOST, Context, ptr
from .lsqr import julia_csc


class InertiaCorrectedLU:
    """``_lu_with_inertia_correction(M, model, st, max_corrections)`` (:402-435): ``num_w`` = primal variables + slack
    variables of the inequality constraints, ``num_cons`` = constraints.  ``K is None`` in the reference (correction
    failed) is ``self.failed`` here."""

    def __init__(self, ctx: Context, M, num_w, num_cons, st=1e-6, max_corrections=50):
        self.ctx = ctx
        self.N = M.shape[0]
        colptr, rowval, nzval = julia_csc(M)
        nc = C.c_int32(0)
        rc = ctx.lib.diffopt_b200_sparse_setup_inertia(ctx.h, self.N, ptr(colptr), ptr(rowval), ptr(nzval), int(num_w), int(num_cons),
                                                       float(st), int(max_corrections), C.byref(nc))
        ctx.check(rc)
        self.corrections = int(nc.value)
        self.failed = rc > 0
        self.factor_ms = ctx.last_kernel_ms

    def solve(self, N):
        R = np.asfortranarray(np.asarray(N, dtype=np.float64).reshape(self.N, -1))
        X = np.empty_like(R, order="F")
        self.ctx.check(self.ctx.lib.diffopt_b200_sparse_solve(self.ctx.h, R.shape[1], ptr(R), ptr(X), HOST))
        self.solve_ms = self.ctx.last_kernel_ms
        return X


def compute_sensitivity(ctx, M, N, num_w, num_cons, st=1e-6, max_corrections=50):
    """``_compute_derivatives_no_relax`` tail (nlp_utilities.jl:436-447): ``ds = -(K \\ N)``, zeros when the inertia
    correction failed.  Returns (ds, K)."""
    K = InertiaCorrectedLU(ctx, M, num_w, num_cons, st, max_corrections)
    Nd = np.asarray(N.todense() if hasattr(N, "todense") else N, dtype=np.float64)
    if K.failed:
        return np.zeros((M.shape[0], Nd.shape[1])), K
    return -K.solve(Nd), K
