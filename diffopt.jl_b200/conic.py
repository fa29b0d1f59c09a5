"""Host-side mirror of the reference's ConicProgram backend (src/ConicProgram/ConicProgram.jl)
on top of the C ABI.  ``ConicModel`` keeps the reference's stored quantities (x, s, y), builds the
gradient cache on the device (``_gradient_cache`` :172-255 -> ``diffopt_b200_conic_setup``) and
exposes ``forward_differentiate!`` (:257-334), ``reverse_differentiate!`` (:336-394) and the
getters (:396-443).  No arithmetic of the path runs on the CPU."""
from __future__ import annotations

import time

import numpy as np

from ._capi import CONE_NONNEG, CONE_PSD, CONE_SOC, CONE_ZERO, HOST, Context, ptr  # noqa: F401
from .lsqr import DEFAULTS, julia_csc


class ConicModel:
    """Geometric conic form  A x + s = b, s in K  with the reference's conventions:
    ``A = -coefficients`` (:179-183), ``b = constants``, ``c`` negated for MAX sense (:206-208)."""

    def __init__(self, ctx: Context, A, b, c, cone_types, cone_dims):
        import scipy.sparse as sp
        self.ctx = ctx
        self.A = sp.csc_matrix(A, dtype=np.float64)
        self.m, self.n = self.A.shape
        self.b = np.ascontiguousarray(b, dtype=np.float64).reshape(self.m)
        self.c = np.ascontiguousarray(c, dtype=np.float64).reshape(self.n)
        self.cone_types = np.ascontiguousarray(cone_types, dtype=np.int32)
        self.cone_dims = np.ascontiguousarray(cone_dims, dtype=np.int64)
        self.x = np.full(self.n, np.nan)
        self.s = np.full(self.m, np.nan)
        self.y = np.full(self.m, np.nan)
        self.gradient_cache = False
        self.forw_grad_cache = None
        self.back_grad_cache = None
        self.diff_time = float("nan")
        self.tolerances = dict(DEFAULTS, maxiter=None)
        self.last_stats = None

    @classmethod
    def from_moi(cls, ctx, coefficients, constants, c, cone_types, cone_dims, max_sense=False):
        """From MOI-style data (function = coefficients x + constants in K), applying the
        reference's sign handling."""
        import scipy.sparse as sp
        return cls(ctx, -sp.csc_matrix(coefficients, dtype=np.float64), constants,
                   -np.asarray(c, float) if max_sense else c, cone_types, cone_dims)

    # MOI.set(model, VariablePrimalStart / ConstraintPrimalStart / ConstraintDualStart, ...)
    def set_variable_primal(self, x):
        self.x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.n)
        self.gradient_cache = False

    def set_constraint_primal(self, s):
        self.s = np.ascontiguousarray(s, dtype=np.float64).reshape(self.m)
        self.gradient_cache = False

    def set_constraint_dual(self, y):
        self.y = np.ascontiguousarray(y, dtype=np.float64).reshape(self.m)
        self.gradient_cache = False

    def _gradient_cache(self):
        if self.gradient_cache:
            return
        if np.isnan(self.y).any():   # ConicProgram.jl:186-196
            raise ValueError("Some constraints are missing a value for the `ConstraintDualStart` attribute.")
        if np.isnan(self.s).any():
            raise ValueError("Some constraints are missing a value for the `ConstraintPrimalStart` attribute.")
        colptr, rowval, nzval = julia_csc(self.A)
        rc = self.ctx.lib.diffopt_b200_conic_setup(
            self.ctx.h, self.n, self.m, ptr(colptr), ptr(rowval), ptr(nzval), ptr(self.b), ptr(self.c),
            ptr(self.x), ptr(self.s), ptr(self.y), len(self.cone_types), ptr(self.cone_types),
            ptr(self.cone_dims), HOST)
        self.ctx.check(rc)
        self.setup_ms = self.ctx.last_kernel_ms
        self.gradient_cache = True

    def _tol(self):
        t = self.tolerances
        return t["atol"], t["btol"], t["conlim"], 0 if t["maxiter"] is None else int(t["maxiter"])

    def vp(self):
        self._gradient_cache()
        out = np.empty(self.m)
        self.ctx.check(self.ctx.lib.diffopt_b200_conic_get_vp(self.ctx.h, ptr(out), HOST))
        return out

    def dpi_apply(self, t, transpose=False):
        self._gradient_cache()
        t = np.ascontiguousarray(t, dtype=np.float64)
        out = np.empty(self.m)
        self.ctx.check(self.ctx.lib.diffopt_b200_conic_dpi_apply(self.ctx.h, ptr(t), int(transpose), ptr(out), HOST))
        return out

    def M_apply(self, t, transpose=False):
        self._gradient_cache()
        t = np.ascontiguousarray(t, dtype=np.float64)
        out = np.empty(self.n + self.m + 1)
        self.ctx.check(self.ctx.lib.diffopt_b200_conic_M_apply(self.ctx.h, ptr(t), int(transpose), ptr(out), HOST))
        return out

    def forward_differentiate(self, dA=None, db=None, dc=None):
        """dA: scipy sparse / dense (m x n) perturbation of the constraint *coefficients as set through
        ForwardConstraintFunction* (used un-negated, :296-305); db: constants; dc: objective."""
        import scipy.sparse as sp
        t0 = time.perf_counter()
        self._gradient_cache()
        if dA is None:
            rows = cols = vals = None
            nnz = 0
        else:
            coo = sp.coo_matrix(dA)
            rows = np.ascontiguousarray(coo.row.astype(np.int64) + 1)
            cols = np.ascontiguousarray(coo.col.astype(np.int64) + 1)
            vals = np.ascontiguousarray(coo.data, dtype=np.float64)
            nnz = vals.size
        db = None if db is None else np.ascontiguousarray(db, dtype=np.float64).reshape(self.m)
        dc = None if dc is None else np.ascontiguousarray(dc, dtype=np.float64).reshape(self.n)
        dx = np.empty(self.n)
        dz = np.empty(self.n + self.m + 1)
        stats = np.zeros(4)
        rc = self.ctx.lib.diffopt_b200_conic_forward(self.ctx.h, nnz, ptr(rows), ptr(cols), ptr(vals), ptr(db), ptr(dc),
                                                     *self._tol(), ptr(dx), ptr(dz), ptr(stats), HOST)
        self.ctx.check(rc)
        n, m = self.n, self.m
        self.forw_grad_cache = (dz[:n], dz[n:n + m], dz[n + m:], dx)
        self.last_stats = dict(istop=int(stats[0]), itn=int(stats[1]), rnorm=stats[2], arnorm=stats[3],
                               kernel_ms=self.ctx.last_kernel_ms)
        self.diff_time = time.perf_counter() - t0

    def reverse_differentiate(self, dx_seed):
        t0 = time.perf_counter()
        self._gradient_cache()
        seed = np.ascontiguousarray(dx_seed, dtype=np.float64).reshape(self.n)
        g = np.empty(self.n + self.m + 1)
        dc = np.empty(self.n)
        db = np.empty(self.m)
        stats = np.zeros(4)
        rc = self.ctx.lib.diffopt_b200_conic_reverse(self.ctx.h, ptr(seed), *self._tol(), ptr(g), ptr(dc), ptr(db),
                                                     ptr(stats), HOST)
        self.ctx.check(rc)
        self.back_grad_cache = dict(g=g, dc=dc, db=db)
        self.last_stats = dict(istop=int(stats[0]), itn=int(stats[1]), rnorm=stats[2], arnorm=stats[3],
                               kernel_ms=self.ctx.last_kernel_ms)
        self.diff_time = time.perf_counter() - t0

    # getters ------------------------------------------------------------------------------------
    def forward_variable_primal(self):
        return self.forw_grad_cache[3]                    # -(du - x dw), :403-412

    def reverse_objective_function(self):
        return self.back_grad_cache["dc"]                 # :396-401

    def get_db(self, rows=None):
        db = self.back_grad_cache["db"]                   # :414-428
        return db if rows is None else db[rows]

    def get_dA(self, rows):
        """g[n+I] x' - vp[I] g[1:n]'  (:430-443): an outer-product view the reference also forms lazily,
        assembled here from the device results for the requested rows only."""
        g = self.back_grad_cache["g"]
        n = self.n
        vp = self.vp()
        rows = np.atleast_1d(rows)
        return np.outer(g[n + rows], self.x) - np.outer(vp[rows], g[:n])


class ConicBatch:
    """Lock-step batch of ``ConicModel``s of equal size (``diffopt_b200_conic_batch_*``): the reference differentiates one
    problem per ``reverse_differentiate!`` call (ConicProgram.jl:336-394); a training loop makes that call once per
    sample of a minibatch.  Here every problem is analysed as usual and ONE persistent kernel advances all the LSQR
    solves -- each with its own operator, stop tests and iteration count."""

    def __init__(self, ctx: Context, models, ctas_per_problem=1):
        self.ctx = ctx
        self.models = list(models)
        self.B = len(self.models)
        self.n, self.m = self.models[0].n, self.models[0].m
        self.tolerances = dict(DEFAULTS, maxiter=None)
        ctx.check(ctx.lib.diffopt_b200_conic_batch_begin(ctx.h, self.B, int(ctas_per_problem)))
        for mdl in self.models:
            colptr, rowval, nzval = julia_csc(mdl.A)
            ctx.check(ctx.lib.diffopt_b200_conic_batch_add(
                ctx.h, mdl.n, mdl.m, ptr(colptr), ptr(rowval), ptr(nzval), ptr(mdl.b), ptr(mdl.c), ptr(mdl.x), ptr(mdl.s),
                ptr(mdl.y), len(mdl.cone_types), ptr(mdl.cone_types), ptr(mdl.cone_dims), HOST))

    def reverse_differentiate(self, dx_seeds):
        """dx_seeds: (B, n).  Returns dict(g (B, n+m+1), dc (B, n), db (B, m), stats (B, 4))."""
        B, n, m = self.B, self.n, self.m
        seeds = np.ascontiguousarray(dx_seeds, dtype=np.float64).reshape(B, n)
        g = np.empty((B, n + m + 1))
        dc = np.empty((B, n))
        db = np.empty((B, m))
        stats = np.zeros((B, 4))
        t = self.tolerances
        rc = self.ctx.lib.diffopt_b200_conic_batch_reverse(
            self.ctx.h, ptr(seeds), t["atol"], t["btol"], t["conlim"], 0 if t["maxiter"] is None else int(t["maxiter"]),
            ptr(g), ptr(dc), ptr(db), ptr(stats), HOST)
        self.ctx.check(rc)
        self.kernel_ms = self.ctx.last_kernel_ms
        return dict(g=g, dc=dc, db=db, stats=stats)
