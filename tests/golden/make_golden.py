#!/usr/bin/env python
"""Writes tests/golden/kat.json: the known-answer vectors that pin the oracle.

Run in the BUILD container only (needs /root/reference for the on-disk fixture
``test/data/*.txt``); the GPU box never runs this.  Every case transcribes the
inputs and the expected literals of one of the reference's own tests
(file:line cited per case).  Primal/dual solutions that the reference obtains
from HiGHS/Ipopt/SCS are written here in closed form (each test in
``tests/test_oracle_kat.py`` re-checks that they satisfy the KKT conditions, so
no solver is needed).
"""
import json
import os
import sys

import numpy as np

REF = "/root/reference/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json")
r2 = np.sqrt(2.0)

cases = {}

# ---------------------------------------------------------------- QP, direct solve
cases["qp_moi_examples_2"] = dict(
    cite="test/quadratic_program.jl:232-293",
    Q=[[4, 1], [1, 2]], q=[1, 1], G=[[-1, 0], [0, -1]], h=[0, 0], A=[[1, 1]], b=[1],
    z=[0.25, 0.75], lam=[0, 0], nu=[-2.75],
    seed=[1.3, 0.5],
    exp=dict(grad_z=[-0.2, 0.2], grad_lam=[0.8, -0.8 / 3], grad_nu=[-0.7],
             dQ=[[-0.05, -0.05], [-0.05, 0.15]], dq=[-0.2, 0.2],
             dG=[[0, 0], [0, 0]], dh=[0, 0], dA=[[0.375, -1.075]], db=[0.7]),
    fwd=dict(dQ=[[-0.05, -0.05], [-0.05, 0.15]], dq=[-0.2, 0.2], dG=[[0, 0], [0, 0]],
             dh=[0, 0], dA=[[0.375, -1.075]], db=[0.7]),
    exp_fwd=dict(dz=[1.4875, -0.075], rhs_z=[-1.28125, 3.25625]),
    tol=2e-4,
)
cases["qp_moi_examples_1"] = dict(
    cite="test/quadratic_program.jl:181-227",
    Q=[[2, 1, 0], [1, 2, 1], [0, 1, 2]], q=[0, 0, 0],
    G=[[-1, -2, -3], [-1, -1, 0]], h=[-4, -1], A=[], b=[],
    z=[4 / 7, 3 / 7, 6 / 7], lam=[5 / 7, 6 / 7], nu=[],
    seed=[1, 1, 1],
    exp=dict(dQ=[[-0.12244895, 0.01530609, -0.11224488],
                 [0.01530609, 0.09183674, 0.07653058],
                 [-0.11224488, 0.07653058, -0.06122449]],
             dq=[-0.2142857, 0.21428567, -0.07142857],
             dG=[[0.05102692, 0.30612244, 0.25510856],
                 [0.06120519, 0.36734693, 0.30610315]],
             dh=[-0.35714284, -0.4285714]),
    tol=2e-4,
)
cases["qp_ineq_eq"] = dict(
    cite="test/quadratic_program.jl:131-176",
    Q=[[1, -1, 1], [-1, 2, -2], [1, -2, 4]], q=[2, -3, 1],
    G=[[0, 0, 1], [0, 1, 0], [1, 0, 0], [0, 0, -1], [0, -1, 0], [-1, 0, 0]],
    h=[1, 1, 1, 0, 0, 0], A=[[1, 1, 1]], b=[0.5],
    z=[0, 0.5, 0], lam=[0, 0, 0, 2, 0, 3.5], nu=[2],
    seed=[1, 1, 1],
    exp=dict(dQ=np.zeros((3, 3)).tolist(), dq=[0, 0, 0], dG=np.zeros((6, 3)).tolist(),
             dh=[0] * 6, dA=[[0, -0.5, 0]], db=[1.0]),
    tol=2e-4,
)
cases["qp_trivial_1"] = dict(
    cite="test/quadratic_program.jl:62-91; docs/src/examples/matrix-inversion-manual.jl:84,153-154",
    Q=[[4, 1], [1, 2]], q=[1, 1], G=[[1, 1]], h=[-1], A=[], b=[],
    z=[-0.25, -0.75], lam=[0.75], nu=[],
    seed=[1, 1],
    exp=dict(dh=[1.0]),
    # forward: only the constant of the constraint moves by one (dh = 1) -> dx = [.25,.75]
    fwd=dict(dQ=[[0, 0], [0, 0]], dq=[0, 0], dG=[[0, 0]], dh=[1.0], dA=[], db=[]),
    exp_fwd=dict(dz=[0.25, 0.75]),
    tol=2e-4,
)

# on-disk fixture (n=10, 25 inequalities all inactive, 10 equalities)
if os.path.isdir(REF):
    rd = lambda name: np.loadtxt(os.path.join(REF, "data", name + ".txt"))
    P, q, G, h, A, b = (rd(k) for k in ["P", "q", "G", "h", "A", "b"])
    z = np.linalg.solve(A, b)
    assert np.max(G @ z - h) < 0, "fixture: all inequalities inactive"
    nu = -np.linalg.solve(A.T, P @ z + q)
    cases["qp_fixture_data"] = dict(
        cite="test/quadratic_program.jl:295-350 (+ test/jump.jl:233-288), test/data/*.txt",
        Q=P.tolist(), q=q.tolist(), G=G.tolist(), h=h.tolist(), A=A.tolist(), b=b.tolist(),
        z=z.tolist(), lam=[0.0] * 25, nu=nu.tolist(), seed=[1.0] * 10,
        exp=dict(dq=rd("dq").tolist(), dh=rd("dh").tolist(), db=rd("db").tolist(),
                 dA=rd("dA").tolist(), dG=rd("dG").tolist(), dQ=rd("dP").tolist()),
        tol=1e-3, note="reference checks only dq, dh, db of the files (low accuracy, tol 1e-3)",
    )
else:
    print("WARNING: /root/reference missing; fixture case not regenerated", file=sys.stderr)

# ---------------------------------------------------------------- LP, LSQR on KKT
G6 = [[3, 2, 1], [2, 5, 3], [-1, 0, 0], [0, -1, 0], [0, 0, -1]]
cases["lp_simplex_example"] = dict(
    cite="test/linear_program.jl:70-102",
    Q=np.zeros((3, 3)).tolist(), q=[-2, -3, -4], G=G6, h=[10, 15, 0, 0, 0], A=[], b=[],
    z=[0, 0, 5], lam=[0, 4 / 3, 2 / 3, 11 / 3, 0], nu=[], seed=[1, 1, 1],
    exp=dict(dq=[0, 0, 0], dh=[0, 1 / 3, -1 / 3, 2 / 3, 0],
             dG=[[0, 0, 0], [0, 0, -5 / 3], [0, 0, 5 / 3], [0, 0, -10 / 3], [0, 0, 0]]),
    tol=1e-2,
)
cases["lp_fixed_variable"] = dict(
    cite="test/linear_program.jl:147-176",
    Q=np.zeros((3, 3)).tolist(), q=[-2, -3, -4],
    G=[[3, 2, 1], [2, 5, 3], [0, -1, 0], [0, 0, -1]], h=[10, 15, 0, 0],
    A=[[1, 0, 0]], b=[0],
    z=[0, 0, 5], lam=[0, 4 / 3, 11 / 3, 0], nu=[-2 / 3], seed=[1, 1, 1],
    exp=dict(dq=[0, 0, 0], dh=[0, 1 / 3, 2 / 3, 0],
             dG=[[0, 0, 0], [0, 0, -5 / 3], [0, 0, -10 / 3], [0, 0, 0]],
             dA=[[0, 0, -5 / 3]], db=[1 / 3]),
    tol=1e-2,
)
cases["lp_nonactive"] = dict(
    cite="test/linear_program.jl:223-246 and :30-48",
    Q=[[0]], q=[1], G=[[-1], [-1]], h=[0, -3], A=[], b=[],
    z=[3], lam=[0, 1], nu=[], seed=[-1],
    exp=dict(dh=[0, 1], grad_z=[0], grad_lam=[0, -1]),
    fwd=dict(dQ=[[0]], dq=[0], dG=[[0], [0]], dh=[0, 1], dA=[], db=[]),
    exp_fwd=dict(dz=[-1]),
    seed2=[1], exp2=dict(dG=[[0], [3]], dh=[0, -1]),
    tol=1e-2,
)

# KAT 8: dispatch LP of test/jump.jl:473-638 (`test_sensitivity_index_issue`).  Variables (v, u, g1, g2, g3, a);
# HiGHS' solution z = (0, 45, 15, 20, 20, 50) is asserted there (:540-543); the duals follow from stationarity
# c + G'lam + A'nu = 0 with lam = 0 on the inactive rows (the vertex is nondegenerate: 4 active inequalities + 2
# equalities = 6 variables, all active multipliers > 0, so the KKT matrix is nonsingular and the solver's duals are these).
# The reference perturbs the constant of each of the 13 constraints by +1 (ForwardConstraintFunction = 1.0, :606-613) and
# compares -dz with +-(column i of lsqr(KKT, rhsKKT)) (:624-636).  In the packed convention of
# QuadraticProgram.jl:374-395 (dh, db = -constant; GreaterThan rows arrive negated, :378-425): rows 1-6 and 11 are
# `>=` constraints -> dh_i = +1; rows 7-10 are `<=` -> dh_i = -1; the equalities -> db_j = -1.
_G8 = [[-1, 0, 0, 0, 0, 0], [0, -1, 0, 0, 0, 0], [0, 0, -1, 0, 0, 0], [0, 0, 0, -1, 0, 0], [0, 0, 0, 0, -1, 0],
       [0, 0, 0, 0, 0, -1], [1, 0, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0],
       [-1, 0, 0, 0, 0, -1]]
cases["lp_dispatch_sensitivity"] = dict(
    cite="test/jump.jl:473-638",
    Q=np.zeros((6, 6)).tolist(), q=[0, 0, 3, 5, 7, 1], G=_G8, h=[0, 0, 0, 0, 0, 0, 50, 15, 20, 25, -50],
    A=[[1, 1, 0, 0, 0, 0], [0, 1, 1, 1, 1, 0]], b=[45, 100],
    z=[0, 45, 15, 20, 20, 50], lam=[6, 0, 0, 0, 0, 0, 0, 4, 2, 0, 1], nu=[7, -7],
    # per constraint i (reference order xRef, :588-602): packed direction entry and the sign s_i of
    # `-dprimal_dcons[:, i] ~ s_i * dprimal_dconsKKT[:, i]` (:624-636)
    directions=[dict(kind="dh", index=i, value=(+1 if i in (0, 1, 2, 3, 4, 5, 10) else -1),
                     kkt_sign=(-1 if i in (0, 1, 2, 3, 4, 5, 10) else +1)) for i in range(11)] +
               [dict(kind="db", index=j, value=-1, kkt_sign=+1) for j in range(2)],
    tol=2e-4,
)

# ---------------------------------------------------------------- conic (A = -coefficients, b = constants)
cases["conic_socp"] = dict(
    cite="test/conic_program.jl:29-116 (eq_vec = true)",
    # variables (x, y, t); rows: Zeros(1): 1 - t ; Nonneg(1): y - 1/sqrt2 ; SOC(3): (t, x, y)
    coefficients=[[0, 0, -1], [0, 1, 0], [0, 0, 1], [1, 0, 0], [0, 1, 0]],
    constants=[1, -1 / r2, 0, 0, 0], c=[1, 0, 0],
    cone_types=[0, 1, 2], cone_dims=[1, 1, 3],
    x=[-1 / r2, 1 / r2, 1], s=[0, 0, 1, -1 / r2, 1 / r2], y=[r2, 1, r2, 1, -1],
    fwd=[dict(dA=[[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 0, 0], [0, 0, 0]], db=[0] * 5,
              dc=[0, 0, 0], exp_dx=[1.12132144, 1 / r2, 1 / r2])],
    tol=2e-4,
)
cases["conic_psd2"] = dict(
    cite="test/conic_program.jl:134-210 (forward) and :801-844 (reverse)",
    # variables X (3); rows: Zeros(1): X[2] - 1 ; PSD(2): X
    coefficients=[[0, 1, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]],
    constants=[-1, 0, 0, 0], c=[1, 0, 1],
    cone_types=[0, 3], cone_dims=[1, 3],
    x=[1, 1, 1], s=[0, 1, 1, 1], y=[2, 1, -1, 1],
    fwd=[dict(dA=np.zeros((4, 3)).tolist(), db=[1, 0, 0, 0], dc=[0, 0, 0], exp_dx=[-1, -1, -1]),
         dict(dA=np.zeros((4, 3)).tolist(), db=[0, 0, 0, 0], dc=[-1, 0, 1], exp_dx=[1, 0, -1])],
    rev=[dict(seed=[1, 0, 0], exp_db_rows=[0], exp_db=[-1.0])],
    tol=2e-4,
)
cases["conic_psd3"] = dict(
    cite="test/conic_program.jl:581-647",
    # one variable x; PSD(3): (x, 1, x, 1, 1, x)
    coefficients=[[1], [0], [1], [0], [0], [1]], constants=[0, 1, 0, 1, 1, 0], c=[1],
    cone_types=[3], cone_dims=[6],
    x=[1], s=[1] * 6, y=[1 / 3, -1 / 6, 1 / 3, -1 / 6, -1 / 6, 1 / 3],
    fwd=[dict(dA=np.zeros((6, 1)).tolist(), db=[1] * 6, dc=[0], exp_dx=[-0.5]),
         dict(dA=np.zeros((6, 1)).tolist(), db=[0] * 6, dc=[1], exp_dx=[0.0])],
    tol=1e-2,
)


def _clean(o):
    if isinstance(o, dict):
        return {k: _clean(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_clean(v) for v in o]
    if isinstance(o, (np.floating, float)):
        return float(o)
    if isinstance(o, (np.integer, int)):
        return int(o)
    return o


with open(OUT, "w") as f:
    json.dump(_clean(cases), f, indent=1)
print("wrote", OUT, "with", len(cases), "cases")
