"""GPU parity tests of the two callers next to the hot path (through the C ABI):
  * NonLinearProgram factorisation with inertia correction + `K \\ N` (NonLinearProgram.jl:356-435, nlp_utilities.jl:436-447)
  * reverse-mode accumulation into parameters (src/parameters.jl:341-534) on the device-summed gradient block."""
import numpy as np
import pytest
import scipy.sparse as sp

import diffopt_b200
from oracle import nlp as onlp
from oracle import qp as oqp
from test_oracle_kat import nlp_inertia_kat_matrix, quadratic_rhs_case

pytestmark = pytest.mark.gpu
RTOL_DIRECT = 1e-8


@pytest.fixture(scope="module")
def ctx():
    return diffopt_b200.Context(0)


def test_kat15_inertia_correction_on_the_device(ctx):
    """The reference's singular KKT Jacobian (test/nlp_program.jl:767-795): the plain factorisation reports it, one
    correction repairs it, and solves against the corrected factorisation match SuperLU on the same corrected matrix."""
    nlp = diffopt_b200.submodule("nlp")
    lsq = diffopt_b200.submodule("lsqr")
    M = sp.csc_matrix(nlp_inertia_kat_matrix())
    with pytest.raises(diffopt_b200.SingularException):
        lsq.SparseFactorization(ctx, M)
    K = nlp.InertiaCorrectedLU(ctx, M, num_w=2, num_cons=3, st=1e-6, max_corrections=50)
    assert not K.failed and K.corrections == 1
    N = np.random.default_rng(0).standard_normal((5, 3))
    D = np.ones(5); D[2:] = -1.0
    import scipy.sparse.linalg as spla
    ref = spla.splu(sp.csc_matrix(M + 1e-6 * sp.diags(D))).solve(N)
    X = K.solve(N)
    assert (np.linalg.norm(X - ref, axis=0) / np.linalg.norm(ref, axis=0)).max() <= RTOL_DIRECT
    ds, _ = nlp.compute_sensitivity(ctx, M, N, 2, 3)
    assert np.allclose(ds, onlp.compute_sensitivity(M, N, 2, 3), rtol=1e-8, atol=0)


def test_regular_and_hopeless_matrices(ctx):
    nlp = diffopt_b200.submodule("nlp")
    R = sp.csc_matrix(nlp_inertia_kat_matrix() + np.eye(5))
    K = nlp.InertiaCorrectedLU(ctx, R, 2, 3)
    assert not K.failed and K.corrections == 0                       # factorised as it is (:417-420)
    Z = sp.csc_matrix((4, 4))                                        # J = c st D is never singular: one correction
    K = nlp.InertiaCorrectedLU(ctx, Z, 2, 2)
    assert not K.failed and K.corrections == 1
    # st = 0 never repairs anything: the reference gives up after max_corrections and returns zeros (:382-386, :440-442)
    M = sp.csc_matrix(nlp_inertia_kat_matrix())
    ds, K = nlp.compute_sensitivity(ctx, M, np.ones((5, 2)), 2, 3, st=0.0, max_corrections=4)
    assert K.failed and K.corrections == 4 and not ds.any()


def test_kkt_jacobian_with_dependent_constraints(ctx):
    """A larger NLP-style KKT Jacobian [W J'; J 0] with two constraints whose Jacobian rows are empty (structurally
    singular, so every LU reports it -- dependent rows that cancel only up to rounding are caught or not depending on the
    pivot order, in UMFPACK as much as here): corrected factorisation and ds = -(K \\ N) for 40 parameter columns vs
    the oracle."""
    nlp = diffopt_b200.submodule("nlp")
    rng = np.random.default_rng(11)
    nw, nc = 300, 120
    W = sp.random(nw, nw, density=0.02, random_state=np.random.RandomState(3), data_rvs=rng.standard_normal)
    W = sp.csc_matrix(W @ W.T + sp.identity(nw) * 0.5)
    J = sp.random(nc - 2, nw, density=0.03, random_state=np.random.RandomState(4), data_rvs=rng.standard_normal).tocsr()
    J = sp.vstack([J, sp.csr_matrix((2, nw))]).tocsc()                # two constraints with an empty Jacobian row
    M = sp.bmat([[W, J.T], [J, None]], format="csc")
    N = rng.standard_normal((nw + nc, 40))
    ds, K = nlp.compute_sensitivity(ctx, M, N, nw, nc)
    ref = onlp.compute_sensitivity(M, N, nw, nc)
    assert not K.failed and K.corrections == onlp.lu_with_inertia_correction(M, nw, nc)[1] == 1
    # the corrected matrix has condition ~ 1 / st: compare through the residual of the system both sides solve
    D = np.ones(nw + nc); D[nw:] = -1.0
    Jc = M + 1e-6 * sp.diags(D)
    assert np.abs(Jc @ ds + N).max() <= 1e-8 * max(1.0, np.abs(N).max())
    assert (np.linalg.norm(ds - ref, axis=0) / np.linalg.norm(ref, axis=0)).max() <= 1e-6


@pytest.mark.parametrize("vals", [(2, 2, 2, 2, 2), (3, 2, 3, 2, 3)])
def test_kat16_parameter_pullback_through_the_device(ctx, vals):
    """test/parameters.jl:317-445: min 2x s.t. 11 t x >= 1 + 3 p q + 5 r^2 + 7 s, reverse seeds 0..3.  The QP backend
    differentiates the inner LP, the device sums the getters over a batch of seeds (shared parameters), and
    diffopt_b200_param_pullback maps that block to (p, q, r, s, t): closed-form answers of the reference's test."""
    qpm = diffopt_b200.submodule("qp")
    cases = [quadratic_rhs_case(*vals, dir_x) for dir_x in range(4)]
    d = {k: np.concatenate([c[0][k] for c in cases]) for k in cases[0][0]}
    _, rev, info = qpm.solve_batch(ctx, d["Q"], d["G"], None, d["h"], d["z"], d["lam"], None, seed=d["seed"])
    assert not info.any()
    terms = qpm.poi_terms(1, 1, 0, [dict(row=("ineq", 0), **cases[0][1])], param_values=np.array(vals, float))
    # per seed (batch of one) ...
    for b, (_, _, expect) in enumerate(cases):
        flat = qpm.shared_param_grads(ctx, d["z"][b:b + 1], d["lam"][b:b + 1], None, rev[b:b + 1], flat=True)
        got = qpm.parameter_pullback(ctx, terms, flat, 5)
        assert np.allclose(got, expect, atol=1e-10), (b, got, expect)
    # ... and accumulated over the batch, as a training loop's `+=` over samples does
    flat = qpm.shared_param_grads(ctx, d["z"], d["lam"], None, rev, flat=True)
    got = qpm.parameter_pullback(ctx, terms, flat, 5)
    assert np.allclose(got, sum(c[2] for c in cases), atol=1e-10)


def test_parameter_pullback_matches_oracle_on_random_terms(ctx):
    """Random parametric terms over a headline-shaped batch: device pull-back of the device-summed block vs the
    reference's loops (oracle.nlp.reverse_parameters) fed with the oracle's own gradients."""
    import bench_data
    qpm = diffopt_b200.submodule("qp")
    n, m, p, B, P = 64, 64, 16, 24, 9
    d = bench_data.qp_batch(B, n, m, p, seed0=515)
    _, rev, info = qpm.solve_batch(ctx, d["Q"], d["G"], d["A"], d["h"], d["z"], d["lam"], d["nu"], seed=d["seed"])
    assert not info.any()
    rng = np.random.default_rng(7)
    pvals = rng.standard_normal(P)
    cons = []
    for kind, rows in (("ineq", m), ("eq", p)):
        for i in rng.choice(rows, size=6, replace=False):
            cons.append(dict(row=(kind, int(i)), p=[(int(rng.integers(P)), float(rng.standard_normal())) for _ in range(2)],
                             pp=[(int(rng.integers(P)), int(rng.integers(P)), float(rng.standard_normal()))],
                             pv=[(int(rng.integers(P)), int(rng.integers(n)), float(rng.standard_normal())) for _ in range(3)]))
    obj = dict(pv=[(int(rng.integers(P)), int(rng.integers(n)), float(rng.standard_normal())) for _ in range(5)])
    flat = qpm.shared_param_grads(ctx, d["z"], d["lam"], d["nu"], rev, flat=True)
    got = qpm.parameter_pullback(ctx, qpm.poi_terms(n, m, p, cons, obj, pvals), flat, P)
    want = np.zeros(P)
    for b in range(B):
        dz, dlam, dnu = oqp.reverse(d["Q"][b], d["G"][b], d["h"][b], d["A"][b], d["z"][b], d["lam"][b], d["nu"][b], d["seed"][b])
        dQ, dq, dG, dh, dA, db = oqp.reverse_param_grads(d["z"][b], d["lam"][b], d["nu"][b], dz, dlam, dnu)
        ocons = []
        for c in cons:
            kind, i = c["row"]
            ocons.append(dict(grad_cte=-(dh[i] if kind == "ineq" else db[i]),
                              grad_coef={v: (dG if kind == "ineq" else dA)[i, v] for v in range(n)}, p=c["p"], pp=c["pp"], pv=c["pv"]))
        oobj = dict(grad_cte=0.0, grad_coef={v: dq[v] for v in range(n)}, pv=obj["pv"])
        want += onlp.reverse_parameters(P, ocons, oobj, pvals)
    assert np.linalg.norm(got - want) <= 1e-8 * np.linalg.norm(want)
