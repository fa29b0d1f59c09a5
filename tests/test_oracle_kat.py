"""Pins the CPU oracle against the reference's own known-answer tests.

Tolerances are the reference's (``atol = rtol = 2e-4`` QP/conic, ``1e-2`` LP, ``1e-3``
fixture), applied like Julia's ``isapprox`` on arrays: ``|x-y| <= max(atol, rtol*max(|x|,|y|))``
in the 2-norm.
"""
import numpy as np
import pytest

from oracle import cones, conic, lsqr, qp


def approx(x, y, tol):
    x = np.asarray(x, float).ravel()
    y = np.asarray(y, float).ravel()
    return np.linalg.norm(x - y) <= max(tol, tol * max(np.linalg.norm(x), np.linalg.norm(y)))


def _qp_arrays(c):
    n, m, p = len(c["z"]), len(c["lam"]), len(c["nu"])
    a = lambda k, shape: np.array(c[k], float).reshape(shape)
    return (a("Q", (n, n)), a("q", n), a("G", (m, n)), a("h", m), a("A", (p, n)), a("b", p),
            a("z", n), a("lam", m), a("nu", p))


QP_CASES = ["qp_moi_examples_2", "qp_moi_examples_1", "qp_ineq_eq", "qp_trivial_1",
            "qp_fixture_data", "lp_simplex_example", "lp_fixed_variable", "lp_nonactive"]
# (KAT 8, "lp_dispatch_sensitivity", has no reverse literals: it has its own test below)


@pytest.mark.parametrize("name", QP_CASES)
def test_qp_kat(kat, name):
    c = kat[name]
    Q, q, G, h, A, b, z, lam, nu = _qp_arrays(c)
    n, m, p = z.size, lam.size, nu.size
    # the closed-form primal/dual really is a KKT point of the reference's problem
    assert np.abs(Q @ z + q + G.T @ lam + A.T @ nu).max() < 1e-12
    assert np.all(G @ z - h <= 1e-12) and np.all(lam >= 0)
    assert np.abs(lam * (G @ z - h)).max(initial=0) < 1e-12
    assert np.abs(A @ z - b).max(initial=0) < 1e-12
    # LP cases take the reference's LSQR branch (QuadraticProgram.jl:333)
    assert qp.is_iterative(Q) == name.startswith("lp")
    tol = c["tol"]
    dz, dl, dn = qp.reverse(Q, G, h, A, z, lam, nu, np.array(c["seed"], float))
    got = dict(zip(["dQ", "dq", "dG", "dh", "dA", "db"], qp.reverse_param_grads(z, lam, nu, dz, dl, dn)))
    got.update(grad_z=dz, grad_lam=dl, grad_nu=dn)
    checked = c["exp"].keys()
    if name == "qp_fixture_data":
        checked = ["dq", "dh", "db"]  # what test/quadratic_program.jl:326-347 checks
    for k in checked:
        assert approx(got[k], c["exp"][k], tol), k
    if "fwd" in c:
        f = c["fwd"]
        a = lambda k, shape: np.array(f[k], float).reshape(shape)
        args = (a("dQ", (n, n)), a("dq", n), a("dG", (m, n)), a("dh", m), a("dA", (p, n)), a("db", p))
        fz, fl, fn = qp.forward(Q, G, h, A, z, lam, nu, *args)
        assert approx(fz, c["exp_fwd"]["dz"], tol)
        if "rhs_z" in c["exp_fwd"]:
            assert approx(qp.forward_rhs(z, lam, nu, *args)[:n], c["exp_fwd"]["rhs_z"], tol)
        # forward/reverse inner-product identity (test/utils.jl:331-337), exact in the algebra
        rb = np.zeros(n + m + p)
        rb[:n] = c["seed"]
        rf = qp.forward_rhs(z, lam, nu, *args)
        assert abs(fz @ rb[:n] - rf @ np.concatenate([dz, dl, dn])) < 1e-9 * (1 + abs(fz @ rb[:n]))
    if "seed2" in c:
        dz, dl, dn = qp.reverse(Q, G, h, A, z, lam, nu, np.array(c["seed2"], float))
        got = dict(zip(["dQ", "dq", "dG", "dh", "dA", "db"], qp.reverse_param_grads(z, lam, nu, dz, dl, dn)))
        for k, e in c["exp2"].items():
            assert approx(got[k], e, tol), k


def _kat8_manual_kkt(c):
    """The hand-built system of test/jump.jl:563-581: OptNet's K = [Q G' A'; D(lam) G D(Gz-h) 0; A 0 0] and the
    right-hand sides [0; D(lam); 0] / [0; 0; I], one column per constraint."""
    Q, q, G, h, A, b, z, lam, nu = _qp_arrays(c)
    n, m, p = z.size, lam.size, nu.size
    K = np.block([[Q, G.T, A.T], [np.diag(lam) @ G, np.diag(G @ z - h), np.zeros((m, p))],
                  [A, np.zeros((p, m)), np.zeros((p, p))]])
    R = np.block([[np.zeros((n, m)), np.zeros((n, p))], [np.diag(lam), np.zeros((m, p))],
                  [np.zeros((p, m)), np.eye(p)]])
    return K, R


def test_kat8_dispatch_lp_forward_directions(kat):
    """test/jump.jl:473-638: 13 successive forward directions (constant of every constraint + 1) against column-wise
    LSQR on the hand-built KKT system, with the reference's per-column sign table (:624-636)."""
    c = kat["lp_dispatch_sensitivity"]
    Q, q, G, h, A, b, z, lam, nu = _qp_arrays(c)
    n, m, p = z.size, lam.size, nu.size
    assert np.abs(Q @ z + q + G.T @ lam + A.T @ nu).max() < 1e-12 and np.abs(A @ z - b).max() < 1e-12
    assert np.all(G @ z - h <= 1e-12) and np.abs(lam * (G @ z - h)).max() < 1e-12
    assert np.all((lam > 0) == (np.abs(G @ z - h) < 1e-12))          # strict complementarity
    K, R = _kat8_manual_kkt(c)
    assert np.linalg.matrix_rank(K) == n + m + p                    # nondegenerate vertex: unique sensitivities
    for i, d in enumerate(c["directions"]):
        dh, db = np.zeros(m), np.zeros(p)
        (dh if d["kind"] == "dh" else db)[d["index"]] = d["value"]
        dz, _, _ = qp.forward(Q, G, h, A, z, lam, nu, np.zeros((n, n)), np.zeros(n), np.zeros((m, n)), dh,
                              np.zeros((p, n)), db)
        col = lsqr.lsqr(K, R[:, i])[:n]
        assert approx(-dz, d["kkt_sign"] * col, c["tol"]), i


def test_kat14_qp_cases_through_the_conic_backend(kat):
    """Cross-backend check of test/utils.jl:369-377 (every qp_test also runs with ConicProgram.Model) on the oracle:
    the QP restated as a conic program (tests/qp_as_conic.py) gives the QP backend's sensitivities for the directions
    the reference's conic path treats consistently (constants of the constraints, the linear objective).  Coefficient
    directions come out with the opposite sign because ConicProgram.jl:296-305 packs dA un-negated while A = -coefficients
    (the reference notes "conic finds different solutions", test/utils.jl:420); kept, and asserted as such."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from qp_as_conic import conic_forward_direction, qp_as_conic
    tight = dict(atol=1e-14, btol=1e-14, conlim=0.0, maxiter=5000)
    table = {"qp_moi_examples_2": ["dq", "db"], "qp_moi_examples_1": ["dq", "dh"], "qp_ineq_eq": ["dh", "db"],
             "qp_trivial_1": ["dq", "dh"], "lp_simplex_example": ["dh"], "lp_fixed_variable": ["dh", "db"],
             "lp_nonactive": ["dh"], "lp_dispatch_sensitivity": ["dq", "dh", "db"]}
    rng = np.random.default_rng(14)
    for name, keys in table.items():
        c = kat[name]
        Q, q, G, h, A, b, z, lam, nu = _qp_arrays(c)
        n, m, p = z.size, lam.size, nu.size
        cp = qp_as_conic(Q, q, G, h, A, b, z, lam, nu)
        assert np.abs(cp["A"] @ cp["x"] + cp["s"] - cp["b"]).max() < 1e-12      # primal feasibility A x + s = b
        assert np.abs(cp["A"].T @ cp["y"] + cp["c"]).max() < 1e-12              # dual feasibility A'y + c = 0
        assert abs(cp["s"] @ cp["y"]) < 1e-12                                    # complementarity
        cache = conic.gradient_cache(cp["A"], cp["b"], cp["c"], cp["x"], cp["s"], cp["y"], cp["cone_types"],
                                     cp["cone_dims"])
        kw_qp = tight if qp.is_iterative(Q) else {}
        for key in keys:
            size = dict(dq=n, dh=m, db=p)[key]
            direction = {key: rng.standard_normal(size)}
            full = dict(dQ=np.zeros((n, n)), dq=np.zeros(n), dG=np.zeros((m, n)), dh=np.zeros(m),
                        dA=np.zeros((p, n)), db=np.zeros(p))
            full.update(direction)
            ref = qp.forward(Q, G, h, A, z, lam, nu, **full, **kw_qp)[0]
            dAc, dbc, dcc = conic_forward_direction(cp, **direction)
            if key == "dq" and cp["quad"]:
                dAc = -dAc   # q sits in the SOC rows' coefficients: the reference's un-negated dA flips it
            dx, _ = conic.forward(cache, dAc, dbc, dcc, **tight)
            assert np.linalg.norm(dx[:n] - ref) <= 1e-8 * max(1.0, np.linalg.norm(ref)), (name, key)


def test_fixture_files_loose_agreement(kat):
    """dA.txt / db.txt agree to the fixture's own (low) accuracy — SURVEY.md §8c (5)."""
    c = kat["qp_fixture_data"]
    Q, q, G, h, A, b, z, lam, nu = _qp_arrays(c)
    dz, dl, dn = qp.reverse(Q, G, h, A, z, lam, nu, np.ones(10))
    g = qp.reverse_param_grads(z, lam, nu, dz, dl, dn)
    assert np.abs(g[4] - np.array(c["exp"]["dA"])).max() < 1e-2
    assert np.abs(g[5] - np.array(c["exp"]["db"])).max() < 1e-2


def _conic(c):
    A = -np.array(c["coefficients"], float)   # ConicProgram.jl:179-183
    b = np.array(c["constants"], float)
    cc = np.array(c["c"], float)
    x, s, y = (np.array(c[k], float) for k in "xsy")
    assert np.abs(A @ x + s - b).max() < 1e-12
    return A, b, cc, x, s, y, conic.gradient_cache(A, b, cc, x, s, y, c["cone_types"], c["cone_dims"])


@pytest.mark.parametrize("name", ["conic_socp", "conic_psd2", "conic_psd3"])
def test_conic_kat(kat, name):
    c = kat[name]
    A, b, cc, x, s, y, cache = _conic(c)
    for f in c.get("fwd", []):
        dx, _ = conic.forward(cache, np.array(f["dA"], float), f["db"], f["dc"])
        assert approx(dx, f["exp_dx"], c["tol"])
    for r in c.get("rev", []):
        g = conic.reverse(cache, r["seed"])
        _, db, _ = conic.reverse_param_grads(cache, g)
        assert approx(db[r["exp_db_rows"]], r["exp_db"], c["tol"])


@pytest.mark.parametrize("name", ["conic_socp", "conic_psd2", "conic_psd3"])
def test_matrix_free_M_equals_reference_M(kat, name):
    c = kat[name]
    A, b, cc, x, s, y, cache = _conic(c)
    Md = cache.M.toarray()
    rng = np.random.default_rng(0)
    for _ in range(3):
        t = rng.normal(size=Md.shape[0])
        assert np.allclose(conic.M_apply(A, b, cc, y - s, c["cone_types"], c["cone_dims"], t), Md @ t, atol=1e-13)
        assert np.allclose(conic.M_apply(A, b, cc, y - s, c["cone_types"], c["cone_dims"], t, True), Md.T @ t, atol=1e-13)


def test_lsqr_min_norm_limit_on_singular_M(kat):
    """M is singular by construction (rank N-1 on the SOCP test): LSQR from x0=0 must give pinv(M) g."""
    c = kat["conic_socp"]
    A, b, cc, x, s, y, cache = _conic(c)
    Md = cache.M.toarray()
    assert np.linalg.matrix_rank(Md) == Md.shape[0] - 1
    f = c["fwd"][0]
    _, dz = conic.forward(cache, np.array(f["dA"], float), f["db"], f["dc"], atol=1e-14, btol=1e-14)
    dA = np.array(f["dA"], float)
    g = np.concatenate([dA.T @ cache.vp, -dA @ x, [0.0]])
    assert np.allclose(dz, np.linalg.pinv(Md) @ g, atol=1e-9)


def test_lsqr_matches_scipy():
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(7)
    A = sp.random(300, 200, density=0.05, random_state=3, format="csr") + sp.eye(300, 200)
    b = rng.normal(size=300)
    x, info = lsqr.lsqr(A, b, return_info=True)
    ref = spla.lsqr(A, b, atol=lsqr.SQRT_EPS, btol=lsqr.SQRT_EPS, conlim=1 / lsqr.SQRT_EPS, iter_lim=300)
    assert info.itn == ref[2] and info.istop == ref[1]
    assert np.allclose(x, ref[0], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("ctype,k", [(cones.ZERO, 4), (cones.NONNEG, 7), (cones.SOC, 6), (cones.PSD, 10), (cones.PSD, 21)])
def test_operator_form_equals_dense_gradient(ctype, k):
    rng = np.random.default_rng(11 + k)
    for trial in range(6):
        v = rng.normal(size=k)
        if ctype == cones.SOC:
            v[0] = [0.1, 5.0, -5.0, 0.0, 0.3, -0.2][trial]   # boundary / interior / polar / generic
        D = cones.project_gradient(v, ctype)
        yv = rng.normal(size=k)
        assert np.allclose(cones.apply_gradient(v, ctype, yv), D @ yv, atol=1e-12)
        assert np.allclose(cones.apply_gradient(v, ctype, yv, transpose=True), D.T @ yv, atol=1e-12)


def test_psd_gradient_is_transposed_jacobian():
    """the reference's PSD block is the TRANSPOSE of d pi / d v (unscaled triangle)."""
    rng = np.random.default_rng(5)
    v = rng.normal(size=6)
    D = cones.project_gradient(v, cones.PSD)
    J = np.empty((6, 6))
    eps = 1e-6
    for j in range(6):
        e = np.zeros(6)
        e[j] = eps
        J[:, j] = (cones.project(v + e, cones.PSD) - cones.project(v - e, cones.PSD)) / (2 * eps)
    assert np.allclose(D.T, J, atol=1e-6)
    assert not np.allclose(D, D.T, atol=1e-3)


# ---- NLP factorisation with inertia correction, parameter pull-back -------------------------------

def nlp_inertia_kat_matrix():
    """The reference's own case, test/nlp_program.jl:767-795 (KAT 15): a KKT Jacobian whose first two columns are
    parallel (lambda2 = 0, mu = 0), so `lu(M; check = false)` reports status 1."""
    x1, x2 = 0.33, 0.33
    lambda1, lambda2 = 0.333, 0.00
    mu_val = 0.00
    return np.array([
        [0, 0, -1, -2, -1],
        [0, 0, -2, -1, 0],
        [-lambda1, -2 * lambda1, (1 - x1 - 2 * x2), 0, 0],
        [-2 * lambda2, -lambda2, 0, (1 - 2 * x1 - x2), 0],
        [mu_val, 0, 0, 0, x1]], dtype=float)


def test_kat15_inertia_correction():
    from oracle import nlp
    import scipy.sparse as sp
    M = sp.csc_matrix(nlp_inertia_kat_matrix())
    K, status = nlp._lu(M)
    assert status == 1                                   # `@assert K.status == 1 # Fail`
    K, nc = nlp.inertia_correction(M, 3, 2, st=1e-6, max_corrections=50)
    assert K is not None and nc == 1                     # `@test K.status == 0 # Success`
    K2, nc2 = nlp.lu_with_inertia_correction(M, 2, 3)
    assert K2 is not None and nc2 == 1
    # a regular matrix is factorised as it is
    R = sp.csc_matrix(nlp_inertia_kat_matrix() + np.eye(5))
    assert nlp.lu_with_inertia_correction(R, 2, 3)[1] == 0


def quadratic_rhs_case(pv, qv, rv, sv, tv, dir_x):
    """test/parameters.jl:317-445 (KAT 16):  min 2x  s.t.  11 t x >= 1 + 3 p q + 5 r^2 + 7 s.  Inner problem in the QP
    backend's form (n = 1, one inequality G z <= h): G = -11 t, h = -(1 + 3pq + 5r^2 + 7s).  Returns the QP data, the
    parametric terms of the constraint as POI holds them (function 11 t x - 3 p q - 5 r^2 - 7 s - 1 in GreaterThan(0),
    flipped to LessThan: -11 t x + 3 p q + 5 r^2 + 7 s + 1 <= 0; MOI's quadratic coefficient of r^2 is 2 x 5) and the
    expected parameter sensitivities."""
    c = 1 + 3 * pv * qv + 5 * rv ** 2 + 7 * sv
    d = dict(Q=np.zeros((1, 1, 1)), q=np.array([[2.0]]), G=np.array([[[-11.0 * tv]]]), h=np.array([[-c]]),
             A=np.zeros((1, 0, 1)), b=np.zeros((1, 0)), z=np.array([[c / (11 * tv)]]), lam=np.array([[2 / (11 * tv)]]),
             nu=np.zeros((1, 0)), seed=np.array([[float(dir_x)]]))
    terms = dict(p=[(3, 7.0)], pp=[(0, 1, 3.0), (2, 2, 10.0)], pv=[(4, 0, -11.0)])   # parameters p, q, r, s, t = 0..4
    expect = dir_x * np.array([3 * qv / (11 * tv), 3 * pv / (11 * tv), 10 * rv / (11 * tv), 7 / (11 * tv), -c / (11 * tv ** 2)])
    return d, terms, expect


@pytest.mark.parametrize("vals", [(2, 2, 2, 2, 2), (2, 3, 2, 3, 3), (3, 2, 3, 2, 2)])
def test_kat16_parameter_pullback(vals):
    from oracle import nlp
    for dir_x in range(4):
        d, terms, expect = quadratic_rhs_case(*vals, dir_x)
        dz, dlam, dnu = qp.reverse(d["Q"][0], d["G"][0], d["h"][0], d["A"][0], d["z"][0], d["lam"][0], d["nu"][0], d["seed"][0])
        dQ, dq, dG, dh, dA, db = qp.reverse_param_grads(d["z"][0], d["lam"][0], d["nu"][0], dz, dlam, dnu)
        # ReverseConstraintFunction of the LessThan row: coefficients dG, constant -dh (QuadraticProgram.jl:307-314)
        con = dict(grad_cte=-dh[0], grad_coef={0: dG[0, 0]}, **terms)
        got = nlp.reverse_parameters(5, [con], None, np.array(vals, float))
        assert np.allclose(got, expect, atol=1e-10), (vals, dir_x, got, expect)


def kat13_psd_pos_problem():
    """test/conic_program.jl:378-579 (KAT 13): 7 variables, Zeros(1) + Nonnegatives(1) + Nonnegatives(6) + PSD triangle(2)
    in the row order the reference's ProductOfSets gives them, MAX sense (c negated, :206-208), the solution literals the
    test asserts (:488-517) and the forward perturbation dA = ones(11, 7), db = ones(11), dc = ones(7) (:452-487, packed
    un-negated).  Expected dx: the diffcp values of :523, compared at atol 0.3 / rtol 0.01 as the reference does."""
    import scipy.sparse as sp
    de, al, r2 = 0.9, 0.8, np.sqrt(2.0)
    coef = np.zeros((11, 7))
    const = np.zeros(11)
    coef[1, :6] = -1.0
    const[1] = 10.0                                             # c1: eta - sum(x[1:6]) >= 0
    coef[2:8, :6] = np.eye(6)                                   # c2: x[1:6] >= 0
    coef[8, :] = [de / 2, al, de, de / 4, de / 8, 0.0, -1.0]    # c3: PSD triangle (1,1), (1,2), (2,2)
    coef[9, [0, 1, 2, 4, 5]] = [-de / (2 * r2), -de / 4, 0.0, -de / (8 * r2), 0.0]
    coef[10, [0, 1, 2, 4, 5, 6]] = [de / 2, de - al, 0.0, de / 8, de / 4, -1.0]
    x = np.array([20 / 3.0, 0.0, 10 / 3.0, 0.0, 0.0, 0.0, 1.90192379])
    s = np.array([0.0, 0.0, 20 / 3.0, 0.0, 10 / 3.0, 0.0, 0.0, 0.0, 4.09807621, -2.12132, 1.09807621])
    y = np.array([0.0, 0.19019238, 0.0, 0.12597667, 0.0, 0.14264428, 0.14264428, 0.01274047, 0.21132487, 0.408248, 0.78867513])
    return dict(A=-sp.csc_matrix(coef), b=const, c=-np.array([0, 0, 0, 0, 0, 0, 1.0]), x=x, s=s, y=y,
                cone_types=[cones.ZERO, cones.NONNEG, cones.NONNEG, cones.PSD], cone_dims=[1, 1, 6, 3],
                dA=np.ones((11, 7)), db=np.ones(11), dc=np.ones(7),
                dx=np.array([-39.6066, 10.8953, -14.9189, 10.9054, 10.883, 10.9118, -21.7508]))


def test_kat13_psd_and_pos_forward_vs_diffcp_literals():
    d = kat13_psd_pos_problem()
    cache = conic.gradient_cache(d["A"], d["b"], d["c"], d["x"], d["s"], d["y"], d["cone_types"], d["cone_dims"])
    dx, _ = conic.forward(cache, d["dA"], d["db"], d["dc"], atol=1e-12, btol=1e-12, conlim=1e12)
    assert np.allclose(dx, d["dx"], atol=0.3, rtol=0.01)
