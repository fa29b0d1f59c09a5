"""CPU restatement (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py) of the NonLinearProgram backend's factorisation
with inertia correction and of the parameter pull-back of src/parameters.jl.

Pinned by the reference's own tests: ``test/nlp_program.jl:767-795`` (a singular 5 x 5 KKT Jacobian that the correction
repairs) and ``test/parameters.jl:32-101, 317-445`` (reverse-mode parameter sensitivities with closed-form answers);
both are transcribed in tests/test_oracle_kat.py."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _lu(J):
    """``lu(J; check = false)``: (factorisation or None, status) with status 1 = singular (UMFPACK's warning; SuperLU
    raises "Factor is exactly singular" on the same zero pivot)."""
    try:
        return spla.splu(sp.csc_matrix(J)), 0
    except RuntimeError:
        return None, 1


def inertia_correction(M, num_cons, num_w, st=1e-6, max_corrections=50):
    """``_inertia_correction`` (NonLinearProgram.jl:356-390), argument order as the reference's.  Returns (K, corrections)."""
    n = M.shape[0]
    d = np.ones(n)
    d[num_w:num_w + num_cons] = -1.0
    D = sp.diags(d)
    J = sp.csc_matrix(M) + st * D
    K, status = _lu(J)
    num_c = 1
    while status == 1 and num_c < max_corrections:
        J = J + st * D
        K, status = _lu(J)
        num_c += 1
    return (K if status == 0 else None), num_c


def lu_with_inertia_correction(M, num_w, num_cons, st=1e-6, max_corrections=50):
    """``_lu_with_inertia_correction`` (:402-435).  Returns (K or None, corrections)."""
    K, status = _lu(M)
    if status == 1:
        return inertia_correction(M, num_cons, num_w, st=st, max_corrections=max_corrections)
    return K, 0


def compute_sensitivity(M, N, num_w, num_cons, st=1e-6, max_corrections=50):
    """nlp_utilities.jl:436-447: ``ds = -(K \\ N)``; zeros when K is nothing."""
    K, _ = lu_with_inertia_correction(M, num_w, num_cons, st, max_corrections)
    Nd = np.asarray(N.todense() if hasattr(N, "todense") else N, dtype=float)
    if K is None:
        return np.zeros((M.shape[0], Nd.shape[1]))
    return -K.solve(Nd)


def reverse_parameters(nparams, constraints, objective, param_values):
    """``reverse_differentiate!(::POI.Optimizer)`` (src/parameters.jl:341-534) in array form.

    constraints: list of dicts, one per parametric constraint, with
        ``grad_cte``  constant of ReverseConstraintFunction of that row,
        ``grad_coef`` mapping variable -> coefficient of ReverseConstraintFunction,
        ``p``  [(param, coefficient)]                affine parameter terms          (:349-360, :404-409)
        ``pp`` [(param1, param2, coefficient)]       parameter x parameter terms     (:410-430)
        ``pv`` [(param, variable, coefficient)]      parameter x variable terms      (:431-437)
    objective: the same keys for the parametric objective (grad of ReverseObjectiveFunction; its constant is 0), with
        the objective's division by 2 for a squared parameter (:497-512).
    Parameters are 0-based here."""
    out = np.zeros(nparams)

    def visit(c, is_objective):
        g = c.get("grad_cte", 0.0)
        for (p, coef) in c.get("p", []):
            out[p] += coef * g
        for (p1, p2, coef) in c.get("pp", []):
            div = 2.0 if (is_objective and p1 == p2) else 1.0
            v1, v2 = out[p1], out[p2]           # both read before either is written, as the reference does
            out[p1] = v1 + coef * g * param_values[p2] / div
            out[p2] = v2 + coef * g * param_values[p1] / div
        for (p, v, coef) in c.get("pv", []):
            out[p] += coef * c["grad_coef"].get(v, 0.0)

    for c in constraints:
        visit(c, False)
    if objective is not None:
        visit(objective, True)
    return out
