import sys, numpy as np
sys.path.insert(0, ".")
import diffopt_b200, bench_data
from oracle import conic as oconic
ctx = diffopt_b200.Context(0)
cm = diffopt_b200.submodule("conic")
d = bench_data.conic_config4()
model = cm.ConicModel(ctx, d["A"], d["b"], d["c"], d["cone_types"], d["cone_dims"])
model.set_variable_primal(d["x"]); model.set_constraint_primal(d["s"]); model.set_constraint_dual(d["y"])
cache = oconic.gradient_cache(d["A"], d["b"], d["c"], d["x"], d["s"], d["y"], d["cone_types"], d["cone_dims"])
rel = lambda a, b: np.linalg.norm(a-b)/np.linalg.norm(b)
print("vp", rel(model.vp(), cache.vp))
T = np.random.default_rng(0).normal(size=12501)
a = model.M_apply(T); b = cache.M @ T
print("M", rel(a, b), "blocks", rel(a[:5000], b[:5000]), rel(a[5000:12500], b[5000:12500]), a[-1], b[-1])
a = model.M_apply(T, transpose=True); b = cache.M.T @ T
print("Mt", rel(a, b), "blocks", rel(a[:5000], b[:5000]), rel(a[5000:12500], b[5000:12500]), a[-1], b[-1])
for it in (1, 2, 3, 10, 50):
    tol = dict(atol=0.0, btol=0.0, conlim=0.0, maxiter=it)
    model.tolerances = tol
    model.reverse_differentiate(d["seed"])
    g = oconic.reverse(cache, d["seed"], **tol)
    print(it, rel(model.back_grad_cache["g"], g), model.last_stats)
