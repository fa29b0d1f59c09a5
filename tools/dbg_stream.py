import sys, os, numpy as np
sys.path.insert(0, ".")
import diffopt_b200, bench_data
from oracle import conic as oconic
ctx = diffopt_b200.Context(0)
cm = diffopt_b200.submodule("conic")
d = bench_data.conic_config4(n=600, n_zero=60, n_nonneg=400, n_soc=40, soc_dim=7, nnz_per_row=6, seed=11)
cache = oconic.gradient_cache(d["A"], d["b"], d["c"], d["x"], d["s"], d["y"], d["cone_types"], d["cone_dims"])
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
for iters in (1, 2, 3, 4, 5, 7, 10, 40):
    g = {}
    for mode in ("persistent", "stream"):
        os.environ["DIFFOPT_B200_LSQR"] = mode
        model = cm.ConicModel(ctx, d["A"], d["b"], d["c"], d["cone_types"], d["cone_dims"])
        model.set_variable_primal(d["x"]); model.set_constraint_primal(d["s"]); model.set_constraint_dual(d["y"])
        model.tolerances = dict(atol=0.0, btol=0.0, conlim=0.0, maxiter=iters)
        model.reverse_differentiate(d["seed"])
        g[mode] = (model.back_grad_cache["g"].copy(), model.last_stats)
    o = oconic.reverse(cache, d["seed"], atol=0.0, btol=0.0, conlim=0.0, maxiter=iters)
    print(iters, "pers-vs-oracle %.2e stream-vs-oracle %.2e pers-vs-stream %.2e" % (rel(g["persistent"][0], o), rel(g["stream"][0], o), rel(g["persistent"][0], g["stream"][0])),
          g["persistent"][1]["rnorm"], g["stream"][1]["rnorm"], g["stream"][1]["itn"])
