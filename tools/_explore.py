import sys, time, numpy as np
sys.path.insert(0, ".")
import bench_data, diffopt_b200
from oracle import conic as oconic, lsqr as olsqr
ctx = diffopt_b200.Context(0)
cm = diffopt_b200.submodule("conic")
rel = lambda a, b: np.linalg.norm(a-b)/np.linalg.norm(b)
def mk(d):
    model = cm.ConicModel(ctx, d["A"], d["b"], d["c"], d["cone_types"], d["cone_dims"])
    model.set_variable_primal(d["x"]); model.set_constraint_primal(d["s"]); model.set_constraint_dual(d["y"])
    return model
# config 4 conditioned
d = bench_data.conic_config4_conditioned()
model = mk(d)
cache = oconic.gradient_cache(d["A"], d["b"], d["c"], d["x"], d["s"], d["y"], d["cone_types"], d["cone_dims"])
for tol in [dict(atol=1.49e-8, btol=1.49e-8, conlim=6.7e7, maxiter=12501), dict(atol=1e-13, btol=1e-13, conlim=0.0, maxiter=12501)]:
    model.tolerances = tol
    model.reverse_differentiate(d["seed"]); model.reverse_differentiate(d["seed"])
    t0=time.time(); g, info = olsqr.lsqr(cache.M, np.concatenate([d["seed"], np.zeros(7500), [-(d["x"]@d["seed"])]]), return_info=True, **tol); dt=time.time()-t0
    print("c4cond", tol["atol"], "gpu", model.last_stats, "oracle", info.istop, info.itn, info.rnorm, "cpu s", dt, "rel", rel(model.back_grad_cache["g"], g), flush=True)
# config 5 variants
for r in (20, 8):
    d = bench_data.maxcut_config5(r=r)
    model = mk(d)
    for it in (1, 2, 5, 10, 20, 50, 200):
        tol = dict(atol=0.0, btol=0.0, conlim=0.0, maxiter=it)
        model.tolerances = tol
        model.reverse_differentiate(d["seed"])
        g = oconic.reverse_matrix_free(d["A"], d["b"], d["c"], d["x"], d["s"], d["y"], d["cone_types"], d["cone_dims"], d["seed"], **tol)
        print("c5 r", r, "it", it, "rel", rel(model.back_grad_cache["g"], g), model.last_stats["rnorm"], flush=True)
    tol = dict(atol=1.49e-8, btol=1.49e-8, conlim=6.7e7, maxiter=3000)
    model.tolerances = tol
    model.reverse_differentiate(d["seed"])
    t0=time.time(); g, info = olsqr.lsqr(oconic.matrix_free_ops(d["A"], d["b"], d["c"], d["x"], d["s"], d["y"], d["cone_types"], d["cone_dims"]), np.concatenate([d["seed"], np.zeros(d["A"].shape[0]), [-(d["x"]@d["seed"])]]), return_info=True, **tol)
    print("c5 r", r, "default tol: gpu", model.last_stats, "oracle", info.istop, info.itn, info.rnorm, "cpu s", time.time()-t0, "rel", rel(model.back_grad_cache["g"], g), flush=True)
