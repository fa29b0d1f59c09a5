import sys, numpy as np
sys.path.insert(0, ".")
import diffopt_b200, bench_data
ctx = diffopt_b200.Context(0)
cm = diffopt_b200.submodule("conic")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
d = bench_data.conic_config4()
model = cm.ConicModel(ctx, d["A"], d["b"], d["c"], d["cone_types"], d["cone_dims"])
model.set_variable_primal(d["x"]); model.set_constraint_primal(d["s"]); model.set_constraint_dual(d["y"])
model.tolerances = dict(atol=0.0, btol=0.0, conlim=0.0, maxiter=iters)
for _ in range(2):
    model.reverse_differentiate(d["seed"])
print("ms", model.last_stats["kernel_ms"], "us/iter", 1e3 * model.last_stats["kernel_ms"] / iters)
