import sys, numpy as np, time
sys.path.insert(0, ".")
import diffopt_b200, bench_data
from oracle import qp as oqp
ctx = diffopt_b200.Context(0)
qpm = diffopt_b200.submodule("qp")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
d = bench_data.qp_batch_fast(B)
for rep in range(2):
    fwd, rev, info = qpm.solve_batch(ctx, d["Q"], d["G"], d["A"], d["h"], d["z"], d["lam"], d["nu"],
                                 fwd_dir=(d["dQ"], d["dq"], d["dG"], d["dh"], d["dA"], d["db"]), seed=d["seed"])
print("kernel ms", ctx.last_kernel_ms, "info any", info.any())
sl = slice(0, 32)
of, orv = oqp.batch_forward_reverse(*[d[k][sl] for k in ["Q","G","A","h","z","lam","nu","seed","dQ","dq","dG","dh","dA","db"]])
re = lambda a,b: (np.linalg.norm(a-b,axis=1)/np.linalg.norm(b,axis=1)).max()
print("fwd err", re(fwd[sl], of), "rev err", re(rev[sl], orv))
