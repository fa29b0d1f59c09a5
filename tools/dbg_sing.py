import sys, os, numpy as np
sys.path.insert(0, ".")
import diffopt_b200, bench_data
ctx = diffopt_b200.Context(0)
qpm = diffopt_b200.submodule("qp")
d = bench_data.qp_batch(8, seed0=31)
act = np.flatnonzero(d["lam"][3] > 0)
d["G"][3, act[1]] = d["G"][3, act[0]]
d["h"][3, act[1]] = d["h"][3, act[0]]
fwd, rev, info = qpm.solve_batch(ctx, d["Q"], d["G"], d["A"], d["h"], d["z"], d["lam"], d["nu"],
                       fwd_dir=(d["dQ"], d["dq"], d["dG"], d["dh"], d["dA"], d["db"]), seed=d["seed"])
print(os.environ.get("DIFFOPT_B200_QP_KERNEL"), "info", info, "max|rev[3]|", np.abs(rev[3]).max(), "max|fwd[3]|", np.abs(fwd[3]).max())
