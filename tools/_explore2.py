import sys, os, numpy as np
sys.path.insert(0, ".")
import bench_data, diffopt_b200
from oracle import qp as oqp
ctx = diffopt_b200.Context(0)
qpm = diffopt_b200.submodule("qp")
def rel_err(a, b): return np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-300)
for na in (50, 55, 57, 63):
    d = bench_data.qp_batch(12, n_active=48, seed0=4100 + na)
    for b in range(12):
        idle = np.flatnonzero(d["lam"][b] == 0)[:na - 48]
        d["lam"][b, idle] = 0.7
    of, orv = oqp.batch_forward_reverse(d["Q"], d["G"], d["A"], d["h"], d["z"], d["lam"], d["nu"], d["seed"], d["dQ"], d["dq"], d["dG"], d["dh"], d["dA"], d["db"])
    conds = [np.linalg.cond(oqp.create_lhs(d["z"][b], d["lam"][b], d["Q"][b], d["G"][b], d["h"][b], d["A"][b])) for b in range(12)]
    for kern in ("ldl", "lu", "generic"):
        os.environ["DIFFOPT_B200_QP_KERNEL"] = kern
        for _ in range(2):
            fwd, rev, info = qpm.solve_batch(ctx, d["Q"], d["G"], d["A"], d["h"], d["z"], d["lam"], d["nu"], fwd_dir=(d["dQ"], d["dq"], d["dG"], d["dh"], d["dA"], d["db"]), seed=d["seed"])
        print(na, kern, "stats", ctx.qp_last_stats(), "fwd", rel_err(fwd, of).max(), "rev", rel_err(rev, orv).max(), "cond max %.2e" % max(conds), flush=True)
