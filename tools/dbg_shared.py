import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench_data, diffopt_b200
ctx = diffopt_b200.Context(0)
qpm = diffopt_b200.submodule("qp")
n, m, p, na = 64, 64, 16, 16
d = bench_data.qp_batch(20, n, m, p, n_active=na, seed0=8800 + n, shared=True)
fd = (d["dQ"], d["dq"], d["dG"], d["dh"], d["dA"], d["db"])
outs = []
for k in range(4):
    f0, r0, i0 = qpm.solve_batch(ctx, d["Q"], d["G"], d["A"], d["h"], d["z"], d["lam"], d["nu"], fwd_dir=fd, seed=d["seed"])
    print("plain", k, ctx.qp_last_stats(), np.abs(f0[:, 64:128][d["lam"] == 0]).max())
    outs.append((f0, r0))
for k in range(3):
    f1, r1, i1 = qpm.solve_batch_ex(ctx, d["Q"][0], d["G"][0], d["A"][0], d["h"], d["z"], d["lam"], d["nu"], fwd_dir=fd,
                                    seed=d["seed"], shared_matrices=True)
    print("shared", k, ctx.qp_last_stats(), np.abs(f1[:, 64:128][d["lam"] == 0]).max(), np.array_equal(f1, outs[-1][0]), np.array_equal(r1, outs[-1][1]),
          np.abs(f1 - outs[-1][0]).max(), np.abs(r1 - outs[-1][1]).max())
for k in range(1, 4):
    print("plain", k, "vs plain 3:", np.array_equal(outs[k][0], outs[3][0]), np.array_equal(outs[k][1], outs[3][1]))
bad = np.argwhere(f1 != outs[-1][0])
print("fwd diff positions (first 10):", bad[:10].tolist(), "count", len(bad))
bad = np.argwhere(r1 != outs[-1][1])
print("rev diff positions (first 10):", bad[:10].tolist(), "count", len(bad))
