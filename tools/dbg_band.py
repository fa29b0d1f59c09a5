"""Config 3 at reduced horizon for ncu: one factorisation + one 64-column solve."""
import sys, numpy as np
sys.path.insert(0, ".")
import diffopt_b200, bench_data
ctx = diffopt_b200.Context(0)
lsq = diffopt_b200.submodule("lsqr")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
d = bench_data.mpc_config3(T=T)
K = d["K"]
R = np.random.default_rng(2).standard_normal((K.shape[0], 64))
F = lsq.SparseFactorization(ctx, K, trans=True)
X = F.solve(R)
print("N", K.shape[0], "bandwidth", F.bandwidth, "residual", np.abs(K.T @ X - R).max())
