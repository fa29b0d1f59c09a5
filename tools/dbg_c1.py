import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, "/root/repo")
import diffopt_b200, bench_data
from oracle import qp as oqp, lsqr as olsqr
ctx = diffopt_b200.Context(0)
lsqr = diffopt_b200.submodule("lsqr")
d = bench_data.lp_config1()
K = sp.csc_matrix(oqp.create_lhs(d["z"], d["lam"], d["Q"], d["G"], d["h"], d["A"]))
rhs = np.zeros(300); rhs[:200] = d["seed"]
xo, info = olsqr.lsqr(K, rhs, return_info=True)
print("oracle itn", info.itn, "istop", info.istop)
for k in range(3):
    x, st = lsqr.lsqr_csc(ctx, K, rhs)
    print("gpu itn", st["itn"], "istop", st["istop"], "ms", ctx.last_kernel_ms, "rel", np.linalg.norm(x - xo) / np.linalg.norm(xo))
