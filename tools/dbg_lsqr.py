import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, ".")
import diffopt_b200
from oracle import lsqr as olsqr
ctx = diffopt_b200.Context(0)
lsqr = diffopt_b200.submodule("lsqr")
rng = np.random.default_rng(3)
A = (sp.random(400, 250, density=0.04, random_state=1) + sp.eye(400, 250)).tocsc()
b = rng.normal(size=400)
x, st = lsqr.lsqr_csc(ctx, A, b)
xo, info = olsqr.lsqr(A, b, return_info=True)
print("gpu", st)
print("cpu", info)
print("rel", np.linalg.norm(x-xo)/np.linalg.norm(xo))
for mi in (63, 64, 65):
    xo2, i2 = olsqr.lsqr(A, b, return_info=True, maxiter=mi)
    x2, s2 = lsqr.lsqr_csc(ctx, A, b, maxiter=mi)
    print(mi, "cpu", i2.istop, i2.itn, i2.rnorm, i2.arnorm, "gpu", s2, "rel", np.linalg.norm(x2-xo2)/np.linalg.norm(xo2))
