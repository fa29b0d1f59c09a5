"""Config 5 (200 x 200 PSD cone) driver for ncu: setup (eigensolver) twice, a few Dpi applies."""
import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, ".")
import diffopt_b200
from oracle import cones as ocones
ctx = diffopt_b200.Context(0)
cm = diffopt_b200.submodule("conic")
dd, r = 200, 20
rng = np.random.default_rng(5)
V = rng.normal(size=(dd, r)); V /= np.linalg.norm(V, axis=1, keepdims=True)
X = V @ V.T
Qf, _ = np.linalg.qr(np.hstack([V, rng.normal(size=(dd, dd - r))]))
W = Qf[:, r:]
Smat = (W * rng.uniform(0.5, 1.5, size=dd - r)) @ W.T
k = dd * (dd + 1) // 2
s = np.concatenate([np.zeros(dd), ocones.vec_symm(X)])
y = np.concatenate([rng.normal(size=dd), ocones.vec_symm(Smat)])
iu = [(i * (i + 1) // 2 + i) for i in range(dd)]
A = sp.vstack([sp.csc_matrix((np.ones(dd), (np.arange(dd), iu)), shape=(dd, k)), -sp.identity(k)]).tocsc()
x = ocones.vec_symm(X)
model = cm.ConicModel(ctx, A, A @ x + s, -(A.T @ y), [ocones.ZERO, ocones.PSD], [dd, k])
model.set_variable_primal(x); model.set_constraint_primal(s); model.set_constraint_dual(y)
model.vp(); model.gradient_cache = False; model.vp()
print("setup ms", model.setup_ms)
t = rng.normal(size=dd + k)
for _ in range(3):
    model.dpi_apply(t)
print("apply ms", ctx.last_kernel_ms)
