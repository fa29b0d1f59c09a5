"""One driver for the ncu captures under profiles/ (run from the repo root on a GPU box):
    python tools/prof_driver.py qp [B] [n m p active] | sparse [T] [nrhs] | lsqr_small [iters] | lsqr_big [scale] [iters] [window] | psd |
                                conic_batch [B] [iters]
Each case runs the library call a couple of times (first call warms buffers) and prints the library's own device time."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data  # noqa: E402
import diffopt_b200  # noqa: E402

ctx = diffopt_b200.Context(0)
case = sys.argv[1] if len(sys.argv) > 1 else "qp"
arg = lambda i, default: type(default)(sys.argv[i]) if len(sys.argv) > i else default


def conic_model(d):
    cm = diffopt_b200.submodule("conic")
    model = cm.ConicModel(ctx, d["A"], d["b"], d["c"], d["cone_types"], d["cone_dims"])
    model.set_variable_primal(d["x"]); model.set_constraint_primal(d["s"]); model.set_constraint_dual(d["y"])
    return model


if case == "qp":
    import torch
    capi = diffopt_b200.submodule("_capi")
    B = arg(2, 4096)
    n, m, p = arg(3, 64), arg(4, 64), arg(5, 16)   # other shapes run the shape-generic LDL' kernel
    d = bench_data.qp_batch_fast(B, n, m, p, n_active=arg(6, int(os.environ.get("ACTIVE", 16))))
    keys = ["Q", "G", "A", "h", "z", "lam", "nu", "dQ", "dq", "dG", "dh", "dA", "db", "seed"]
    mats = {"Q", "G", "A", "dQ", "dG", "dA"}
    dev = {k: torch.from_numpy(np.ascontiguousarray(d[k].transpose(0, 2, 1) if k in mats else d[k])).cuda() for k in keys}
    fo = torch.empty((B, n + m + p), dtype=torch.float64, device="cuda"); ro = torch.empty_like(fo)
    io = torch.zeros(B, dtype=torch.int32, device="cuda")
    ms = []
    for rep in range(int(os.environ.get("REPS", 4))):
        rc = ctx.lib.diffopt_b200_qp_batch_solve(ctx.h, B, n, m, p, *[capi.vp(dev[k].data_ptr()) for k in keys],
                                                 capi.vp(fo.data_ptr()), capi.vp(ro.data_ptr()), capi.vp(io.data_ptr()), capi.DEVICE)
        assert rc == 0
        ms.append(ctx.last_kernel_ms)
    print("device-resident kernel ms: min %.4f" % min(ms[1:]), "stats", ctx.qp_last_stats())
elif case == "sparse":
    lsq = diffopt_b200.submodule("lsqr")
    d = bench_data.mpc_config3(T=arg(2, 10_000))
    K = d["K"]
    R = np.asfortranarray(np.random.default_rng(2).standard_normal((K.shape[0], arg(3, 256))))
    F = lsq.SparseFactorization(ctx, K, trans=True)
    for _ in range(2):
        X = F.solve(R)
    print("N", K.shape[0], "factor ms", F.factor_ms, "solve ms", F.solve_ms, "residual", np.abs(K.T @ X[:, :4] - R[:, :4]).max())
elif case in ("lsqr_small", "lsqr_big"):
    if case == "lsqr_small":
        iters, d = arg(2, 200), bench_data.conic_config4()
    else:
        scale, iters = arg(2, 200), arg(3, 20)
        win = int(sys.argv[4]) if len(sys.argv) > 4 else None
        d = bench_data.conic_config4(n=5000 * scale, n_zero=500 * scale, n_nonneg=4000 * scale, n_soc=300 * scale, col_window=win)
    model = conic_model(d)
    model.tolerances = dict(atol=0.0, btol=0.0, conlim=0.0, maxiter=iters)
    for _ in range(2):
        model.reverse_differentiate(d["seed"])
    print("ms", model.last_stats["kernel_ms"], "us/iter", 1e3 * model.last_stats["kernel_ms"] / iters)
elif case == "psd":
    d = bench_data.maxcut_config5()
    model = conic_model(d)
    model.vp(); model.gradient_cache = False; model.vp()
    print("setup ms", model.setup_ms)
    t = np.random.default_rng(0).normal(size=len(d["b"]))
    for _ in range(3):
        model.dpi_apply(t)
    print("apply ms", ctx.last_kernel_ms)
elif case == "conic_batch":
    cm = diffopt_b200.submodule("conic")
    B, iters = arg(2, 148), arg(3, 20)
    models, seeds = [], []
    for k in range(B):
        d = bench_data.conic_config4(seed=4000 + k)
        models.append(conic_model(d)); seeds.append(d["seed"])
    batch = cm.ConicBatch(ctx, models)
    batch.tolerances = dict(atol=0.0, btol=0.0, conlim=0.0, maxiter=iters)
    for _ in range(2):
        batch.reverse_differentiate(np.stack(seeds))
    print("B", B, "iters", iters, "ms", batch.kernel_ms)
else:
    raise SystemExit(__doc__)
