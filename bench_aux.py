#!/usr/bin/env python
"""bench_aux.py -- secondary measurements of the hot path (BASELINE.json configs 1, 4, 5).

The driver's contract line comes from bench.py (config 2).  This script reports, one JSON line per config:
  config1  LP n=200, m=100: reverse mode through the LSQR-on-KKT branch (QuadraticProgram.jl:333-335, :488)
  config4  conic n=5000, m=7500 (zeros + nonneg + 300 x SOC(10)): reverse mode, LSQR on the matrix-free M
  config4x the same generator scaled until A no longer fits L2 (the HBM-meaningful form of config 4)
  config5  max-cut SDP, 200 x 200 PSD cone: eigendecomposition (setup), one Dpi apply, fixed-iteration reverse solve
Device time is the library's own CUDA-event bracket around its kernels (ctx.last_kernel_ms).  HBM roofline uses the
SURVEY 8(d) algorithmic bytes per LSQR iteration: 2 (12 nnz(M) + 4 (N+1)) + 88 N with nnz(M) as the reference builds
it.  The CPU leg times the oracle port (scipy LSQR on the explicit M) for a bounded number of iterations."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def lsqr_bytes_per_iter(nnzM, N):
    return 2 * (12 * nnzM + 4 * (N + 1)) + 88 * N


def run_conic(ctx, name, d, iters, cpu_iters, emit=True):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import diffopt_b200
    from oracle import conic as oconic
    cm = diffopt_b200.submodule("conic")
    model = cm.ConicModel(ctx, d["A"], d["b"], d["c"], d["cone_types"], d["cone_dims"])
    model.set_variable_primal(d["x"]); model.set_constraint_primal(d["s"]); model.set_constraint_dual(d["y"])
    if iters is None:      # the reference's defaults, to convergence
        sq = float(np.sqrt(np.finfo(float).eps))
        model.tolerances = dict(atol=sq, btol=sq, conlim=1 / sq, maxiter=None)
    else:
        model.tolerances = dict(atol=0.0, btol=0.0, conlim=0.0, maxiter=iters)
    model.reverse_differentiate(d["seed"])            # warm-up (includes setup)
    if iters is None:
        iters = model.last_stats["itn"]
    setup_ms = model.setup_ms
    ms = []
    for _ in range(3):
        model.reverse_differentiate(d["seed"])
        ms.append(model.last_stats["kernel_ms"])
    ms = min(ms)
    n, m = model.n, model.m
    N = n + m + 1
    nnzA = d["A"].nnz
    nnzM = 2 * nnzA + 2 * n + 3 * m   # as the reference assembles it: A'Dpi, -A, I - Dpi, c, b, -c', -b'Dpi (diagonal-ish Dpi)
    by = lsqr_bytes_per_iter(nnzM, N)
    line = {"config": name, "n": n, "m": m, "N": N, "nnz_A": int(nnzA), "lsqr_iterations": iters,
            "device_ms": ms, "us_per_iteration": 1e3 * ms / iters, "setup_ms": setup_ms,
            "roofline": {"bound": "hbm", "achieved": by * iters / (ms * 1e-3) / 1e9, "peak": hbm_peak(), "unit": "GB/s",
                         "algorithmic_bytes_per_iteration": by},
            "istop": model.last_stats["istop"], "rnorm": model.last_stats["rnorm"]}
    line["roofline"]["frac"] = line["roofline"]["achieved"] / line["roofline"]["peak"]
    line["working_set_MB"] = (12 * 2 * nnzA + 8 * 8 * N) / 1e6
    if cpu_iters is None:
        cache = oconic.gradient_cache(d["A"], d["b"], d["c"], d["x"], d["s"], d["y"], d["cone_types"], d["cone_dims"])
        dz = np.concatenate([d["seed"], np.zeros(m), [-(d["x"] @ d["seed"])]])
        t0 = time.perf_counter()
        ref = spla.lsqr(sp.csr_matrix(cache.M), dz, atol=model.tolerances["atol"], btol=model.tolerances["btol"],
                        conlim=model.tolerances["conlim"], iter_lim=N)
        dt = time.perf_counter() - t0
        g = model.back_grad_cache["g"]
        line["converged"] = {"gpu_iterations": iters, "cpu_iterations": int(ref[2]), "gpu_istop": model.last_stats["istop"],
                             "cpu_istop": int(ref[1]), "rel_err_vs_cpu_lsqr": float(np.linalg.norm(g - ref[0]) / np.linalg.norm(ref[0]))}
        line["cpu_baseline"] = {"ms": 1e3 * dt, "us_per_iteration": 1e6 * dt / max(int(ref[2]), 1), "kind": "port", "cores": 1,
                                "sample": "scipy.sparse.linalg.lsqr on the explicit M to convergence at the same tolerances"}
    elif cpu_iters:
        cache = oconic.gradient_cache(d["A"], d["b"], d["c"], d["x"], d["s"], d["y"], d["cone_types"], d["cone_dims"]) \
            if hasattr(oconic, "gradient_cache") else None
        if cache is not None:
            dz = np.concatenate([d["seed"], np.zeros(m), [-(d["x"] @ d["seed"])]])
            M = sp.csr_matrix(cache.M)
            t0 = time.perf_counter()
            spla.lsqr(M, dz, atol=0.0, btol=0.0, conlim=0.0, iter_lim=cpu_iters)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"us_per_iteration": 1e6 * dt / cpu_iters, "kind": "port", "cores": 1,
                                    "sample": f"scipy.sparse.linalg.lsqr on the explicit M (reference construction), {cpu_iters} iterations"}
    if emit:
        print(json.dumps(line), flush=True)
    return line


def run_conic_batch(ctx, B=512, iters=100, ctas=1, emit=True):
    """SURVEY 8(d) config 4 in its HBM-meaningful form: a lock-step batch of B independent config-4-sized problems (own
    sparsity, solution and seed each) advanced by one persistent kernel.  Fixed iteration count so the algorithmic bytes
    are exact; parity of the batch against single-problem solves is tests/test_conic_gpu.py's business."""
    import bench_data
    import diffopt_b200
    cm = diffopt_b200.submodule("conic")
    t0 = time.perf_counter()
    models, seeds, nnzA = [], [], 0
    for k in range(B):
        d = bench_data.conic_config4(seed=4000 + k)
        mdl = cm.ConicModel(ctx, d["A"], d["b"], d["c"], d["cone_types"], d["cone_dims"])
        mdl.set_variable_primal(d["x"]); mdl.set_constraint_primal(d["s"]); mdl.set_constraint_dual(d["y"])
        models.append(mdl); seeds.append(d["seed"]); nnzA += d["A"].nnz
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    batch = cm.ConicBatch(ctx, models, ctas_per_problem=ctas)
    setup_s = time.perf_counter() - t0
    batch.tolerances = dict(atol=0.0, btol=0.0, conlim=0.0, maxiter=iters)
    seeds = np.stack(seeds)
    out = batch.reverse_differentiate(seeds)
    ms = []
    for _ in range(3):
        out = batch.reverse_differentiate(seeds)
        ms.append(batch.kernel_ms)
    ms = min(ms)
    n, m = batch.n, batch.m
    N = n + m + 1
    nnzM = 2 * nnzA + B * (2 * n + 3 * m)
    by = 2 * (12 * nnzM + 4 * B * (N + 1)) + 88 * N * B
    # spot parity: problem 0 through the single-problem path at the same iteration count
    mdl = models[0]
    mdl.tolerances = batch.tolerances
    mdl.reverse_differentiate(seeds[0])
    g1 = mdl.back_grad_cache["g"]
    line = {"config": f"4b: lock-step batch of {B} independent config-4 problems (n=5000, m=7500 each), reverse, {iters} LSQR "
                      f"iterations each, one persistent kernel, {ctas} CTA(s) per problem",
            "B": B, "lsqr_iterations": iters, "device_ms": ms, "us_per_lockstep_iteration": 1e3 * ms / iters,
            "problem_iterations_per_s": B * iters / (ms * 1e-3), "single_problem_us_per_iteration": 1e3 * mdl.last_stats["kernel_ms"] / iters,
            "roofline": {"bound": "hbm", "achieved": by * iters / (ms * 1e-3) / 1e9, "peak": hbm_peak(), "unit": "GB/s",
                         "algorithmic_bytes_per_lockstep_iteration": by},
            "working_set_MB": (12 * 2 * nnzA + 8 * 8 * N * B) / 1e6,
            "rel_diff_problem0_vs_single_problem_path": float(np.linalg.norm(out["g"][0] - g1) / np.linalg.norm(g1)),
            "host_generate_s": gen_s, "batch_setup_s": setup_s}
    line["roofline"]["frac"] = line["roofline"]["achieved"] / line["roofline"]["peak"]
    if emit:
        print(json.dumps(line), flush=True)
    return line


def config3(ctx, portfolio=False, nrhs=256, cpu=True):
    """BASELINE config 3: ONE sparse KKT system (MPC QP, N = 240 000; or the portfolio arrowhead, N = 200 201), 256
    forward directions against one factorisation.  Device times are the library's own CUDA-event brackets
    (factorisation: all level launches; solve: both sweeps, right-hand sides resident in HBM)."""
    import scipy.sparse.linalg as spla
    import torch
    import diffopt_b200
    import bench_data
    lsq = diffopt_b200.submodule("lsqr")
    capi = diffopt_b200.submodule("_capi")
    if portfolio:
        d = bench_data.portfolio_config3()
        name = f"3p: sparse portfolio QP (n={d['n']}, m={d['m']}, p={d['p']}): arrowhead KKT"
    else:
        T = int(os.environ.get("DIFFOPT_AUX_MPC_T", 10_000))
        d = bench_data.mpc_config3(T=T)
        name = f"3: sparse MPC QP T={T} (n={d['n']}, m={d['m']}, p={d['p']})"
    K = d["K"]
    N = K.shape[0]
    rng = np.random.default_rng(33)
    R = np.zeros((N, nrhs), order="F")
    R[rng.integers(0, N, size=8 * nrhs), np.repeat(np.arange(nrhs), 8)] = rng.standard_normal(8 * nrhs)   # sparse directions
    t0 = time.perf_counter()
    F = lsq.SparseFactorization(ctx, K, trans=True)           # forward mode solves with LHS' (:438)
    setup_wall_ms = 1e3 * (time.perf_counter() - t0)
    F2 = lsq.SparseFactorization(ctx, K, trans=True)          # second run: buffers allocated, kernels loaded
    factor_ms = F2.factor_ms
    dev = torch.device("cuda", ctx.device)
    Rd = torch.from_numpy(np.ascontiguousarray(R.T)).to(dev)  # (nrhs, N) row-major == N x nrhs column-major
    Xd = torch.empty_like(Rd)
    ms = []
    for _ in range(5):
        ctx.check(ctx.lib.diffopt_b200_sparse_solve(ctx.h, nrhs, capi.vp(Rd.data_ptr()), capi.vp(Xd.data_ptr()), capi.DEVICE))
        ms.append(ctx.last_kernel_ms)
    solve_ms = min(ms[1:])
    X = np.asfortranarray(Xd.cpu().numpy().T)
    res = float((np.linalg.norm(K.T @ X[:, :8] - R[:, :8], axis=0) / np.linalg.norm(R[:, :8], axis=0)).max())
    st = F2.stats
    solve_bytes = 2 * 12 * st["nnz_lu"] + 2 * N * nrhs * 8      # (nnz L + nnz U) * 12 per sweep pair + 2 N nrhs 8
    line = {"config": name + f", N={N}, nnz={K.nnz}, {nrhs} forward directions, one factorisation",
            "method": st["method"], "fronts": st["fronts"], "tree_levels": st["levels"], "largest_front": st["max_front"],
            "nnz_LU_stored": st["nnz_lu"], "factor_flop": st["factor_flops"], "delayed_pivot_repeats": st["delayed_pivot_retries"],
            "analysis_host_ms": st["analysis_ms"], "factor_device_ms": factor_ms, "solve_device_ms": solve_ms,
            "total_device_ms": factor_ms + solve_ms, "first_setup_wall_ms_incl_host_analysis_and_h2d": setup_wall_ms,
            "factor_tflops": st["factor_flops"] / (factor_ms * 1e-3) / 1e12,
            "max_rel_residual_first_8_columns": res,
            "roofline": {"bound": "hbm", "kernel": "multi-RHS solve (both sweeps)", "achieved": solve_bytes / (solve_ms * 1e-3) / 1e9,
                         "peak": hbm_peak(), "unit": "GB/s", "algorithmic_bytes": solve_bytes,
                         "note": "SURVEY 8(d) bytes: factors once per sweep pair + the N x nrhs block read and written once; "
                                 "the implementation moves the block four times (b -> y -> x), so 0.5 is its ceiling"}}
    line["roofline"]["frac"] = line["roofline"]["achieved"] / line["roofline"]["peak"]
    if cpu:
        # CPU: factor once + 256 columns; reference-faithful (refactorise per direction) extrapolated from 4 directions
        Kt = K.T.tocsc()
        t0 = time.perf_counter(); lu = spla.splu(Kt); f_ms = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter(); Xc = lu.solve(np.ascontiguousarray(R[:, :32])); s_ms = 1e3 * (time.perf_counter() - t0) * (nrhs / 32)
        t0 = time.perf_counter()
        for k in range(2):
            spla.splu(Kt).solve(R[:, k])
        faithful_ms = 1e3 * (time.perf_counter() - t0) / 2 * nrhs
        line["rel_err_vs_superlu_32_columns"] = float((np.linalg.norm(X[:, :32] - Xc, axis=0) / np.linalg.norm(Xc, axis=0)).max())
        line["cpu_baseline"] = {"factor_ms": f_ms, "solve_256_ms_extrapolated_from_32": s_ms, "total_ms": f_ms + s_ms,
                                "reference_faithful_ms_extrapolated_from_2": faithful_ms, "kind": "port", "cores": 1,
                                "sample": "scipy splu (SuperLU stands in for UMFPACK); faithful = one factorisation per direction as QuadraticProgram.jl:438"}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2s,3,3p,4,4c,4x,5")
    args = ap.parse_args()
    import bench_data
    import diffopt_b200
    ctx = diffopt_b200.Context(0)
    todo = args.configs.split(",")
    if "1" in todo:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        from oracle import qp as oqp
        qpm = diffopt_b200.submodule("qp")
        d = bench_data.lp_config1()
        model = qpm.QPModel(ctx, d["Q"], d["q"], d["G"], d["h"], d["A"], d["b"])
        model.set_variable_primal(d["z"]); model.set_constraint_dual_le(-d["lam"]); model.set_constraint_dual_eq(-d["nu"])
        model.reverse_differentiate(d["seed"])
        ms = []
        for _ in range(3):
            model.reverse_differentiate(d["seed"])
            ms.append(ctx.last_kernel_ms)
        K = sp.csc_matrix(oqp.create_lhs(d["z"], d["lam"], d["Q"], d["G"], d["h"], d["A"]))
        rhs = np.zeros(300); rhs[:200] = d["seed"]
        t0 = time.perf_counter()
        ref = spla.lsqr(K, rhs, atol=1.49e-8, btol=1.49e-8, conlim=6.7e7, iter_lim=300)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        got = np.concatenate(model.back_grad_cache)
        want = np.concatenate(oqp.reverse(d["Q"], d["G"], d["h"], d["A"], d["z"], d["lam"], d["nu"], d["seed"]))
        st = getattr(model, "last_stats", None) or {}
        print(json.dumps({"config": "1: LP n=200 m=100, reverse via LSQR on the KKT matrix (default tolerances)",
                          "device_ms": min(ms), "lsqr_iterations": st.get("itn"), "rel_err_vs_oracle": float(
                              np.linalg.norm(got - want) / np.linalg.norm(want)),
                          "cpu_baseline": {"ms": cpu_ms, "iterations": int(ref[2]), "kind": "port", "cores": 1,
                                           "sample": "scipy.sparse.linalg.lsqr on the CSC KKT matrix"},
                          "note": "N=300: launch/grid-sync latency bound, no roofline claim"}), flush=True)
    if "3" in todo:
        print(json.dumps(config3(ctx)), flush=True)
    if "3nocpu" in todo:
        print(json.dumps(config3(ctx, cpu=False)), flush=True)
    if "3p" in todo:
        print(json.dumps(config3(ctx, portfolio=True)), flush=True)
    if "3pnocpu" in todo:
        print(json.dumps(config3(ctx, portfolio=True, cpu=False)), flush=True)
    if "4" in todo:
        run_conic(ctx, "4: conic n=5000 m=7500 (zeros 500 + nonneg 4000 + 300 x SOC(10)), reverse, matrix-free M",
                  bench_data.conic_config4(), iters=2000, cpu_iters=300)
    if "4c" in todo:
        d = bench_data.conic_config4_conditioned()
        run_conic(ctx, "4c: config 4 on the well-conditioned generator (75 % of the nonnegative rows active, solution scaled 0.02: "
                       "cond(M) ~ 1e4), reverse, the reference's default LSQR tolerances to convergence", d, iters=None, cpu_iters=None)
    for key in todo:
        if key.startswith("4b"):      # 4b, 4b:B, 4b:B:ctas
            parts = key.split(":")
            run_conic_batch(ctx, B=int(parts[1]) if len(parts) > 1 else 512, ctas=int(parts[2]) if len(parts) > 2 else 1)
    if "4x" in todo:
        run_conic(ctx, "4x: config-4 generator scaled 200x (n=1e6, m=1.5e6, nnz(A)=1.5e7): A exceeds L2",
                  bench_data.conic_config4(n=1_000_000, n_zero=100_000, n_nonneg=800_000, n_soc=60_000), iters=200, cpu_iters=0)
    if "4xb" in todo:
        run_conic(ctx, "4xb: scaled 200x like 4x, but stage-structured sparsity (row i touches variables within +-2000 of "
                       "i n/m, as in MPC / network / PDE-constrained programs): x gathers have locality",
                  bench_data.conic_config4(n=1_000_000, n_zero=100_000, n_nonneg=800_000, n_soc=60_000, col_window=2000),
                  iters=200, cpu_iters=0)
    if "5b" in todo:
        run_psd_batch(ctx)
    if "5" in todo:
        config5(ctx)
    if "2s" in todo:
        dense_single(ctx)


def dense_single(ctx, emit=True, sizes=((120, 120, 30), (400, 400, 100), (900, 900, 200))):
    """The stock `solve_system` plug point on ONE dense QP (the reference's `LHS \\ RHS`, QuadraticProgram.jl:486-492, through
    diffopt_b200_kkt_solve_csc): KKT matrices of order 270 / 900 / 2000 with dense Q, G, A, four right-hand sides -- the blocked
    LU over the whole GPU against SuperLU (the oracle's stand-in for UMFPACK) and dense LAPACK on the host."""
    import bench_data
    import diffopt_b200
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from oracle import qp as oqp
    lsq = diffopt_b200.submodule("lsqr")
    rows = []
    for n, m, p in sizes:
        d = bench_data.qp_batch(1, n, m, p, n_active=p, seed0=5)
        K = sp.csc_matrix(oqp.create_lhs(d["z"][0], d["lam"][0], d["Q"][0], d["G"][0], d["h"][0], d["A"][0]))
        N = K.shape[0]
        R = np.random.default_rng(0).standard_normal((N, 4))
        lsq.solve_csc(ctx, K, R)
        X = lsq.solve_csc(ctx, K, R)
        dev_ms = ctx.last_kernel_ms
        t0 = time.perf_counter(); lu = spl.splu(K); lu.solve(R); t1 = time.perf_counter(); np.linalg.solve(K.toarray(), R); t2 = time.perf_counter()
        rows.append({"N": N, "nnz": int(K.nnz), "device_ms": dev_ms, "max_residual": float(np.abs(K @ X - R).max()),
                     "cpu_superlu_ms": 1e3 * (t1 - t0), "cpu_dense_lapack_ms": 1e3 * (t2 - t1)})
    line = {"config": "2s: one dense QP through the solve_system drop-in (kkt_solve_csc), 4 right-hand sides", "systems": rows}
    if emit:
        print(json.dumps(line), flush=True)
    return line


def config5(ctx, emit=True):
    """BASELINE config 5: max-cut SDP with one 200 x 200 PSD cone -- eigendecomposition (setup), one Dpi apply, 50 LSQR iterations."""
    import diffopt_b200
    import scipy.sparse as sp
    from oracle import cones as ocones
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
    model.vp()
    model.gradient_cache = False   # time the second setup: buffers allocated, modules loaded
    model.vp()
    setup_ms = model.setup_ms
    t = rng.normal(size=dd + k)
    model.dpi_apply(t)
    ap_ms = []
    for _ in range(3):
        model.dpi_apply(t)
        ap_ms.append(ctx.last_kernel_ms)
    model.tolerances = dict(atol=0.0, btol=0.0, conlim=0.0, maxiter=50)
    model.reverse_differentiate(rng.normal(size=k))
    t0 = time.perf_counter()
    w, U = np.linalg.eigh(X - Smat)
    cpu_eig_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    ocones.Dpi_apply(y - s, [ocones.ZERO, ocones.PSD], [dd, k], t)
    cpu_apply_ms = 1e3 * (time.perf_counter() - t0)
    line = ({"config": "5: max-cut SDP, 200 x 200 PSD cone (20 100 triangle rows + 200 zero rows)",
                      "setup_ms_incl_eigendecomposition": setup_ms, "dpi_apply_ms": min(ap_ms),
                      "dpi_apply_gflops": 4 * 2 * dd ** 3 / (min(ap_ms) * 1e-3) / 1e9,
                      "reverse_50_lsqr_iterations_ms": model.last_stats["kernel_ms"],
                      "cpu_baseline": {"eigh_ms": cpu_eig_ms, "dpi_apply_ms_incl_eigh": cpu_apply_ms, "kind": "port",
                                       "sample": "numpy eigh + operator-form apply; the reference's dense 20100^2 Jacobian "
                                                 "(3.2 GB, ~1e12 flop) is not formed"}})
    if emit:
        print(json.dumps(line), flush=True)
    return line


def run_psd_batch(ctx):
    """Batched micro-config of SURVEY 8(d) config 5: 512 PSD cones of side 16 / 32, eigendecompositions in one launch."""
    import scipy.sparse as sp
    import diffopt_b200
    from oracle import cones as ocones
    cm = diffopt_b200.submodule("conic")
    rng = np.random.default_rng(55)
    for side in (16, 32):
        mats = []
        for _ in range(512):
            Xm = rng.normal(size=(side, side))
            mats.append((Xm + Xm.T) / 2)
        dims = [side * (side + 1) // 2] * 512
        k = sum(dims)
        y = np.concatenate([ocones.vec_symm(m) for m in mats])
        model = cm.ConicModel(ctx, (-sp.identity(k)).tocsc(), np.zeros(k), np.zeros(k), [ocones.PSD] * 512, dims)
        model.set_variable_primal(np.zeros(k)); model.set_constraint_primal(np.zeros(k)); model.set_constraint_dual(y)
        model.vp()
        model.gradient_cache = False   # second setup: buffers allocated, kernels warm
        model.vp()
        setup_ms = model.setup_ms
        t = rng.normal(size=k)
        model.dpi_apply(t)
        model.dpi_apply(t)
        ap = ctx.last_kernel_ms
        t0 = time.perf_counter()
        for m_ in mats:
            np.linalg.eigh(m_)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        print(json.dumps({"config": f"5b: 512 PSD cones of side {side} (batched eigendecomposition + B + pi)",
                          "setup_ms_incl_eigendecomposition": setup_ms, "eig_per_s": 512 / (setup_ms * 1e-3),
                          "dpi_apply_ms": ap,
                          "cpu_baseline": {"eigh_512_ms": cpu_ms, "kind": "port", "cores": 1,
                                           "sample": "numpy.linalg.eigh in a loop"}}), flush=True)


if __name__ == "__main__":
    main()
