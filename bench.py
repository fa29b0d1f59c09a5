#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 sensitivity hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config 2 of BASELINE.json): a batch of 4096 dense random QPs, n=64, m_eq=16,
m_ineq=64 (KKT order N=144).  One *solve* = KKT assembly + one LU factorisation + one forward
(LHS') and one reverse (LHS) sensitivity solve for one instance; one *step* = one pass over the
batch.  With N GPUs every rank owns its own 4096-instance batch (weak scaling, no data-path
collective; the whole-job value is the sum).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_VAR, M_INEQ, P_EQ = 64, 64, 16
KKT_N = N_VAR + M_INEQ + P_EQ
FLOP_PER_SOLVE = 2 * KKT_N**3 // 3 + 2 * (2 * KKT_N**2)          # 2 073 600
BYTES_PER_SOLVE = 8 * ((N_VAR**2 + M_INEQ * N_VAR + P_EQ * N_VAR + N_VAR + 2 * M_INEQ + P_EQ) * 2 - M_INEQ
                       + N_VAR + 2 * KKT_N)                        # ~153 KB: inputs + direction + seed + outputs
METRIC = "KKT sensitivity solves/sec (4096xQP n=64)"
FIELDS = ["Q", "G", "A", "h", "z", "lam", "nu", "dQ", "dq", "dG", "dh", "dA", "db", "seed"]
SHAPES = dict(Q=(N_VAR, N_VAR), G=(M_INEQ, N_VAR), A=(P_EQ, N_VAR), dQ=(N_VAR, N_VAR), dG=(M_INEQ, N_VAR),
              dA=(P_EQ, N_VAR))


def host_buffers(B, seed, pinned_alloc=None):
    """Instance-major, column-major-per-instance fp64 buffers exactly as the C ABI takes them."""
    import bench_data
    d = bench_data.qp_batch_fast(B, N_VAR, M_INEQ, P_EQ, seed=seed)
    out = {}
    for k in FIELDS:
        a = d[k]
        if k in SHAPES:
            a = a.transpose(0, 2, 1)
        if pinned_alloc is not None:
            buf = pinned_alloc(a.shape)
            np.copyto(buf, a)
            out[k] = buf
        else:
            out[k] = np.ascontiguousarray(a)
    return out, d


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=2.0):
        """Block until nvidia-smi has produced its first line (so that the timed region is actually sampled)."""
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.01)

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def fp64_peaks():
    """Runs tools/fp64_peak (cuBLAS DGEMM 8192^3 + raw DMMA/DFMA issue rates) on this GPU."""
    tool = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        out = subprocess.run([tool], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on host cores (oracle port; Julia/UMFPACK are not in the image)

def _cpu_chunk(args):
    """Reference-faithful per-instance work: sparse KKT (create_LHS_matrix) and one sparse LU
    factorisation PER differentiate call (QuadraticProgram.jl:335 and :438 both call `\\`)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from oracle import qp as oqp
    lo, hi = args
    d = _CPU_DATA
    n, m = N_VAR, M_INEQ
    for b in range(lo, hi):
        K = sp.csc_matrix(oqp.create_lhs(d["z"][b], d["lam"][b], d["Q"][b], d["G"][b], d["h"][b], d["A"][b]))
        rf = oqp.forward_rhs(d["z"][b], d["lam"][b], d["nu"][b], d["dQ"][b], d["dq"][b], d["dG"][b], d["dh"][b],
                             d["dA"][b], d["db"][b])
        rb = np.zeros(KKT_N)
        rb[:n] = d["seed"][b]
        spla.splu(K.T.tocsc()).solve(rf)      # forward_differentiate!: LHS' \ RHS
        spla.splu(K).solve(rb)                # reverse_differentiate!: LHS \ RHS
    return hi - lo


_CPU_DATA = None


def cpu_reference_rate(d, sample, procs):
    """solves/s of the CPU port on `sample` instances using `procs` worker processes (fork: the
    workers inherit the inputs, nothing is pickled inside the timed region)."""
    import multiprocessing as mp
    global _CPU_DATA
    _CPU_DATA = {k: v[:sample] for k, v in d.items()}
    bounds = np.linspace(0, sample, procs + 1).astype(int)
    chunks = [(int(bounds[i]), int(bounds[i + 1])) for i in range(procs) if bounds[i + 1] > bounds[i]]
    if procs == 1:
        _cpu_chunk((0, min(4, sample)))
        t0 = time.perf_counter()
        _cpu_chunk(chunks[0])
        return sample / (time.perf_counter() - t0)
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_cpu_chunk, [(0, min(4, sample))] * procs)  # warm the workers (imports)
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, chunks)
        dt = time.perf_counter() - t0
    return sample / dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    import bench_data
    procs = os.cpu_count() or 1
    sample = int(os.environ.get("DIFFOPT_BENCH_CPU_SAMPLE", 1024))
    d = bench_data.qp_batch_fast(sample, N_VAR, M_INEQ, P_EQ, seed=2026)
    for _ in range(args.warmup):
        cpu_reference_rate(d, min(sample, 64 * procs), procs)
    dt = 0.0
    done = 0
    for _ in range(args.steps):
        dt += sample / cpu_reference_rate(d, sample, procs)   # times the worker map only, not pool start-up
        done += sample
    val = done / dt
    sample_txt = (f"{sample} of 4096 instances per step; per instance: sparse KKT build + splu factorisation and "
                  f"solve for the forward call (LHS') and again for the reverse call (LHS), like "
                  f"QuadraticProgram.jl:335,438; SuperLU (scipy) stands in for UMFPACK; {procs} worker processes")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps * (4096 / sample),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "4096 x dense QP n=64 m_ineq=64 m_eq=16 (KKT N=144), forward + reverse "
                                   "sensitivities; CPU arm runs a bounded sample", "batch_per_gpu": 4096},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": procs, "kind": "port", "sample": sample_txt},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------

def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import diffopt_b200
    capi = diffopt_b200.submodule("_capi")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = diffopt_b200.Context(local_rank)
    lib = ctx.lib
    B = args.batch
    N = KKT_N

    hb, logical = host_buffers(B, seed=2026 + rank, pinned_alloc=capi.pinned_empty)
    db = {k: torch.from_numpy(v).to(dev) for k, v in hb.items()}     # resident inputs (torch = device memory)
    fwd = torch.empty((B, N), dtype=torch.float64, device=dev)
    rev = torch.empty((B, N), dtype=torch.float64, device=dev)
    info = torch.zeros(B, dtype=torch.int32, device=dev)
    fwd_h = capi.pinned_empty((B, N))
    rev_h = capi.pinned_empty((B, N))
    info_h = np.zeros(B, dtype=np.int32)
    in_bytes = sum(v.nbytes for v in hb.values())
    out_bytes = fwd_h.nbytes + rev_h.nbytes + info_h.nbytes

    def dptr(t):
        return capi.vp(t.data_ptr())

    dev_args = [dptr(db[k]) for k in FIELDS] + [dptr(fwd), dptr(rev), dptr(info)]

    def step_device():
        # stream-ordered form of the call (device-resident inputs): batches run back to back on the ctx stream, the
        # status of the last one is collected by diffopt_b200_synchronize after the timed region
        rc = lib.diffopt_b200_qp_batch_solve_async(ctx.h, B, N_VAR, M_INEQ, P_EQ, *dev_args)
        if rc != 0:
            raise RuntimeError(f"qp_batch_solve_async rc={rc}: {lib.diffopt_b200_last_error(ctx.h).decode()}")

    def step_device_blocking():
        rc = lib.diffopt_b200_qp_batch_solve(ctx.h, B, N_VAR, M_INEQ, P_EQ, *dev_args, capi.DEVICE)
        if rc != 0:
            raise RuntimeError(f"qp_batch_solve rc={rc}: {lib.diffopt_b200_last_error(ctx.h).decode()}")

    def finish_device():
        rc = lib.diffopt_b200_synchronize(ctx.h)
        if rc != 0:
            raise RuntimeError(f"qp_batch_solve_async status {rc}: {lib.diffopt_b200_last_error(ctx.h).decode()}")

    def step_e2e():
        rc = lib.diffopt_b200_qp_batch_solve(
            ctx.h, B, N_VAR, M_INEQ, P_EQ, *[capi.ptr(hb[k]) for k in FIELDS], capi.ptr(fwd_h), capi.ptr(rev_h),
            capi.ptr(info_h), capi.HOST)
        if rc != 0:
            raise RuntimeError(f"qp_batch_solve(host) rc={rc}: {lib.diffopt_b200_last_error(ctx.h).decode()}")

    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """CUDA events on the library's own stream, barrier + synchronize on both sides."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = 0.0
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
            kms += ctx.last_kernel_ms
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, kms / steps

    sampler = ClockSampler(local_rank) if rank == 0 else None   # started early: nvidia-smi needs ~0.2 s to come up
    for _ in range(max(args.warmup, 3)):
        step_device()
    finish_device()
    if sampler:
        sampler.wait_first()
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    ms, _ = timed(step_device, args.steps)
    finish_device()
    t1 = time.perf_counter()
    launches = ctx.launch_count - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    kernel_ms = ms / args.steps   # device time per launch: the timed region holds nothing but the K launches
    # the blocking form of the same call (what the reference-facing API does): a few calls, for the record
    step_device_blocking()
    ms_blocking, call_kernel_ms = timed(step_device_blocking, 5)

    # parity spot-check of what was just timed (oracle as checker only)
    if rank == 0:
        from oracle import qp as oqp
        sl = slice(0, 8)
        of, orv = oqp.batch_forward_reverse(*[logical[k][sl] for k in ["Q", "G", "A", "h", "z", "lam", "nu", "seed",
                                                                       "dQ", "dq", "dG", "dh", "dA", "db"]])
        gf, gr = fwd[sl].cpu().numpy(), rev[sl].cpu().numpy()
        err = max((np.linalg.norm(gf - of, axis=1) / np.linalg.norm(of, axis=1)).max(),
                  (np.linalg.norm(gr - orv, axis=1) / np.linalg.norm(orv, axis=1)).max())
        if not err <= 1e-8:
            raise RuntimeError(f"parity check failed inside bench: rel err {err:.3e}")
    else:
        err = None

    # end-to-end through the C ABI with host (pinned) buffers: H2D + kernel + D2H every step
    for _ in range(2):
        step_e2e()
    e2e_steps = max(2, min(args.steps, 10))
    ms_e2e, _ = timed(step_e2e, e2e_steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * B * args.steps / (ms * 1e-3)
    e2e_val = world * B * e2e_steps / (ms_e2e * 1e-3)
    peaks = fp64_peaks() if world == 1 else {}
    achieved = B * FLOP_PER_SOLVE / (kernel_ms * 1e-3) / 1e12
    peak = peaks.get("dgemm8192_tflops_sustained")
    prof = _profile_facts()
    hbm_peak = _hbm_peak()
    hbm_ach = B * BYTES_PER_SOLVE / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if peak else None, "traffic": prof.get("dram_bytes_per_launch"),
                "kernel": "qp_kkt_sqd_kernel (KKT assembly + blocked LDL' on DMMA + 2 solves; pivoted-LU fallback kernel "
                          "launched behind it, plus the active-set scan), avg device time per launch %.3f ms over the "
                          "timed region; blocking call: %.3f ms per call (%.3f ms of it between the library's own "
                          "events)" % (kernel_ms, ms_blocking / 5, call_kernel_ms),
                "algorithmic_flop_per_solve": FLOP_PER_SOLVE, "algorithmic_bytes_per_solve": BYTES_PER_SOLVE,
                "executed_flop_per_solve_estimate": prof.get("executed_flop_per_solve"),
                "note": "achieved = SURVEY.md 8(d) algorithmic flop (LU of the N=144 KKT + 2x2 triangular solves) / device "
                        "time; the kernel reaches it by eliminating column singletons and factorising the symmetric "
                        "quasi-definite reduced system with LDL' (fewer executed flop, see DESIGN.md); traffic from "
                        "profiles/ (ncu capture of the same kernel, per launch)",
                "peak_source": "measured in this run: cuBLAS DGEMM 8192^3 sustained 4 s (tools/fp64_peak); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "fp64_peaks": peaks,
                "hbm_secondary": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": hbm_ach / hbm_peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)"}}
    line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "4096 x dense QP n=64 m_ineq=64 m_eq=16 (KKT N=144), forward + reverse "
                                   "sensitivities (BASELINE.json configs[1])",
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"instances sharded x{world}",
                       "l2_policy": "inputs (627 MB/step) larger than the 126 MB L2",
                       "call": "diffopt_b200_qp_batch_solve_async x K + diffopt_b200_synchronize (stream-ordered, "
                               "device-resident inputs); e2e uses the blocking host-buffer call"},
            "roofline": roofline, "e2e": {"value": e2e_val, "unit": "solves/s", "h2d_bytes_per_step": in_bytes,
                                          "d2h_bytes_per_step": out_bytes, "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": launches, "clocks": clocks, "parity_rel_err": err}
    if world == 1 and not args.no_cpu:
        procs = os.cpu_count() or 1
        sample = int(os.environ.get("DIFFOPT_BENCH_CPU_SAMPLE", 1024))
        sub = {k: logical[k][:sample] for k in logical}
        rate = cpu_reference_rate(sub, sample, procs)
        line["cpu_baseline"] = {
            "value": rate, "unit": "solves/s", "cores": procs, "kind": "port",
            "sample": f"first {sample} of the 4096 instances; per instance sparse KKT build + splu factorise+solve "
                      f"for forward (LHS') and again for reverse (LHS) as QuadraticProgram.jl:335,438 do; "
                      f"SuperLU stands in for UMFPACK; {procs} worker processes"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _profile_facts():
    """Numbers read off the committed ncu capture of the dominant kernel (profiles/qp_sqd_facts.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "qp_sqd_facts.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def _hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"]
    except Exception:
        return 6650.0  # B200_PROFILING.md fallback


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
