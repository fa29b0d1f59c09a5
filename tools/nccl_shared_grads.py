"""torchrun --nproc-per-node N tools/nccl_shared_grads.py : shared-parameter reverse gradients over N GPUs.
Each rank solves its instance shard with the CUDA library (shared Q, G, A), sums its shard's parameter gradients on the
device and all-reduces them with the ctx-owned NCCL communicator (device resident, diffopt_b200_qp_batch_shared_grads
with DIFFOPT_QP_ALLREDUCE); rank 0 checks the total against the CPU oracle (checker only)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench_data
import diffopt_b200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sh = diffopt_b200.submodule("sharding")
qpm = diffopt_b200.submodule("qp")
ctx = diffopt_b200.Context(local)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d = bench_data.qp_batch(B, shared=True, seed0=17)
sh.nccl_init(ctx, rank, world)
rev, total = sh.sharded_reverse_shared_params_device(ctx, d["Q"][0], d["G"][0], d["A"][0], d["h"], d["z"], d["lam"], d["nu"],
                                                     d["seed"], rank, world)
lo, hi = sh.shard_range(B, rank, world)
ok = torch.ones(1, device=dev)
if rank == 0:
    from oracle import qp as oqp
    want = [0.0] * 6
    for b in range(B):
        dz, dl, dn = oqp.reverse(d["Q"][b], d["G"][b], d["h"][b], d["A"][b], d["z"][b], d["lam"][b], d["nu"][b], d["seed"][b])
        g = oqp.reverse_param_grads(d["z"][b], d["lam"][b], d["nu"][b], dz, dl, dn)
        want = [w + x for w, x in zip(want, g)]
        if lo <= b < hi:
            ref = np.concatenate([dz, dl, dn])
            assert np.linalg.norm(rev[b - lo] - ref) <= 1e-8 * np.linalg.norm(ref)
    err = max(np.abs(t - w).max() / max(np.abs(w).max(), 1e-300) for t, w in zip(total, want))
    print(f"nccl shared-parameter gradients over {world} GPUs, B={B}: max rel err {err:.2e}", flush=True)
    if not err <= 1e-8:
        ok.zero_()
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if ok.item() == 1 else 1)
