"""Multi-GPU layer of the batched QP path: one process per GPU, instances sharded by contiguous
ranges, NO data-path collective for (dz, dlam, dnu).  The only exchange is the sum over the batch of the
reverse-mode gradients of parameters that are SHARED across instances (OptNet-style layers: the `+=` over samples in
docs/src/examples/polyhedral_project.jl:95-104, the per-parameter accumulation of src/parameters.jl:355-360): each
rank reduces its shard on the device (``diffopt_b200_qp_batch_param_grads(reduce_over_batch=1)``) and ONE fp64
all-reduce of <= n^2 + mn + pn + n + m + p doubles (75 KB at the headline shape) finishes it.

``torch.distributed`` is plumbing only (NCCL over NVLink on GPUs, gloo in the CPU tests); the arithmetic of the
shard reduction is the CUDA kernel.  Host logic here is backend agnostic so it can be covered on CPU.
"""
from __future__ import annotations

import numpy as np


def shard_range(B: int, rank: int, world: int):
    """Contiguous instance range [lo, hi) of ``rank``: sizes differ by at most one, larger shards first."""
    if not (0 <= rank < world) or B < 0:
        raise ValueError((B, rank, world))
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


PARAM_KEYS = ("dQ", "dq", "dG", "dh", "dA", "db")


def pack_param_grads(grads):
    """(dQ, dq, dG, dh, dA, db) -> one flat fp64 buffer (so that a single collective carries all of them)."""
    shapes = [np.shape(g) for g in grads]
    flat = np.concatenate([np.asarray(g, dtype=np.float64).ravel() for g in grads])
    return flat, shapes


def unpack_param_grads(flat, shapes):
    out, off = [], 0
    for s in shapes:
        k = int(np.prod(s))
        out.append(np.asarray(flat[off:off + k]).reshape(s))
        off += k
    return tuple(out)


def allreduce_shared_param_grads(local_sums, group=None, device=None):
    """Sum of per-rank shard sums over all ranks (every rank gets the total).  ``local_sums`` is the tuple
    returned by ``QPBatch.param_grads(rev, reduce_over_batch=True)`` on this rank's shard.  With the NCCL backend the
    buffer is moved to ``device`` (this rank's GPU) for the collective; with gloo it stays on the host."""
    import torch
    import torch.distributed as dist
    flat, shapes = pack_param_grads(local_sums)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return unpack_param_grads(flat, shapes)
    t = torch.from_numpy(flat)
    if dist.get_backend(group) == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return unpack_param_grads(t.cpu().numpy(), shapes)


def sharded_reverse_shared_params(make_batch, data, seed, rank, world, group=None, device=None):
    """Reverse mode for a layer whose Q, G, A (and/or q, h, b) are shared by all B instances:
    rank solves its shard with ``make_batch(shard_of_data)`` (a ``QPBatch``-like object: ``reverse``, ``param_grads``),
    reduces the parameter gradients over its shard on the device and all-reduces the sums.
    Returns (rev_local, total_param_grads)."""
    B = len(seed)
    lo, hi = shard_range(B, rank, world)
    batch = make_batch({k: v[lo:hi] for k, v in data.items()})
    dz, dl, dn = batch.reverse(seed[lo:hi])
    rev = np.hstack([dz, dl, dn])
    local = batch.param_grads(rev, reduce_over_batch=True) if hi > lo else None
    if local is None:  # empty shard: contribute zeros of the right shapes
        n, m, p = batch.n, batch.m, batch.p
        local = (np.zeros((n, n)), np.zeros(n), np.zeros((m, n)), np.zeros(m), np.zeros((p, n)), np.zeros(p))
    return rev, allreduce_shared_param_grads(local, group=group, device=device)


def nccl_init(ctx, rank, world, group=None):
    """Creates the ctx-owned NCCL communicator (``diffopt_b200_nccl_init``): rank 0 draws the unique id, the id travels
    through ``torch.distributed`` (any backend -- it is 128 bytes of host data), every rank joins.  After this,
    ``diffopt_b200_qp_batch_shared_grads(..., DIFFOPT_QP_ALLREDUCE)`` reduces on the device without the host."""
    import ctypes as C

    import torch
    import torch.distributed as dist
    buf = C.create_string_buffer(128)
    if rank == 0:
        rc = ctx.lib.diffopt_b200_nccl_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise RuntimeError(f"diffopt_b200_nccl_unique_id failed ({rc})")
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if world > 1:
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0, group=group)
        t = t.cpu()
    ident = bytes(t.numpy().tobytes())
    ctx.check(ctx.lib.diffopt_b200_nccl_init(ctx.h, world, rank, C.c_char_p(ident)))


def sharded_reverse_shared_params_device(ctx, Q, G, A, h, z, lam, nu, seed, rank, world):
    """The GPU form of ``sharded_reverse_shared_params``: Q, G, A are ONE instance shared by all B problems.  This rank
    solves its contiguous shard (``diffopt_b200_qp_batch_solve_ex`` with DIFFOPT_QP_SHARED_MATRICES), sums the
    shared-parameter gradients over the shard on the device and -- after ``nccl_init(ctx, rank, world)`` -- all-reduces
    the 75 KB of sums with NCCL without leaving the GPUs.  Returns (rev_local, (dQ, dq, dG, dh, dA, db) totals)."""
    from . import qp as qpm
    B = len(z)
    lo, hi = shard_range(B, rank, world)
    sl = slice(lo, hi)
    _, rev, info = qpm.solve_batch_ex(ctx, Q, G, A, h[sl], z[sl], lam[sl], nu[sl], seed=seed[sl], shared_matrices=True)
    if info.any():
        raise ArithmeticError(f"singular instance {int(np.flatnonzero(info)[0]) + lo}")
    total = qpm.shared_param_grads(ctx, z[sl], lam[sl], nu[sl], rev, allreduce=world > 1)
    return rev, total
