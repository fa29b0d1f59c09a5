"""Host logic of the multi-GPU layer on CPU: shard ranges, packing, and the shared-parameter all-reduce with
world_size 2 over gloo (the GPU arithmetic is replaced by an oracle-backed stand-in batch; this is test code)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import diffopt_b200
from oracle import qp as oqp


def test_shard_ranges_cover_the_batch_exactly():
    sh = diffopt_b200.submodule("sharding")
    for B in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            r = [sh.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


def test_pack_roundtrip():
    sh = diffopt_b200.submodule("sharding")
    rng = np.random.default_rng(0)
    g = (rng.standard_normal((4, 4)), rng.standard_normal(4), rng.standard_normal((3, 4)), rng.standard_normal(3),
         rng.standard_normal((2, 4)), rng.standard_normal(2))
    flat, shapes = sh.pack_param_grads(g)
    assert flat.size == 16 + 4 + 12 + 3 + 8 + 2
    for a, b in zip(g, sh.unpack_param_grads(flat, shapes)):
        assert np.array_equal(a, b)


class _OracleBatch:
    """Stand-in with QPBatch's interface whose arithmetic is the CPU oracle (CPU test only)."""

    def __init__(self, d):
        self.d = d
        self.B = len(d["z"])
        self.n, self.m, self.p = d["Q"].shape[1], d["G"].shape[1], d["A"].shape[1]

    def reverse(self, seed):
        out = [oqp.reverse(self.d["Q"][b], self.d["G"][b], self.d["h"][b], self.d["A"][b], self.d["z"][b],
                           self.d["lam"][b], self.d["nu"][b], seed[b]) for b in range(self.B)]
        n, m, p = self.n, self.m, self.p
        stack = lambda k, w: np.array([o[k] for o in out]).reshape(self.B, w)
        return stack(0, n), stack(1, m), stack(2, p)

    def param_grads(self, rev, reduce_over_batch=False):
        n, m = self.n, self.m
        g = [oqp.reverse_param_grads(self.d["z"][b], self.d["lam"][b], self.d["nu"][b], rev[b, :n], rev[b, n:n + m],
                                     rev[b, n + m:]) for b in range(self.B)]
        assert reduce_over_batch
        return tuple(sum(x[k] for x in g) for k in range(6))


def _data(B=7, n=6, m=5, p=2):
    import bench_data
    d = bench_data.qp_batch(B, n, m, p, n_active=2, seed0=11, shared=True)
    return {k: d[k] for k in ("Q", "G", "A", "h", "z", "lam", "nu")}, d["seed"]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = diffopt_b200.submodule("sharding")
        data, seed = _data()
        rev, total = sh.sharded_reverse_shared_params(_OracleBatch, data, seed, rank, world)
        lo, hi = sh.shard_range(len(seed), rank, world)
        np.savez(os.path.join(out, f"r{rank}.npz"), rev=rev, lo=lo, hi=hi, **{k: v for k, v in zip(sh.PARAM_KEYS, total)})
    finally:
        dist.destroy_process_group()


def test_shared_param_allreduce_world2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sh = diffopt_b200.submodule("sharding")
    data, seed = _data()
    whole = _OracleBatch(data)
    dz, dl, dn = whole.reverse(seed)
    rev_all = np.hstack([dz, dl, dn])
    want = whole.param_grads(rev_all, reduce_over_batch=True)
    got_rev = np.zeros_like(rev_all)
    for r in range(2):
        f = np.load(tmp_path / f"r{r}.npz")
        got_rev[int(f["lo"]):int(f["hi"])] = f["rev"]
        for k, w in zip(sh.PARAM_KEYS, want):     # every rank holds the full sum
            assert np.allclose(f[k], w, rtol=1e-12, atol=1e-12), k
    assert np.allclose(got_rev, rev_all, rtol=0, atol=0)   # sharding does not change per-instance results
