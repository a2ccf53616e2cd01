"""Host-side logic of the column-sharded commitment (latticeum_b200/sharded.py) on CPU: world_size 2 and 3 over gloo.
The local engine is a stand-in built on the oracle (tests may use the oracle; the product engine is DeviceScheme on a
GPU) -- what is under test is the shard arithmetic, the all-gather exchange and the mod-q fold order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latticeum_b200.sharded import ShardedCommitPipeline, ShardedAjtaiScheme, shard_bounds
from oracle import c_oracle as CO

Q = 2**64 - 2**32 + 1
KAPPA, W_TOTAL, L, B = 5, 23, 5, 1 << 15


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 23, 98815, 2**20):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def to_t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64))


def to_np(t):
    return t.numpy().view(np.uint64)


class OracleEngine:
    """CPU stand-in with DeviceScheme's interface."""

    def __init__(self, A_local):
        self.A = A_local
        self.kappa = A_local.shape[0]

    def new_commitment(self, batch=1):
        shape = (self.kappa, 24) if batch == 1 else (batch, self.kappa, 24)
        return torch.empty(shape, dtype=torch.int64)

    def witness_commit(self, w_local, cm):
        _, f = CO.witness_from_w_ccs(to_np(w_local), B, L)
        cm.copy_(to_t(CO.commit(self.A, f)))
        return cm

    def commit_ntt(self, f_local, cm):
        f = to_np(f_local)
        if f.ndim == 3:
            cm.copy_(to_t(np.stack([CO.commit(self.A, x) for x in f])))
        else:
            cm.copy_(to_t(CO.commit(self.A, f)))
        return cm

    def fold_partials(self, parts, out):
        p = to_np(parts).astype(object)
        out.copy_(to_t((p.sum(axis=0) % Q).astype(np.uint64)))
        return out


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A = CO.fill_uniform((KAPPA, W_TOTAL * L, 24), 1)
        w = CO.fill_uniform((W_TOTAL, 24), 2)
        lo, hi = shard_bounds(W_TOTAL, world, rank)
        eng = OracleEngine(np.ascontiguousarray(A[:, lo * L : hi * L]))
        sh = ShardedAjtaiScheme(eng)
        assert (sh.world, sh.rank) == (world, rank)
        cm = sh.witness_commit(to_t(w[lo:hi]))
        _, f = CO.witness_from_w_ccs(w, B, L)
        fs = np.stack([f, CO.fill_uniform((W_TOTAL * L, 24), 3)])
        cms = sh.commit_ntt(to_t(np.ascontiguousarray(fs[:, lo * L : hi * L])))
        # the pipelined front end: three steps through two slots, tickets waited for in order
        pipe = ShardedCommitPipeline(sh, hi - lo, depth=2)
        outs = []
        t0 = pipe.submit(to_t(w[lo:hi]))
        t1 = pipe.submit(to_t(w[lo:hi]))
        outs.append(to_np(pipe.wait(t0)).copy())
        t2 = pipe.submit(to_t(w[lo:hi]))
        outs.append(to_np(pipe.wait(t1)).copy())
        outs.append(to_np(pipe.wait(t2)).copy())
        assert np.array_equal(outs[0], to_np(cm)) and np.array_equal(outs[1], to_np(cm)) and np.array_equal(outs[2], to_np(cm))
        results[rank] = (to_np(cm).copy(), to_np(cms).copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_commit_equals_unsharded(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    A = CO.fill_uniform((KAPPA, W_TOTAL * L, 24), 1)
    w = CO.fill_uniform((W_TOTAL, 24), 2)
    _, f = CO.witness_from_w_ccs(w, B, L)
    exp = CO.commit(A, f)
    exp2 = CO.commit(A, CO.fill_uniform((W_TOTAL * L, 24), 3))
    for r in range(world):
        cm, cms = results[r]
        assert np.array_equal(cm, exp), f"rank {r}"
        assert np.array_equal(cms[0], exp) and np.array_equal(cms[1], exp2), f"rank {r}"


def test_single_rank_needs_no_process_group():
    A = CO.fill_uniform((KAPPA, W_TOTAL * L, 24), 1)
    w = CO.fill_uniform((W_TOTAL, 24), 2)
    sh = ShardedAjtaiScheme(OracleEngine(A), world=1, rank=0)
    cm = sh.witness_commit(to_t(w))
    assert np.array_equal(to_np(cm), CO.commit(A, CO.witness_from_w_ccs(w, B, L)[1]))
