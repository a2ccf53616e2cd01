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

    K = 15

    def decompose_commit(self, f_coeff_local, cm, cms, side=None):
        # the local K-1 matrix commits of this column block; y_0 only when the caller hands over a commitment
        fc = to_np(f_coeff_local)
        cm_np = to_np(cm) if cm is not None else np.zeros((self.kappa, 24), np.uint64)
        _, pf, out = CO.decompose_commit(self.A, fc, cm_np, 2, self.K, want_planes=True)
        self.planes = getattr(self, "planes", {})
        self.planes[side or 0] = pf
        if cm is None:
            cms[1:].copy_(to_t(out[1:]))
        else:
            cms.copy_(to_t(out))
        return cms

    def y0(self, cm, cms):
        acc = np.zeros((self.kappa, 24), dtype=object)
        c = to_np(cms).astype(object)
        for k in range(self.K - 1, 0, -1):  # decomposition.rs:189-197: fold_rev((acc + y_k) * b), b = 2
            acc = (acc + c[k]) * 2 % Q
        cms[0].copy_(to_t(((to_np(cm).astype(object) - acc) % Q).astype(np.uint64)))
        return cms

    def fold_witness(self, rho):
        f0 = CO.compute_f0(to_np(rho), [self.planes[s][k] for s in (0, 1) for k in range(self.K)])
        return to_t(f0), to_t(CO.icrt(f0))

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
        # the fold step, column-sharded: K-1 partial commitments per side exchanged at once, y_0 from the totals
        n = W_TOTAL * L
        rng = np.random.Generator(np.random.PCG64(5))
        fold = {}
        for side in (0, 1):
            small = rng.integers(-(1 << 14), (1 << 14) + 1, size=(n, 24), dtype=np.int64)
            fc = np.where(small < 0, small.view(np.uint64) + np.uint64(Q), small.view(np.uint64))  # wraps to small + q
            cm_full = CO.commit(A, CO.crt(fc))
            ys = sh.decompose_commit(to_t(np.ascontiguousarray(fc[lo * L : hi * L])), to_t(cm_full), side=side)
            fold[side] = to_np(ys).copy()
        rho = CO.fill_uniform((30, 24), 6)
        f0_local, f0c_local = sh.fold_witness(to_t(rho))
        results[rank] = (to_np(cm).copy(), to_np(cms).copy(), fold, to_np(f0_local).copy(), to_np(f0c_local).copy())
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
    n = W_TOTAL * L
    rng = np.random.Generator(np.random.PCG64(5))
    exp_fold, planes = {}, []
    for side in (0, 1):
        small = rng.integers(-(1 << 14), (1 << 14) + 1, size=(n, 24), dtype=np.int64)
        fc = np.where(small < 0, small.view(np.uint64) + np.uint64(Q), small.view(np.uint64))  # wraps to small + q
        _, pf, ys = CO.decompose_commit(A, fc, CO.commit(A, CO.crt(fc)), 2, 15, want_planes=True)
        exp_fold[side] = ys
        planes += [pf[k] for k in range(15)]
    f0 = CO.compute_f0(CO.fill_uniform((30, 24), 6), planes)
    for r in range(world):
        cm, cms, fold, f0_local, f0c_local = results[r]
        assert np.array_equal(cm, exp), f"rank {r}"
        assert np.array_equal(cms[0], exp) and np.array_equal(cms[1], exp2), f"rank {r}"
        assert np.array_equal(fold[0], exp_fold[0]) and np.array_equal(fold[1], exp_fold[1]), f"rank {r}: sharded decompose_commit"
        lo, hi = shard_bounds(W_TOTAL, world, r)
        assert np.array_equal(f0_local, f0[lo * L : hi * L]) and np.array_equal(f0c_local, CO.icrt(f0[lo * L : hi * L]))


def test_single_rank_needs_no_process_group():
    A = CO.fill_uniform((KAPPA, W_TOTAL * L, 24), 1)
    w = CO.fill_uniform((W_TOTAL, 24), 2)
    sh = ShardedAjtaiScheme(OracleEngine(A), world=1, rank=0)
    cm = sh.witness_commit(to_t(w))
    assert np.array_equal(to_np(cm), CO.commit(A, CO.witness_from_w_ccs(w, B, L)[1]))
