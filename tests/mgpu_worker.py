"""Multi-GPU parity worker (run under torchrun, one rank per GPU; driven by tests/test_multigpu.py).

Every rank commits its column block with the CUDA engine, the partial commitments are exchanged (LAT_EXCHANGE = nccl |
p2p) and folded mod q, and the result is compared with the ORACLE's commitment of the whole witness against the whole
matrix.  Covered: the device path with step overlap on (the exchange kernel inside the programmatic-launch chain), the
host-buffer pipeline (lat_ajtai_set_peers + submit/wait) on the legacy stream with skewed ranks, the strong-scaled
commit_ntt of BASELINE configs[4]'s shape, and the column-sharded fold step (decompose_commit per side + fold_witness).
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeum_b200 as LB
from latticeum_b200 import _capi as capi
from latticeum_b200.device import DeviceScheme
from latticeum_b200.sharded import ShardedAjtaiScheme, ShardedCommitPipeline, shard_bounds
from oracle import c_oracle as CO  # checker

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mode = os.environ.get("LAT_EXCHANGE", "auto")
KAPPA, L, K = 32, 5, 15
failures = []


def check(name, ok):
    if not ok:
        failures.append(name)
    if rank == 0:
        print(f"mgpu_worker world={world} exchange={mode}: {name}: {'ok' if ok else 'MISMATCH'}", flush=True)


def shard_scheme(A, lo, hi, n_total):
    """This rank's engine on columns [lo, hi), uploaded straight out of the full-width host matrix."""
    s = LB.AjtaiCommitmentScheme(KAPPA, hi - lo, device=local)
    st = capi.lib().lat_ajtai_upload_rows(s._h, 0, KAPPA, A.ctypes.data + lo * 24 * 8, n_total)
    assert st == 0, capi.last_error()
    return s


# ---- 1. Witness::from_w_ccs + commit, column-sharded, consecutive steps overlapped on the device --------------------------
W_TOTAL = 4001 * world + 3
n_total = W_TOTAL * L
A = CO.fill_uniform((KAPPA, n_total, 24), 11)  # same seeded inputs on every rank
w = CO.fill_uniform((W_TOTAL, 24), 12)
lo, hi = shard_bounds(W_TOTAL, world, rank)
scheme = shard_scheme(A, lo * L, hi * L, n_total)
eng = DeviceScheme(scheme)  # bound to torch's current stream = the LEGACY stream
sh = ShardedAjtaiScheme(eng, exchange=mode)
if mode in ("p2p", "nccl"):
    assert sh.exchange == mode, sh.exchange
eng.set_step_overlap(True)
w_dev = eng.to_device(w[lo:hi])
_, f_full = CO.witness_from_w_ccs(w, 1 << 15, L)
exp = CO.commit(A, f_full)
outs = [sh.witness_commit(w_dev).clone() for _ in range(6)]  # several epochs: both mailbox slots, chained kernels
torch.cuda.synchronize()
check("device path, step overlap on", all(np.array_equal(DeviceScheme.to_numpy(o), exp) for o in outs))

# ---- 2. host-buffer pipeline on the legacy stream, ranks skewed -----------------------------------------------------------
pipe = ShardedCommitPipeline(sh, hi - lo)
w_pin = torch.from_numpy(w[lo:hi].view(np.int64).copy()).pin_memory()
got, tickets = [], []
for k in range(3 * pipe.depth):
    if k >= pipe.depth:
        got.append(DeviceScheme.to_numpy(pipe.wait(tickets[k - pipe.depth])).copy())
    if k in (0, 1, pipe.depth + 1):
        time.sleep(0.05 * rank)  # skew: the first submits of every rank happen at different times
    tickets.append(pipe.submit(w_pin))
for k in range(2 * pipe.depth, 3 * pipe.depth):
    got.append(DeviceScheme.to_numpy(pipe.wait(tickets[k])).copy())
torch.cuda.synchronize()
check(f"host-buffer pipeline ({'native submit/wait' if pipe.native else 'chained' if pipe.chain else 'events'}), skewed ranks",
      len(got) == 3 * pipe.depth and all(np.array_equal(g, exp) for g in got))
eng.set_step_overlap(False)
scheme.close()

# ---- 3. configs[4] shape: one CRT-form witness, strong-scaled by columns (commit_ntt) ---------------------------------------
n2 = (1 << 15) + 5
A2 = CO.fill_uniform((KAPPA, n2, 24), 21)
f2 = CO.fill_uniform((2, n2, 24), 22)
lo2, hi2 = shard_bounds(n2, world, rank)
scheme2 = shard_scheme(A2, lo2, hi2, n2)
eng2 = DeviceScheme(scheme2)
sh2 = ShardedAjtaiScheme(eng2, exchange=mode)
cm1 = sh2.commit_ntt(eng2.to_device(f2[0, lo2:hi2]))
cmb = sh2.commit_ntt(eng2.to_device(np.ascontiguousarray(f2[:, lo2:hi2])))
torch.cuda.synchronize()
e0, e1 = CO.commit(A2, f2[0]), CO.commit(A2, f2[1])
check("strong-scaled commit_ntt", np.array_equal(DeviceScheme.to_numpy(cm1), e0))
check("strong-scaled commit_ntt batch", np.array_equal(DeviceScheme.to_numpy(cmb), np.stack([e0, e1])))

# ---- 4. the fold step, column-sharded: decompose_commit on both sides + fold_witness ----------------------------------------
if hasattr(sh2, "decompose_commit"):
    rng = np.random.Generator(np.random.PCG64(31))
    Q = 2**64 - 2**32 + 1
    sides = []
    for side in range(2):
        small = rng.integers(-(1 << 14), (1 << 14) + 1, size=(n2, 24), dtype=np.int64)
        sides.append(np.where(small < 0, small.view(np.uint64) + np.uint64(Q), small.view(np.uint64)))  # wraps to small + q
    rho = CO.fill_uniform((2 * K, 24), 33)
    ok = True
    planes_full = []
    for side, fc in enumerate(sides):
        cm_full = CO.commit(A2, CO.crt(fc))
        cms = sh2.decompose_commit(eng2.to_device(fc[lo2:hi2]), eng2.to_device(cm_full), side=side)
        torch.cuda.synchronize()
        pc, pf, exp_cms = CO.decompose_commit(A2, fc, cm_full, 2, K, want_planes=True)
        planes_full.append(pf)
        ok = ok and np.array_equal(DeviceScheme.to_numpy(cms), exp_cms)
    check("sharded decompose_commit (both sides)", ok)
    f0_local, f0c_local = sh2.fold_witness(eng2.to_device(rho))
    torch.cuda.synchronize()
    f0 = CO.compute_f0(rho, [pf[k] for pf in planes_full for k in range(K)])
    check("sharded fold_witness", np.array_equal(DeviceScheme.to_numpy(f0_local), f0[lo2:hi2])
          and np.array_equal(DeviceScheme.to_numpy(f0c_local), CO.icrt(f0[lo2:hi2])))
scheme2.close()

flag = torch.tensor([0 if failures else 1], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
if failures:
    print(f"rank {rank}: FAILED {failures}", flush=True)
sys.exit(0 if int(flag.item()) == 1 else 1)
