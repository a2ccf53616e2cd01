"""Multi-GPU parity check of the column-sharded commitment (run under torchrun, one rank per GPU):
every rank commits its column block with the CUDA engine, the partials are all-gathered over NCCL and folded mod q,
and rank 0 compares the result with the oracle's commitment of the WHOLE witness against the WHOLE matrix."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeum_b200 as LB
from latticeum_b200.device import DeviceScheme
from latticeum_b200.sharded import ShardedAjtaiScheme, shard_bounds
from oracle import c_oracle as CO  # checker

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
KAPPA, W_TOTAL, L = 32, 4001 * world + 3, 5
n_total = W_TOTAL * L
A = CO.fill_uniform((KAPPA, n_total, 24), 11)      # same seeded inputs on every rank
w = CO.fill_uniform((W_TOTAL, 24), 12)
lo, hi = shard_bounds(W_TOTAL, world, rank)
scheme = LB.AjtaiCommitmentScheme(KAPPA, (hi - lo) * L, device=local)
# upload this rank's column block straight out of the full-width host matrix (row_stride = full width)
from latticeum_b200 import _capi as capi
st = capi.lib().lat_ajtai_upload_rows(scheme._h, 0, KAPPA, A.ctypes.data + lo * L * 24 * 8, n_total)
assert st == 0, capi.last_error()
eng = DeviceScheme(scheme)
mode = os.environ.get("LAT_EXCHANGE", "auto")
sh = ShardedAjtaiScheme(eng, exchange=mode)
w_dev = eng.to_device(w[lo:hi])
for _ in range(5):  # several epochs: exercises both receive slots
    cm = sh.witness_commit(w_dev)
torch.cuda.synchronize()
got = DeviceScheme.to_numpy(cm)
ok = True
if rank == 0:
    _, f = CO.witness_from_w_ccs(w, 1 << 15, L)
    exp = CO.commit(A, f)
    ok = bool(np.array_equal(got, exp))
    print(f"mgpu_check world={world} exchange={sh.exchange}: sharded commitment {'==' if ok else '!='} oracle (n_total={n_total})", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
