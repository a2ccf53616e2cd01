"""Sweeps the MAC kernel's launch knobs (stages / resident CTAs / grid) at the zkVM shape and prints the CUDA-event
duration of mac_kernel alone for each setting.  Run on a GPU box: python tools/tune_mac.py [planes]
Compile-time knobs (tile size LAT_TJ_BYTES, LAT_MAC_TRACE, ...) go through latticeum_b200.build.build_variant(tag, defines)
and LAT_LIB=latticeum_b200/lib/variants/<tag>/liblattice_ajtai.so; tools/ab_mac.py times the default plan of a library."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeum_b200 as LB
from latticeum_b200.device import DeviceScheme

planes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
KAPPA, N = 32, 98815
rng = np.random.default_rng(0)
scheme = LB.AjtaiCommitmentScheme(KAPPA, N)
for i in range(KAPPA):
    row = rng.integers(0, 2**63, size=(1, N, 24), dtype=np.uint64)
    scheme.upload_rows(i, row)
eng = DeviceScheme(scheme)
shape = (N, 24) if planes == 1 else (planes, N, 24)
f = torch.from_numpy(rng.integers(0, 2**63, size=shape, dtype=np.int64)).cuda()
cm = eng.new_commitment(planes)
w = torch.from_numpy(rng.integers(0, 2**63, size=(N // 5, 24), dtype=np.int64)).cuda()
MODE = os.environ.get("TUNE_MODE", "commit")


def one():
    if MODE == "witness":
        eng.witness_commit(w, cm)
    else:
        eng.commit_ntt(f, cm)


def run(tag, reps=30):
    for _ in range(5):
        one()
    torch.cuda.synchronize()
    eng.set_profiling(True)
    eng.mac_profile()
    for _ in range(reps):
        one()
    s, c = eng.mac_profile()
    eng.set_profiling(False)
    ms = s / c
    print(f"{tag:40s} mac {ms*1e3:8.1f} us  {KAPPA*N*192/ms/1e6:8.1f} GB/s/plane-equivalent", flush=True)


for stages in (3, 4, 6, 8):
    for gx in (148, 296):
        os.environ["LAT_MAC_STAGES"] = str(stages)
        os.environ["LAT_MAC_GRIDX"] = str(gx)
        try:
            run(f"stages={stages} gridx={gx}")
        except Exception as e:
            print(f"stages={stages} gridx={gx} ERR {str(e)[:80]}")
            torch.cuda.synchronize()
