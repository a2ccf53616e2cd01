"""A/B timing of mac_kernel (default plan) for the library selected by LAT_LIB: planes = 1 (Karatsuba, 3-word matrix) and 14
(Toom-3, 5-word matrix, 12 + 2 planes) at the zkVM shape.
    for v in a b; do LAT_LIB=latticeum_b200/lib/variants/$v/liblattice_ajtai.so python tools/ab_mac.py; done"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeum_b200 as LB
from latticeum_b200.device import DeviceScheme

KAPPA, N = 32, 98815
rng = np.random.default_rng(0)
scheme = LB.AjtaiCommitmentScheme(KAPPA, N)
row = rng.integers(0, 2**63, size=(1, N, 24), dtype=np.uint64)
for i in range(KAPPA):
    scheme.upload_rows(i, row)
eng = DeviceScheme(scheme)
res = []
for planes in (1, 14):
    shape = (N, 24) if planes == 1 else (planes, N, 24)
    f = torch.from_numpy(rng.integers(0, 2**63, size=shape, dtype=np.int64)).cuda()
    cm = eng.new_commitment(planes)
    for _ in range(5):
        eng.commit_ntt(f, cm)
    torch.cuda.synchronize()
    eng.set_profiling(True)
    eng.mac_profile()
    for _ in range(30):
        eng.commit_ntt(f, cm)
    s, c = eng.mac_profile()
    eng.set_profiling(False)
    res.append(f"planes={planes}: {s / 30 * 1e3:8.1f} us per call in {c // 30} launch(es)")  # 14 planes = 12 + 2: two launches
print((os.environ.get("LAT_LIB") or "x/default/x").split("/")[-2], " ".join(res), flush=True)
