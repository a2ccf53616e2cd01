"""Time the standalone negacyclic NTT over 2^24 coefficients for every supported d; CUDA events."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latticeum_b200 import _capi as capi

L = capi.lib()
rng = np.random.default_rng(0)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
LG_TOTAL = 24
PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


x = torch.from_numpy(rng.integers(0, 2**63, size=1 << LG_TOTAL, dtype=np.int64)).cuda()
y = torch.empty_like(x)
rows = []
for lg_d in [int(a) for a in sys.argv[1:]] or range(1, 15):
    polys = 1 << (LG_TOTAL - lg_d)
    t_f = timeit(lambda: L.lat_ntt_negacyclic_dev(x.data_ptr(), polys, lg_d, 0, y.data_ptr(), stream))
    t_i = timeit(lambda: L.lat_ntt_negacyclic_dev(x.data_ptr(), polys, lg_d, 1, y.data_ptr(), stream))
    floor_us = (1 << LG_TOTAL) * 16 / 6548.8e9 * 1e6
    rows.append({"log2_d": lg_d, "fwd_us": round(t_f * 1e3, 1), "inv_us": round(t_i * 1e3, 1),
                 "hbm_floor_us": round(floor_us, 1), "fwd_frac": round(floor_us / (t_f * 1e3), 3),
                 "inv_frac": round(floor_us / (t_i * 1e3), 3)})
    print(rows[-1], flush=True)
out = os.environ.get("NTT_OUT")
if out:
    json.dump(rows, open(out, "w"), indent=1)
