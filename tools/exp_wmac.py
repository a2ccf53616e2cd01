"""Tuning: where does the pipelined step-commit kernel (wmac_kernel) spend its time?  Needs the debug build:
    python -c "import latticeum_b200.build as b; print(b.build_variant('wdbg', ['LAT_WMAC_DEBUG', 'LAT_MAC_TRACE']))"
    LAT_LIB=latticeum_b200/lib/variants/wdbg/liblattice_ajtai.so python tools/exp_wmac.py
Results with the debug bits set are WRONG on purpose (arithmetic skipped); only the timing is of interest."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeum_b200 as LB
from latticeum_b200 import _capi as capi
from latticeum_b200.device import DeviceScheme

KAPPA, N = 32, 98815
rng = np.random.default_rng(0)
scheme = LB.AjtaiCommitmentScheme(KAPPA, N)
for i in range(KAPPA):
    scheme.upload_rows(i, rng.integers(0, 2**63, size=(1, N, 24), dtype=np.uint64))
eng = DeviceScheme(scheme)
w = torch.from_numpy(rng.integers(0, 2**63, size=(N // 5, 24), dtype=np.int64)).cuda()
cm = eng.new_commitment()
raw = C.CDLL(capi.LIB_PATH)


def timed(reps=50):
    for _ in range(5):
        eng.witness_commit(w, cm)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.witness_commit(w, cm)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for dbg in (0, 7, 15, 31, 16):
    assert raw.lat_debug_wmac(dbg) == 0
    print(f"debug bits {dbg}: step {timed():7.1f} us", flush=True)
raw.lat_debug_wmac(0)
# per-CTA timeline of one launch
eng.witness_commit(w, cm)
torch.cuda.synchronize()
nct = 296
buf = (C.c_ulonglong * (nct * 8))()
assert raw.lat_debug_mac_trace(buf, nct) == 0
full = np.frombuffer(buf, dtype=np.uint64).reshape(nct, 8)
t = full[:, :5].astype(np.int64)
t = (t - t[:, 0].min()) / 1e3
for k, nm in enumerate(["enter", "first tile", "loop done", "published", "exit"]):
    c = t[:, k]
    print(f"{nm:12s} min {c.min():8.2f}  median {np.median(c):8.2f}  max {c.max():8.2f} us")
