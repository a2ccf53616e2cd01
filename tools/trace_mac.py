"""Per-CTA timeline of one mac_kernel launch (tuning builds with -DLAT_MAC_TRACE only):
    python -c "import latticeum_b200.build as b; print(b.build_variant('trace', ['LAT_MAC_TRACE']))"
    LAT_LIB=latticeum_b200/lib/variants/trace/liblattice_ajtai.so python tools/trace_mac.py [planes]
Prints, relative to the first CTA's entry: when CTAs enter, get their first tile, leave the loop, have published their
partials, and exit -- i.e. where the fixed cost of a launch goes."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeum_b200 as LB
from latticeum_b200 import _capi as capi
from latticeum_b200.device import DeviceScheme

planes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
KAPPA, N = 32, int(os.environ.get("TRACE_N", 98815))
rng = np.random.default_rng(0)
scheme = LB.AjtaiCommitmentScheme(KAPPA, N)
row = rng.integers(0, 2**63, size=(1, N, 24), dtype=np.uint64)
for i in range(KAPPA):
    scheme.upload_rows(i, row)
eng = DeviceScheme(scheme)
shape = (N, 24) if planes == 1 else (planes, N, 24)
f = torch.from_numpy(rng.integers(0, 2**63, size=shape, dtype=np.int64)).cuda()
cm = eng.new_commitment(planes)
for _ in range(5):
    eng.commit_ntt(f, cm)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.commit_ntt(f, cm)
b.record()
torch.cuda.synchronize()
print(f"commit_ntt (fext + mac) event time {a.elapsed_time(b)*1e3:.1f} us")
raw = C.CDLL(capi.LIB_PATH)
nct = int(os.environ.get("TRACE_CTAS", 296 if planes == 1 else 148))
buf = (C.c_ulonglong * (nct * 8))()
assert raw.lat_debug_mac_trace(buf, nct) == 0
full = np.frombuffer(buf, dtype=np.uint64).reshape(nct, 8)
smid = full[:, 5].astype(np.int64)
t = full[:, :5].astype(np.int64)
t0 = t[:, 0].min()
t = (t - t0) / 1e3
names = ["enter", "first tile", "loop done", "published", "exit"]
for k, nm in enumerate(names):
    c = t[:, k]
    print(f"{nm:12s} min {c.min():8.2f}  p10 {np.percentile(c,10):8.2f}  median {np.median(c):8.2f}  p90 {np.percentile(c,90):8.2f}  max {c.max():8.2f} us")
d = t[:, 2] - t[:, 1]
print(f"loop time    min {d.min():8.2f}  median {np.median(d):8.2f}  max {d.max():8.2f} us")
d = t[:, 3] - t[:, 2]
print(f"finish+REDs+fence  min {d.min():8.2f}  median {np.median(d):8.2f}  max {d.max():8.2f} us")

# imbalance: within an SM (its CTAs) or across SMs?
loop_done = t[:, 2]
per_sm = {}
for c in range(nct):
    per_sm.setdefault(int(smid[c]), []).append(float(loop_done[c]))
last = np.array([max(v) for v in per_sm.values()])
first = np.array([min(v) for v in per_sm.values()])
print(f"SMs used {len(per_sm)}; CTAs per SM {sorted(set(len(v) for v in per_sm.values()))}")
print(f"per SM, last CTA done:  min {last.min():8.2f} median {np.median(last):8.2f} max {last.max():8.2f} us")
print(f"per SM, first CTA done: min {first.min():8.2f} median {np.median(first):8.2f} max {first.max():8.2f} us")
order = np.argsort(list(per_sm.keys()))
keys = np.array(list(per_sm.keys()))[order]
print("smid : last-done (us), every 8th SM:", [(int(k), round(float(last[order][i]), 1)) for i, k in enumerate(keys)][::8])
print("smid of CTAs 0..7:", smid[:8].tolist(), " CTAs", nct // 2, "..:", smid[nct // 2 : nct // 2 + 8].tolist())
pairs = {}
for c in range(nct):
    pairs.setdefault(int(smid[c]), []).append(c)
diffs = sorted(set(abs(v[1] - v[0]) for v in pairs.values() if len(v) == 2))
print("blockIdx distance between the two CTAs of an SM:", diffs[:10])
slow_is_higher = sum(1 for v in pairs.values() if len(v) == 2 and loop_done[max(v)] > loop_done[min(v)])
print("SMs where the higher-numbered CTA finishes later:", slow_is_higher, "of", len(pairs))
