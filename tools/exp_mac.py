"""Step-level and kernel-level timing of the zkVM step commit for the library selected by LAT_LIB (A/B of kernel variants):
    for v in a b; do LAT_LIB=latticeum_b200/lib/variants/$v/liblattice_ajtai.so python tools/exp_mac.py; done"""
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
for i in range(KAPPA):
    scheme.upload_rows(i, rng.integers(0, 2**63, size=(1, N, 24), dtype=np.uint64))
eng = DeviceScheme(scheme)
w = torch.from_numpy(rng.integers(0, 2**63, size=(N // 5, 24), dtype=np.int64)).cuda()
f = torch.from_numpy(rng.integers(0, 2**63, size=(N, 24), dtype=np.int64)).cuda()
cm = eng.new_commitment()


def timed(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def mac_us(fn, reps):
    eng.set_profiling(True)
    eng.mac_profile()
    for _ in range(reps):
        fn()
    s, c = eng.mac_profile()
    eng.set_profiling(False)
    return s / c * 1e3


tag = (os.environ.get("LAT_LIB") or "x/default/x").split("/")[-2]
out = [tag]
for name, fn in (("witness_commit", lambda: eng.witness_commit(w, cm)), ("commit_ntt", lambda: eng.commit_ntt(f, cm))):
    for rep in range(2):
        out.append(f"{name}: step {timed(fn, 50):7.1f} us, mac {mac_us(fn, 30):7.1f} us")
eng.set_step_overlap(True)
out.append(f"witness_commit overlapped: step {timed(lambda: eng.witness_commit(w, cm), 50):7.1f} us")
print(" | ".join(out), flush=True)
