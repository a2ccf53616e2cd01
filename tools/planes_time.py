import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.getcwd())
import latticeum_b200 as LB
from latticeum_b200 import _capi as capi
from latticeum_b200.device import DeviceScheme
KAPPA, N, K = 32, 98815, 15
rng = np.random.default_rng(0)
scheme = LB.AjtaiCommitmentScheme(KAPPA, N)
eng = DeviceScheme(scheme)
v = rng.integers(-(2**14), 2**14 + 1, size=(N, 24), dtype=np.int64)
fc = np.where(v < 0, v.view(np.uint64) + np.uint64(LB.scheme.Q), v.view(np.uint64))
fc_dev = eng.to_device(fc)
L = capi.lib()
def run():
    L.lat_ajtai_decompose_commit_dev(scheme._h, fc_dev.data_ptr(), N, None, None, None, None)
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): run()
b.record(); torch.cuda.synchronize()
print((os.environ.get("LAT_LIB") or "x/default/x").split("/")[-2], "pack+planes", round(a.elapsed_time(b) / 20 * 1e3, 1), "us")
