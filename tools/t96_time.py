"""Times the kernels built on the Z/(2^96+1) transform network: CRT / iCRT sweep at 2^22 elements, pack + planes of one
fold-step side, and the step's witness kernel + matrix-vector kernel (device-resident).  For A/B runs with LAT_LIB."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeum_b200 as LB
from latticeum_b200 import _capi as capi
from latticeum_b200.device import DeviceScheme

L = capi.lib()
rng = np.random.default_rng(0)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
tag = (os.environ.get("LAT_LIB") or "x/default/x").split("/")[-2]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


cnt = 1 << 22
x = torch.from_numpy(rng.integers(0, 2**63, size=(cnt, 24), dtype=np.int64)).cuda()
y = torch.empty_like(x)
t_crt = timeit(lambda: L.lat_ring_crt_dev(x.data_ptr(), cnt, y.data_ptr(), stream))
t_icrt = timeit(lambda: L.lat_ring_icrt_dev(x.data_ptr(), cnt, y.data_ptr(), stream))
del x, y
KAPPA, N, WL = 32, 98815, 19763
scheme = LB.AjtaiCommitmentScheme(KAPPA, N)
row = rng.integers(0, 2**63, size=(1, N, 24), dtype=np.uint64)
for i in range(KAPPA):
    scheme.upload_rows(i, row)
eng = DeviceScheme(scheme)
v = rng.integers(-(2**14), 2**14 + 1, size=(N, 24), dtype=np.int64)
fc = np.where(v < 0, v.view(np.uint64) + np.uint64(LB.scheme.Q), v.view(np.uint64))
fc_dev = eng.to_device(fc)
t_planes = timeit(lambda: L.lat_ajtai_decompose_commit_dev(scheme._h, fc_dev.data_ptr(), N, None, None, None, None))
w = torch.from_numpy(rng.integers(0, 2**63, size=(WL, 24), dtype=np.int64)).cuda()
cm = eng.new_commitment(1)
t_step = timeit(lambda: L.lat_ajtai_witness_from_w_ccs_dev(scheme._h, w.data_ptr(), WL, None, None, cm.data_ptr()), reps=50)
t_wit = timeit(lambda: L.lat_ajtai_witness_from_w_ccs_dev(scheme._h, w.data_ptr(), WL, None, None, None), reps=50)
print(f"{tag}: crt {t_crt:.1f} us  icrt {t_icrt:.1f} us  pack+planes {t_planes:.1f} us  step {t_step:.1f} us  witness-only {t_wit:.1f} us", flush=True)
