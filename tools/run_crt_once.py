"""Two launches each of the batched CRT and iCRT at 2^20 ring elements: a small target for ncu."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latticeum_b200 import _capi as capi

L = capi.lib()
rng = np.random.default_rng(0)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
cnt = 1 << 20
x = torch.from_numpy(rng.integers(0, 2**63, size=(cnt, 24), dtype=np.int64)).cuda()
y = torch.empty_like(x)
for _ in range(2):
    assert L.lat_ring_crt_dev(x.data_ptr(), cnt, y.data_ptr(), stream) == 0
    assert L.lat_ring_icrt_dev(x.data_ptr(), cnt, y.data_ptr(), stream) == 0
torch.cuda.synchronize()
print("ok")
