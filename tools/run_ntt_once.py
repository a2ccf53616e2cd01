"""A few launches of the standalone negacyclic NTT (default d = 64 and 4096; 2^24 coefficients): a small target for ncu."""
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
for lg_d in [int(v) for v in sys.argv[1:]] or (6, 12):
    polys = 1 << (24 - lg_d)
    x = torch.from_numpy(rng.integers(0, 2**63, size=(polys, 1 << lg_d), dtype=np.int64)).cuda()
    y = torch.empty_like(x)
    for inverse in (0, 1):
        for _ in range(2):
            assert L.lat_ntt_negacyclic_dev(x.data_ptr(), polys, lg_d, inverse, y.data_ptr(), stream) == 0
torch.cuda.synchronize()
print("ok")
