"""Both fold sides (decompose_witness + commit_witnesses) and the fold of the 2K planes (compute_f_0 + iCRT) at the zkVM
shape: a small target for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeum_b200 as LB
from latticeum_b200.device import DeviceScheme
KAPPA, N, K = 32, 98815, 15
rng = np.random.default_rng(0)
scheme = LB.AjtaiCommitmentScheme(KAPPA, N)
for i in range(KAPPA):
    scheme.upload_rows(i, rng.integers(0, 2**63, size=(1, N, 24), dtype=np.uint64))
eng = DeviceScheme(scheme)
v = rng.integers(-(2**14), 2**14 + 1, size=(N, 24), dtype=np.int64)
fc = np.where(v < 0, np.uint64(LB.scheme.Q) - (-v).astype(np.uint64), v.astype(np.uint64)).astype(np.uint64)
fc_dev = eng.to_device(fc)
cm = eng.new_commitment()
cms = torch.empty((K, KAPPA, 24), dtype=torch.int64, device="cuda")
eng.commit_ntt(eng.to_device(rng.integers(0, 2**63, size=(N, 24), dtype=np.uint64)), cm)
from latticeum_b200 import _capi as capi
L = capi.lib()
for side in (0, 1):
    assert L.lat_ajtai_select_side(scheme._h, side) == 0
    eng.decompose_commit(fc_dev, cm, cms)
rho = eng.to_device(rng.integers(0, 2**63, size=(2 * K, 24), dtype=np.uint64))
f0 = torch.empty((N, 24), dtype=torch.int64, device="cuda")
f0c = torch.empty_like(f0)
assert L.lat_ajtai_fold_witness_dev(scheme._h, rho.data_ptr(), f0.data_ptr(), f0c.data_ptr()) == 0
eng.synchronize()
print("ok")
