"""One fold-side call (decompose_witness + commit_witnesses) at the zkVM shape: a small target for ncu."""
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
for _ in range(3):
    eng.decompose_commit(fc_dev, cm, cms)
eng.synchronize()
print("ok")
