"""Two dependent IVC fold steps through lat_ajtai_fold_step_begin / _finish at the zkVM shape (kappa = 32, n = 98 815): a small
target for the ncu launch list of the whole step (profiles/*_fold_step_launches.csv)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeum_b200 as LB

KAPPA, WL = 32, 19763
DP = LB.GoldiLocksDP
N = WL * DP.L
Q = LB.scheme.Q
rng = np.random.default_rng(0)
scheme = LB.AjtaiCommitmentScheme(KAPPA, N)
for i in range(KAPPA):
    scheme.upload_rows(i, rng.integers(0, 2**63, size=(1, N, 24), dtype=np.uint64))
fs = LB.FoldStep(scheme)
small = np.clip(np.rint(rng.normal(0.0, 350.0, size=(N, 24))), -(2**14), 2**14).astype(np.int64)
acc = np.where(small < 0, small.view(np.uint64) + np.uint64(Q), small.view(np.uint64))
fs.set_accumulator(acc, scheme.commit_coeff(acc))
for step in range(2):
    w = rng.integers(0, 2**63, size=(WL, 24), dtype=np.uint64) % np.uint64(Q)
    cm, ys0, ys1, d16 = fs.begin(w)
    # short challenges, as the protocol's rho_i are: the folded witness stays under the norm bound
    r = rng.integers(-2, 3, size=(2 * DP.K, 24)).astype(np.int64)
    r[:, 1:] = 0
    rho = np.where(r < 0, r.view(np.uint64) + np.uint64(Q), r.view(np.uint64))
    try:
        fs.finish(rho)
    except LB.DigitOverflow:
        fs.set_accumulator(acc, scheme.commit_coeff(acc))
print("ok")
