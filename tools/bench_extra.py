"""Secondary measurements (BASELINE.json configs[2], configs[3]): the fold-step batch (decompose + 14 commits per side)
and the standalone CRT/iCRT sweep, device-resident, CUDA-event timed.  Writes gpurun_out/extra.json."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latticeum_b200 as LB
from latticeum_b200 import _capi as capi
from latticeum_b200.device import DeviceScheme

KAPPA, N, K = 32, 98815, 15
rng = np.random.default_rng(0)
out = {}


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# ---- transform sweep (config 3) -------------------------------------------------------------------------------------
L = capi.lib()
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
sweep = []
for lg in range(10, 23, 2):
    cnt = 1 << lg
    x = torch.from_numpy(rng.integers(0, 2**63, size=(cnt, 24), dtype=np.int64)).cuda()
    y = torch.empty_like(x)
    t_crt = timeit(lambda: L.lat_ring_crt_dev(x.data_ptr(), cnt, y.data_ptr(), stream))
    t_icrt = timeit(lambda: L.lat_ring_icrt_dev(x.data_ptr(), cnt, y.data_ptr(), stream))
    sweep.append({"log2_count": lg, "crt_us": t_crt * 1e3, "icrt_us": t_icrt * 1e3,
                  "crt_GBps": cnt * 384 / t_crt / 1e6, "icrt_GBps": cnt * 384 / t_icrt / 1e6})
    print(sweep[-1], flush=True)
    del x, y
out["transform_sweep"] = sweep

# ---- standalone power-of-two negacyclic NTT (not on the drop-in path) ---------------------------------------
ntt = []
for lg_d, lg_total in ((6, 24), (8, 24), (10, 24), (12, 24), (14, 24)):
    d, polys = 1 << lg_d, 1 << (lg_total - lg_d)
    x = torch.from_numpy(rng.integers(0, 2**63, size=(polys, d), dtype=np.int64)).cuda()
    y = torch.empty_like(x)
    t_f = timeit(lambda: L.lat_ntt_negacyclic_dev(x.data_ptr(), polys, lg_d, 0, y.data_ptr(), stream))
    t_i = timeit(lambda: L.lat_ntt_negacyclic_dev(x.data_ptr(), polys, lg_d, 1, y.data_ptr(), stream))
    # wide multiplies per butterfly: 5 in a general stage (4 for the product, 1 in the fold), 1 in the (up to 4) shift stages
    wide = polys * (d // 2) * (5 * max(lg_d - 4, 0) + min(lg_d, 4))
    ntt.append({"log2_d": lg_d, "polys": polys, "fwd_us": t_f * 1e3, "inv_us": t_i * 1e3,
                "coeffs_per_s": polys * d / (t_f * 1e-3), "GBps": polys * d * 16 / t_f / 1e6,
                "imad_pipe_frac": wide / (t_f * 1e-3) / 9.154e12})
    print(ntt[-1], flush=True)
    del x, y
out["ntt_negacyclic"] = ntt

# ---- fold-step batch (config 2) ---------------------------------------------------------------------------------------
scheme = LB.AjtaiCommitmentScheme(KAPPA, N)
for i in range(KAPPA):
    scheme.upload_rows(i, rng.integers(0, 2**63, size=(1, N, 24), dtype=np.uint64))
eng = DeviceScheme(scheme)
v = rng.integers(-(2**14), 2**14 + 1, size=(N, 24), dtype=np.int64)           # dense: every bit-plane about half full
fc = np.where(v < 0, np.uint64(LB.scheme.Q) - (-v).astype(np.uint64), v.astype(np.uint64)).astype(np.uint64)
fc_dev = eng.to_device(fc)
cm = eng.new_commitment()
cms = torch.empty((K, KAPPA, 24), dtype=torch.int64, device="cuda")
eng.commit_ntt(eng.to_device(rng.integers(0, 2**63, size=(N, 24), dtype=np.uint64)), cm)
eng.set_profiling(True)
eng.mac_profile()
t_side = timeit(lambda: eng.decompose_commit(fc_dev, cm, cms), reps=10)
mac_ms, mac_n = eng.mac_profile()
eng.set_profiling(False)
eng.synchronize()
# the 14 commits run in Toom-3 form (20 wide multiplies per Fq3 product) as two launches, 12 + 2 planes; timeit made 13 calls
calls = 13
mac_call_ms = mac_ms / calls
wide = 14 * KAPPA * N * 8 * 20
out["fold_side"] = {"what": "decompose_witness + commit_witnesses for one side: 15 planes, 14 matrix commits, y_0",
                    "ms": t_side, "mac_kernels_ms_per_call": mac_call_ms, "mac_launches_per_call": mac_n / calls,
                    "commitments_per_s": 14 / (t_side * 1e-3), "ring_elems_per_s": 14 * N / (t_side * 1e-3),
                    "imad_wide_executed_per_s": wide / (mac_call_ms * 1e-3), "imad_wide_peak_per_s": 9.154e12,
                    "imad_pipe_frac": wide / (mac_call_ms * 1e-3) / 9.154e12,
                    "karatsuba_equiv_frac": 14 * KAPPA * N * 8 * 24 / (mac_call_ms * 1e-3) / 9.154e12}
print(out["fold_side"], flush=True)
# pack + planes kernels alone (no matrix commits): the decomposition's own cost
t_planes = timeit(lambda: L.lat_ajtai_decompose_commit_dev(scheme._h, fc_dev.data_ptr(), N, None, None, None, None), reps=10)
out["planes_only"] = {"what": "pack_coeff + planes_fx_kernel (15 planes, extended layout in Toom-3 form, 569 MB written)", "ms": t_planes,
                      "GBps_written": K * N * 48 * 8 / t_planes / 1e6}
print(out["planes_only"], flush=True)
# 28-witness batch through commit_ntt_batch (both sides in one launch)
fs = torch.from_numpy(rng.integers(0, 2**63, size=(28, N, 24), dtype=np.int64)).cuda()
cms28 = torch.empty((28, KAPPA, 24), dtype=torch.int64, device="cuda")
eng.set_profiling(True)
eng.mac_profile()
t28 = timeit(lambda: eng.commit_ntt(fs, cms28), reps=5)
mac_ms, mac_n = eng.mac_profile()
wide28 = 28 * KAPPA * N * 8 * 20  # Toom-3, one launch of 7 groups x 4 witnesses; timeit made 8 calls
out["batch28"] = {"ms": t28, "mac_kernels_ms_per_call": mac_ms / 8, "mac_launches_per_call": mac_n / 8,
                  "imad_pipe_frac": wide28 / (mac_ms / 8 * 1e-3) / 9.154e12}
print(out["batch28"], flush=True)
# fold of the 2K resident planes (compute_f_0 + iCRT), both sides filled by decompose_commit
scheme_h = scheme._h
for side in (0, 1):
    assert L.lat_ajtai_select_side(scheme_h, side) == 0
    eng.decompose_commit(fc_dev, cm, cms)
rho = eng.to_device(rng.integers(0, 2**63, size=(2 * K, 24), dtype=np.uint64))
f0 = torch.empty((N, 24), dtype=torch.int64, device="cuda")
f0c = torch.empty_like(f0)
t_fold = timeit(lambda: L.lat_ajtai_fold_witness_dev(scheme_h, rho.data_ptr(), f0.data_ptr(), f0c.data_ptr()), reps=10)
out["fold_witness"] = {"what": "compute_f_0 over 2K = 30 resident planes + iCRT (n = 98 815)", "ms": t_fold,
                       "bytes_read": 2 * K * N * 48 * 8, "GBps": 2 * K * N * 48 * 8 / t_fold / 1e6}
print(out["fold_witness"], flush=True)
del fs, cms28
scheme.close()

# ---- configs[2] shape on one GPU: kappa = 32, n = 2^20 (A = 6.44 GB), single commit ---------------------------------------
N20 = 1 << 20
big = LB.AjtaiCommitmentScheme(KAPPA, N20)
row = rng.integers(0, 2**63, size=(1, N20, 24), dtype=np.uint64)
for i in range(KAPPA):
    big.upload_rows(i, row)  # the same random row 32 times: timing only
eng2 = DeviceScheme(big)
f20 = torch.from_numpy(rng.integers(0, 2**63, size=(N20, 24), dtype=np.int64)).cuda()
cm20 = eng2.new_commitment()
eng2.set_profiling(True)
eng2.mac_profile()
t20 = timeit(lambda: eng2.commit_ntt(f20, cm20), reps=10)
mac_ms, mac_n = eng2.mac_profile()
out["commit_n_2_20"] = {"what": "commit_ntt, kappa = 32, n = 2^20, one GPU (A = 6.44 GB streamed)", "ms": t20,
                        "mac_kernel_ms": mac_ms / mac_n, "GBps": KAPPA * N20 * 192 / (mac_ms / mac_n) / 1e6,
                        "hbm_frac": KAPPA * N20 * 192 / (mac_ms / mac_n) / 1e6 / 6548.8}
print(out["commit_n_2_20"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.environ.get("EXTRA_OUT", "gpurun_out/extra.json"), "w"), indent=1)
