"""GPU parity tests: the CUDA engine, called through its C ABI (ctypes -> liblattice_ajtai.so), against the C oracle
on identical seeded inputs, against the reference's known-answer vectors, and -- at the zkVM's full sizes -- through
size-independent properties.  Bit-exact everywhere: this is integer arithmetic mod q (no tolerance)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import latticeum_b200 as LB
from latticeum_b200 import _capi as capi
from latticeum_b200 import scheme as S
from oracle import c_oracle as CO

Q = S.Q
KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
DP = LB.GoldiLocksDP


def ptr(a):
    return a.ctypes.data


def gpu_crt(x, inverse=False):
    x = np.ascontiguousarray(x, dtype=np.uint64)
    out = np.empty_like(x)
    fn = capi.lib().lat_ring_icrt if inverse else capi.lib().lat_ring_crt
    assert fn(ptr(x), x.size // 24, ptr(out), 0) == 0, capi.last_error()
    return out


def make_scheme(A, mont=False, params=DP):
    A = np.ascontiguousarray(A, dtype=np.uint64)
    return LB.AjtaiCommitmentScheme.new(S.to_mont(A) if mont else A, params=params, mont=mont)


def maybe_mont(x, mont):
    return S.to_mont(x) if mont else x


def unmont(x, mont):
    return CO.from_mont(x) if mont else x


# ---- CRT / iCRT ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kat", KATS["crt_pre_homogenize"], ids=lambda k: k["name"])
def test_crt_icrt_reference_kats(kat):
    # GOLD/ntt.rs:563-787; expected arrays are in the pre-homogenize layout
    coeffs = np.array([c % Q for c in kat["coeffs"]], dtype=np.uint64).reshape(1, 24)
    slots = np.array(kat["slots_dehomogenized"], dtype=np.uint64)
    got = gpu_crt(coeffs)[0]
    assert CO.dehomogenize(got).tolist() == slots.tolist()
    assert gpu_crt(CO.homogenize(slots).reshape(1, 24), inverse=True)[0].tolist() == coeffs[0].tolist()


@pytest.mark.parametrize("count", [1, 2, 127, 128, 129, 1000, 4097])
def test_crt_icrt_vs_oracle_ragged(count):
    x = CO.fill_uniform((count, 24), 100 + count)
    assert np.array_equal(gpu_crt(x), CO.crt(x))
    assert np.array_equal(gpu_crt(x, inverse=True), CO.icrt(x))


def test_crt_edge_values_and_empty():
    edge = np.array([[0] * 24, [Q - 1] * 24, [1] * 24, [Q // 2] * 24, [Q // 2 + 1] * 24, [2**32 - 1] * 24, [2**32] * 24],
                    dtype=np.uint64)
    assert np.array_equal(gpu_crt(edge), CO.crt(edge))
    assert np.array_equal(gpu_crt(edge, inverse=True), CO.icrt(edge))
    empty = np.empty((0, 24), np.uint64)
    assert capi.lib().lat_ring_crt(None, 0, None, 0) == 0
    assert gpu_crt(empty).shape == (0, 24)


def test_crt_icrt_roundtrip_large():
    # GOLD/ntt.rs:789-806 (CRT . iCRT = id, 10^6 times) -- here 2^20 elements in one batch, both orders
    x = CO.fill_uniform((1 << 20, 24), 7)
    y = gpu_crt(x)
    assert np.array_equal(gpu_crt(y, inverse=True), x)
    assert np.array_equal(gpu_crt(gpu_crt(x, inverse=True)), x)
    # and spot-check the big batch against the oracle
    idx = np.arange(0, 1 << 20, 4099)
    assert np.array_equal(y[idx], CO.crt(x[idx]))


def test_mul_crt_property():
    # GOLD/mod.rs:231-247: crt(a) * crt(b) = crt(a*b); checked with the scalar monomial b = X (a rotation)
    a = CO.fill_uniform((64, 24), 8)
    # a * X mod X^24 - X^12 + 1: shift up, X^24 = X^12 - 1
    ax = np.zeros_like(a)
    ax[:, 1:] = a[:, :-1]
    top = a[:, 23].astype(object)
    ax[:, 12] = ((ax[:, 12].astype(object) + top) % Q).astype(np.uint64)
    ax[:, 0] = ((-top) % Q).astype(np.uint64)
    x_poly = np.zeros((1, 24), np.uint64)
    x_poly[0, 1] = 1
    cx = gpu_crt(x_poly)[0]
    ca = gpu_crt(a)
    prod = S._fq3_mul_scalar_vec(ca, cx, False)
    assert np.array_equal(prod, gpu_crt(ax))


# ---- commit --------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mont", [False, True], ids=["canonical", "montgomery"])
def test_commit_ntt_reference_closed_form(mont):
    # LF/commitment/commitment_scheme.rs:150-185 at its real size: kappa = 9, n = 2^15, A_ij = i*n + j, witness = 2
    k = KATS["commit_ntt_closed_form"]
    kappa, n = k["kappa"], k["n"]
    A = np.zeros((kappa, n, 24), np.uint64)
    idx = np.arange(kappa, dtype=np.uint64)[:, None] * np.uint64(n) + np.arange(n, dtype=np.uint64)[None, :]
    A[:, :, 0::3] = idx[:, :, None]
    scheme = make_scheme(A, mont)
    w = np.tile(S.ntt_from_scalar(2, mont), (n, 1))
    cm = scheme.commit_ntt(w)
    for i in range(kappa):
        exp = S.ntt_from_scalar(n * (2 * i * n + (n - 1)), mont)
        assert cm.as_ref()[i].tolist() == exp.tolist()
    with pytest.raises(LB.WrongWitnessLength) as e:  # commitment_scheme.rs:64-69
        scheme.commit_ntt(w[:-1])
    assert (e.value.got, e.value.expected) == (n - 1, n)
    scheme.close()


@pytest.mark.parametrize("kappa,n", [(1, 1), (3, 2), (4, 5), (9, 63), (17, 64), (20, 257), (32, 1000), (33, 130), (64, 77)])
@pytest.mark.parametrize("mont", [False, True], ids=["canonical", "montgomery"])
def test_commit_vs_oracle_ragged(kappa, n, mont):
    A = CO.fill_uniform((kappa, n, 24), 1000 + kappa)
    f = CO.fill_uniform((n, 24), 2000 + n)
    scheme = make_scheme(A, mont)
    assert scheme.kappa() == kappa and scheme.width() == n
    cm = scheme.commit(maybe_mont(f, mont))
    assert np.array_equal(unmont(cm.as_ref(), mont), CO.commit(A, f))
    # all-zero witness -> zero commitment (initialize_accumulator, ZKVM/main.rs:316-330)
    assert not scheme.commit(np.zeros((n, 24), np.uint64)).as_ref().any()
    scheme.close()


def test_degenerate_rand_matrix_all_rows_equal():
    # AjtaiCommitmentScheme::rand is vec![vec![R::rand(rng); n]; kappa] (LF/commitment/commitment_scheme.rs:56-58): ONE sampled
    # ring element in every entry, so every row of every commitment equals a * sum_j f_j.  The engine must not
    # special-case it -- and must get it right.
    kappa, n = 32, 4099
    scheme = LB.AjtaiCommitmentScheme.rand(kappa, n, seed=9)
    a = S._uniform((1, 24), 9)[0]
    f = CO.fill_uniform((n, 24), 201)
    cm = scheme.commit_ntt(f).as_ref()
    total = np.zeros(24, dtype=object)
    for row in f.astype(object):
        total = (total + row) % Q
    exp = S._fq3_mul_scalar_vec(total.astype(np.uint64).reshape(1, 24), a, False)[0]
    assert all(np.array_equal(cm[i], exp) for i in range(kappa))
    scheme.close()


@pytest.mark.parametrize("seed", range(6))
def test_commit_random_shapes(seed):
    # random (kappa, n) including every row-group count of the matrix layout, both representations, batch of 1..3
    rng = np.random.default_rng(1000 + seed)
    kappa = int(rng.integers(1, 41))
    n = int(rng.integers(1, 3000))
    count = int(rng.integers(1, 4))
    mont = bool(rng.integers(0, 2))
    A = CO.fill_uniform((kappa, n, 24), 2000 + seed)
    fs = CO.fill_uniform((count, n, 24), 3000 + seed)
    scheme = make_scheme(A, mont)
    got = scheme.commit_ntt_batch(maybe_mont(fs, mont))
    for p in range(count):
        assert np.array_equal(unmont(got[p].as_ref(), mont), CO.commit(A, fs[p])), (kappa, n, count, mont, p)
    scheme.close()


def test_commit_extreme_values():
    # every operand q-1: the lazy accumulators see the largest possible products
    kappa, n = 32, 4096
    A = np.full((kappa, n, 24), Q - 1, dtype=np.uint64)
    f = np.full((n, 24), Q - 1, dtype=np.uint64)
    scheme = make_scheme(A)
    assert np.array_equal(scheme.commit(f).as_ref(), CO.commit(A, f))
    scheme.close()


def test_commit_batch_and_errors():
    kappa, n = 8, 300
    A = CO.fill_uniform((kappa, n, 24), 5)
    scheme = make_scheme(A)
    for count in (1, 2, 3, 14):
        fs = CO.fill_uniform((count, n, 24), 60 + count)
        cms = scheme.commit_ntt_batch(fs)
        for k in range(count):
            assert np.array_equal(cms[k].as_ref(), CO.commit(A, fs[k]))
    with pytest.raises(LB.WrongWitnessLength):
        scheme.commit_ntt_batch(CO.fill_uniform((2, n + 1, 24), 1))
    scheme.close()
    # a matrix with missing rows refuses to commit
    s2 = LB.AjtaiCommitmentScheme(kappa, n)
    s2.upload_rows(0, A[:3])
    with pytest.raises(LB.EngineError):
        s2.commit(CO.fill_uniform((n, 24), 1))
    s2.upload_rows(3, A[3:])
    assert np.array_equal(s2.commit(fs[0]).as_ref(), CO.commit(A, fs[0]))
    s2.close()


@pytest.mark.parametrize("kappa", [1, 8, 9, 16, 17, 24, 28, 32, 33])
@pytest.mark.parametrize("mont", [False, True])
def test_commit_batch_toom_form_every_split(kappa, mont):
    # launches with several witnesses run in Toom-3 form on the 5-word matrix (csrc/goldilocks.cuh ToomAcc): 4, 2 or 1
    # witnesses per thread, counts above 4 split into a multiple of 4 and the rest -- every split shape, every row-group
    # geometry (kappa 1 .. 33), a ragged last tile, both representations
    n = 131 if kappa != 32 else 517
    A = CO.fill_uniform((kappa, n, 24), 300 + kappa)
    scheme = make_scheme(A, mont)
    for count in (2, 3, 4, 5, 6, 7, 8, 9, 12, 15):
        fs = CO.fill_uniform((count, n, 24), 900 + count)
        cms = scheme.commit_ntt_batch(maybe_mont(fs, mont))
        for k in range(count):
            assert np.array_equal(unmont(cms[k].as_ref(), mont), CO.commit(A, fs[k])), (kappa, count, k)
    scheme.close()


def test_commit_batch_toom_form_extreme_values_and_reupload():
    # q-1 everywhere: the largest point values (a(2) = 7 (q-1) mod q, a(-1) = q-1) and the largest lazy sums; then a row
    # is replaced and the 5-word copy of the matrix must follow
    kappa, n, count = 32, 2048, 6
    A = np.full((kappa, n, 24), Q - 1, dtype=np.uint64)
    fs = np.full((count, n, 24), Q - 1, dtype=np.uint64)
    fs[1] = 0
    fs[2, ::2] = 1
    scheme = make_scheme(A)
    cms = scheme.commit_ntt_batch(fs)
    for k in range(count):
        assert np.array_equal(cms[k].as_ref(), CO.commit(A, fs[k])), k
    A2 = A.copy()
    A2[5] = CO.fill_uniform((n, 24), 77)
    A2[31] = 0
    scheme.upload_rows(5, A2[5:6])
    scheme.upload_rows(31, A2[31:32])
    assert np.array_equal(scheme.commit(fs[0]).as_ref(), CO.commit(A2, fs[0]))  # single witness: the 3-word matrix
    cms = scheme.commit_ntt_batch(fs)
    for k in range(count):
        assert np.array_equal(cms[k].as_ref(), CO.commit(A2, fs[k])), k
    scheme.close()


def test_upload_column_shard_with_stride():
    # a column shard of a wider host matrix: row_stride = full width
    kappa, n_total, lo, hi = 5, 100, 30, 71
    A = CO.fill_uniform((kappa, n_total, 24), 11)
    f = CO.fill_uniform((n_total, 24), 12)
    s = LB.AjtaiCommitmentScheme(kappa, hi - lo)
    L = capi.lib()
    assert L.lat_ajtai_upload_rows(s._h, 0, kappa, A.ctypes.data + lo * 24 * 8, n_total) == 0
    got = s.commit(f[lo:hi]).as_ref()
    assert np.array_equal(got, CO.commit(np.ascontiguousarray(A[:, lo:hi]), f[lo:hi]))
    s.close()


def test_commit_coeff_and_decompose_and_commit():
    # LF/commitment/commitment_scheme.rs:107-139
    kappa, wl = 6, 40
    n = wl * DP.L
    A = CO.fill_uniform((kappa, n, 24), 21)
    scheme = make_scheme(A)
    fcoef = CO.fill_uniform((n, 24), 22)
    assert np.array_equal(scheme.commit_coeff(fcoef).as_ref(), CO.commit(A, CO.crt(fcoef)))
    w = CO.fill_uniform((wl, 24), 23)
    f_coeff, f = CO.witness_from_w_ccs(w, DP.B, DP.L)
    exp = CO.commit(A, f)
    assert np.array_equal(scheme.decompose_and_commit_ntt(w).as_ref(), exp)
    assert np.array_equal(scheme.decompose_and_commit_coeff(CO.icrt(w)).as_ref(), exp)
    with pytest.raises(LB.WrongWitnessLength):
        scheme.decompose_and_commit_ntt(w[:-1])
    scheme.close()


# ---- Witness::from_w_ccs ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mont", [False, True], ids=["canonical", "montgomery"])
@pytest.mark.parametrize("kappa,wl", [(4, 4), (9, 129), (32, 1000)])
def test_witness_from_w_ccs_vs_oracle(kappa, wl, mont):
    n = wl * DP.L
    A = CO.fill_uniform((kappa, n, 24), 31)
    w = CO.fill_uniform((wl, 24), 32)
    # mix in scalar-embedded elements, zeros and tie-rule coefficients
    w[0] = CO.scalar_elem(12345)
    w[1] = 0
    tie = np.zeros((1, 24), np.uint64)
    tie[0, :4] = [2**14, 2**14 + 1, Q - 2**14, Q - 2**14 - 1]
    w[2] = CO.crt(tie)[0]
    scheme = make_scheme(A, mont)
    wit, cm = LB.Witness.from_w_ccs(scheme, maybe_mont(w, mont), commit=True)
    f_coeff, f = CO.witness_from_w_ccs(w, DP.B, DP.L)
    assert np.array_equal(unmont(wit.f_coeff, mont), f_coeff)
    assert np.array_equal(unmont(wit.f, mont), f)
    assert np.array_equal(unmont(cm.as_ref(), mont), CO.commit(A, f))
    assert wit.commit(scheme) == cm
    # outputs are optional
    w2 = LB.Witness.from_w_ccs(scheme, maybe_mont(w, mont), want_f=False, want_f_coeff=False)
    assert w2.f is None and w2.f_coeff is None
    scheme.close()


# ---- decompose_witness + commit_witnesses -------------------------------------------------------------------------------
def signed_to_fq(v):
    v = np.asarray(v, dtype=np.int64)
    return np.where(v < 0, np.uint64(Q) - (-v).astype(np.uint64), v.astype(np.uint64)).astype(np.uint64)


def fq_to_signed(x):
    x = np.asarray(x, dtype=np.uint64)
    return np.where(x > np.uint64(Q // 2), -((np.uint64(Q) - x).astype(np.int64)), x.astype(np.int64))


def small_coeffs(n, seed, bound):
    rng = np.random.default_rng(seed)
    v = rng.integers(-bound, bound + 1, size=(n, 24), dtype=np.int64)
    return signed_to_fq(v), v


@pytest.mark.parametrize("mont", [False, True], ids=["canonical", "montgomery"])
@pytest.mark.parametrize("kappa,n", [(4, 20), (9, 333), (32, 2000)])
def test_decompose_commit_vs_oracle(kappa, n, mont):
    A = CO.fill_uniform((kappa, n, 24), 41)
    fc, _ = small_coeffs(n, 42, 2**15 - 1)
    fc[0, :3] = [2**15 - 1, Q - (2**15 - 1), 0]
    f = CO.crt(fc)
    cm = CO.commit(A, f)
    pc, pf, cms = CO.decompose_commit(A, fc, cm, 2, DP.K)
    scheme = make_scheme(A, mont)
    wit_s, ys = LB.LFDecompositionProver.decompose_and_commit(
        scheme, maybe_mont(fc, mont), LB.Commitment(maybe_mont(cm, mont), mont))
    assert len(wit_s) == DP.K == len(ys)
    for k in range(DP.K):
        assert np.array_equal(unmont(wit_s[k].f_coeff, mont), pc[k]), k
        assert np.array_equal(unmont(wit_s[k].f, mont), pf[k]), k
        assert np.array_equal(unmont(ys[k].as_ref(), mont), cms[k]), k
    # the reference's own self-consistency test (LF/nifs/decomposition/tests/mod.rs:203-236):
    # homomorphic y_0 == committing plane 0 directly
    assert wit_s[0].commit(scheme) == ys[0]
    # commit_witnesses on materialised witnesses gives the same
    ys2 = LB.LFDecompositionProver.commit_witnesses(scheme, wit_s, LB.Commitment(maybe_mont(cm, mont), mont))
    assert ys2 == ys
    # decompose_witness alone
    wit_s2 = LB.LFDecompositionProver.decompose_witness(scheme, LB.Witness(None, None, maybe_mont(fc, mont), mont))
    assert all(np.array_equal(a.f, b.f) for a, b in zip(wit_s, wit_s2))
    scheme.close()


def test_decompose_commit_digit_overflow_is_reported():
    # |c| = 2^15 needs 16 binary digits: the reference panics (mod.rs:80), the engine returns LAT_E_DIGIT_OVERFLOW
    kappa, n = 4, 50
    A = CO.fill_uniform((kappa, n, 24), 51)
    scheme = make_scheme(A)
    for bad in (2**15, Q - 2**15, 2**40, Q // 2):
        fc, _ = small_coeffs(n, 52, 100)
        fc[n - 1, 23] = bad
        with pytest.raises(LB.DigitOverflow):
            LB.LFDecompositionProver.decompose_and_commit(scheme, fc, LB.Commitment.zeroed(kappa))
    # and the engine stays usable afterwards
    fc, _ = small_coeffs(n, 53, 2**15 - 1)
    cm = CO.commit(A, CO.crt(fc))
    _, ys = LB.LFDecompositionProver.decompose_and_commit(scheme, fc, LB.Commitment(cm))
    assert np.array_equal(ys[3].as_ref(), CO.decompose_commit(A, fc, cm, 2, DP.K)[2][3])
    scheme.close()


def test_resident_witness_path():
    # from_w_ccs leaves the limbs on the device; decompose_commit_resident reuses them without re-upload
    kappa, wl = 8, 60
    n = wl * DP.L
    A = CO.fill_uniform((kappa, n, 24), 61)
    w = CO.fill_uniform((wl, 24), 62)
    scheme = make_scheme(A)
    wit, cm = LB.Witness.from_w_ccs(scheme, w, commit=True)
    cms = np.empty((DP.K, kappa, 24), np.uint64)
    st = capi.lib().lat_ajtai_decompose_commit_resident(scheme._h, cm.as_ref().ctypes.data, None, None, cms.ctypes.data)
    assert st == 0, capi.last_error()
    exp = CO.decompose_commit(A, wit.f_coeff, cm.as_ref(), 2, DP.K, want_planes=False)[2]
    assert np.array_equal(cms, exp)
    scheme.close()


# ---- pipelined steps (submit / wait) ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("mont", [False, True], ids=["canonical", "montgomery"])
@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
def test_commit_pipeline_matches_blocking_calls(mont, pinned):
    kappa, wl, steps = 9, 300, 11
    n = wl * DP.L
    A = CO.fill_uniform((kappa, n, 24), 81)
    scheme = make_scheme(A, mont)
    pipe = LB.CommitPipeline(scheme)
    ws = []
    for k in range(steps):
        w = maybe_mont(CO.fill_uniform((wl, 24), 800 + k), mont)
        if pinned:
            buf = LB.pinned_empty((wl, 24))
            buf[:] = w
            w = buf
        ws.append(w)
    got, tickets = {}, []
    for k in range(steps):
        if len(tickets) == pipe.depth:  # keep the pipeline full: wait for the oldest, submit the next
            t = tickets.pop(0)
            got[t] = pipe.wait(t)
        tickets.append(pipe.submit(ws[k]))
    for t in tickets:
        got[t] = pipe.wait(t)
    assert sorted(got) == list(range(steps))
    for k in range(steps):
        _, f = CO.witness_from_w_ccs(unmont(np.array(ws[k]), mont), DP.B, DP.L)
        assert np.array_equal(unmont(got[k].as_ref(), mont), CO.commit(A, f)), f"step {k}"
    # the last submitted step is the handle's resident witness
    wit, cm = LB.Witness.from_w_ccs(scheme, ws[-1], commit=True)
    assert cm == got[steps - 1]
    scheme.close()


def test_commit_pipeline_errors():
    # two limbs of 2^15 only, so that a coefficient can overflow the padding at all (2^75 > q with the zkVM's L = 5)
    kappa, wl = 4, 40
    params = LB.DecompositionParams(B=1 << 15, L=2, B_SMALL=2, K=15)
    n = wl * params.L
    A = CO.fill_uniform((kappa, n, 24), 83)
    scheme = make_scheme(A, params=params)
    pipe = LB.CommitPipeline(scheme)
    small = np.zeros((wl, 24), np.uint64)
    small[:, :] = CO.fill_uniform((wl, 24), 84) % np.uint64(1 << 28)
    w = CO.crt(small)
    with pytest.raises(LB.WrongWitnessLength):
        pipe.submit(w[:-1])
    tickets = [pipe.submit(w) for _ in range(pipe.depth)]
    with pytest.raises(LB.EngineError):  # full
        pipe.submit(w)
    assert capi.lib().lat_ajtai_wait(scheme._h, 10_000) == capi.LAT_E_INVALID_ARGUMENT  # unknown ticket
    # a digit overflow is reported by the wait of the step that caused it, and only that one
    bad = w.copy()
    bad[3] = CO.crt(np.full((1, 24), 1 << 31, np.uint64))[0]  # 2^31 needs three limbs of 2^15
    first = pipe.wait(tickets.pop(0))
    tb = pipe.submit(bad)
    for t in tickets:
        assert pipe.wait(t) == first
    with pytest.raises(LB.DigitOverflow):
        pipe.wait(tb)
    assert pipe.wait(pipe.submit(w)) == first
    scheme.close()


@pytest.mark.parametrize("log2_B,L,K", [(8, 8, 9), (12, 6, 13), (15, 5, 15), (4, 3, 5), (1, 2, 2)])
@pytest.mark.parametrize("mont", [False, True], ids=["canonical", "montgomery"])
def test_other_decomposition_parameters(log2_B, L, K, mont):
    # DecompositionParams other than the zkVM's (LF/decomposition_parameters.rs:11-20): the kernels take B, L, K at run time
    params = LB.DecompositionParams(B=1 << log2_B, L=L, B_SMALL=2, K=K)
    kappa, wl = 6, 211
    n = wl * L
    A = CO.fill_uniform((kappa, n, 24), 300 + log2_B)
    scheme = make_scheme(A, mont, params)
    # w_ccs whose coefficients fit in L digits of base B: |c| <= (B/2) * (B^L - 1) / (B - 1)
    bound = (1 << (log2_B - 1)) * ((1 << (log2_B * L)) - 1) // ((1 << log2_B) - 1) if log2_B > 1 else (1 << L) - 1
    bound = min(bound, 2**62)
    rng = np.random.default_rng(log2_B * 100 + L)
    coeff = rng.integers(-bound, bound + 1, size=(wl, 24), dtype=np.int64)
    coeff[0], coeff[1] = bound, -bound
    w = CO.crt(signed_to_fq(coeff))
    wit, cm = LB.Witness.from_w_ccs(scheme, maybe_mont(w, mont), commit=True)
    f_coeff, f = CO.witness_from_w_ccs(w, params.B, L)
    assert np.array_equal(unmont(wit.f_coeff, mont), f_coeff) and np.array_equal(unmont(wit.f, mont), f)
    e_cm = CO.commit(A, f)
    assert np.array_equal(unmont(cm.as_ref(), mont), e_cm)
    if log2_B <= K:  # the limbs then fit the K bit planes: decompose_witness + commit_witnesses on the resident witness
        _, ys = LB.LFDecompositionProver.decompose_and_commit(scheme, maybe_mont(f_coeff, mont), cm, want_planes=False, side=1)
        _, pf1, e_ys = CO.decompose_commit(A, f_coeff, e_cm, 2, K)
        assert all(np.array_equal(unmont(ys[k].as_ref(), mont), e_ys[k]) for k in range(K))
        # and the fold step on top: an accumulator with |c| < 2^(K-1), short challenges
        acc = rng.integers(-(1 << (K - 1)) + 1, 1 << (K - 1), size=(n, 24), dtype=np.int64)
        acc_fc = signed_to_fq(acc // 8 if K > 4 else acc)
        acc_cm = CO.commit(A, CO.crt(acc_fc))
        fs = LB.FoldStep(scheme)
        fs.set_accumulator(maybe_mont(acc_fc, mont), LB.Commitment(maybe_mont(acc_cm, mont), mont))
        cm2, ys0, ys1, d16 = fs.begin(maybe_mont(w, mont))
        _, pf0, e_ys0 = CO.decompose_commit(A, acc_fc, acc_cm, 2, K)
        assert np.array_equal(unmont(cm2.as_ref(), mont), e_cm) and np.array_equal(d16.astype(np.int64), fq_to_signed(f_coeff))
        assert all(np.array_equal(unmont(ys0[k].as_ref(), mont), e_ys0[k]) and np.array_equal(unmont(ys1[k].as_ref(), mont), e_ys[k])
                   for k in range(K))
        rho = CO.crt(signed_to_fq(rng.integers(-1, 2, size=(2 * K, 24))))  # tiny challenges keep f_0 inside 2^K for small K too
        e_f0 = CO.compute_f0(rho, [pf0[k] for k in range(K)] + [pf1[k] for k in range(K)])
        if np.abs(fq_to_signed(CO.icrt(e_f0))).max() < (1 << K):
            cm0, f0d, f0, _ = fs.finish(maybe_mont(rho, mont), want_f0=True)
            assert np.array_equal(unmont(f0, mont), e_f0) and np.array_equal(f0d.astype(np.int64), fq_to_signed(CO.icrt(e_f0)))
            assert np.array_equal(unmont(cm0.as_ref(), mont), CO.commit(A, e_f0))
        else:
            with pytest.raises(LB.DigitOverflow):
                fs.finish(maybe_mont(rho, mont))
    scheme.close()


def test_step_overlap_gives_identical_commitments():
    """lat_ajtai_set_step_overlap: consecutive device-resident steps whose kernels overlap (next witness kernel
    under the draining matrix-vector kernel, alternating witness buffers) commit exactly what serialised steps do,
    also when other entry points are interleaved."""
    import torch
    from latticeum_b200.device import DeviceScheme

    kappa, wl, steps = 32, 4000, 12
    n = wl * DP.L
    A = CO.fill_uniform((kappa, n, 24), 85)
    scheme = make_scheme(A)
    eng = DeviceScheme(scheme)
    ws = [eng.to_device(CO.fill_uniform((wl, 24), 860 + k)) for k in range(steps)]
    exp = []
    for k in range(steps):
        cm = eng.new_commitment()
        eng.witness_commit(ws[k], cm)
        exp.append(cm.clone())
    torch.cuda.synchronize()
    eng.set_step_overlap(True)
    outs = [eng.new_commitment() for _ in range(steps)]
    for rep in range(3):
        for k in range(steps):
            eng.witness_commit(ws[k], outs[k])
            if rep == 1 and k % 4 == 3:  # an unrelated call between two steps: f of a random vector, batch of 2
                f = eng.to_device(CO.fill_uniform((2, n, 24), 900 + k))
                eng.commit_ntt(f, eng.new_commitment(2))
        torch.cuda.synchronize()
        for k in range(steps):
            assert torch.equal(outs[k], exp[k]), (rep, k)
    _, f = CO.witness_from_w_ccs(ws[3].cpu().numpy().view(np.uint64), DP.B, DP.L)
    assert np.array_equal(exp[3].cpu().numpy().view(np.uint64), CO.commit(A, f))
    eng.set_step_overlap(False)
    scheme.close()


# ---- the zkVM's full size: oracle on a bounded part + size-independent properties -----------------------------------------
@pytest.fixture(scope="module")
def zkvm():
    A = CO.fill_uniform((LB.KAPPA, LB.N, 24), 1)
    scheme = LB.AjtaiCommitmentScheme.new(A)
    yield A, scheme
    scheme.close()


def steady_state_w(seed):
    """SURVEY 8d mix (i): ~5.5 % scalar-embedded elements, the rest dense uniform CRT-form elements."""
    w = CO.fill_uniform((LB.W_SIZE, 24), seed)
    nsc = 1088
    vals = CO.fill_uniform((nsc,), seed + 1)
    w[:nsc] = 0
    w[:nsc, 0::3] = vals[:, None]
    return w


def test_zkvm_step_commit_vs_oracle(zkvm):
    A, scheme = zkvm
    w = steady_state_w(70)
    wit, cm = LB.Witness.from_w_ccs(scheme, w, commit=True)
    f_coeff, f = CO.witness_from_w_ccs(w, DP.B, DP.L)
    assert np.array_equal(wit.f_coeff, f_coeff) and np.array_equal(wit.f, f)
    assert np.array_equal(cm.as_ref(), CO.commit(A, f))
    # first-step mix (ii): scalars only; and all-zero (iii)
    w2 = w.copy()
    w2[1088:] = 0
    _, cm2 = LB.Witness.from_w_ccs(scheme, w2, commit=True, want_f=False, want_f_coeff=False)
    assert np.array_equal(cm2.as_ref(), CO.commit(A, CO.witness_from_w_ccs(w2, DP.B, DP.L)[1]))
    _, cm0 = LB.Witness.from_w_ccs(scheme, np.zeros_like(w), commit=True, want_f=False, want_f_coeff=False)
    assert not cm0.as_ref().any()


def test_zkvm_fold_step_properties(zkvm):
    A, scheme = zkvm
    w = steady_state_w(80)
    wit, cm = LB.Witness.from_w_ccs(scheme, w, commit=True)
    wit_s, ys = LB.LFDecompositionProver.decompose_and_commit(scheme, wit.f_coeff, cm)
    # (1) homomorphic y_0 equals the direct commitment of plane 0        LF/nifs/decomposition/tests/mod.rs:203-236
    assert wit_s[0].commit(scheme) == ys[0]
    # (2) recomposed commitment equals cm                                 LF/nifs/decomposition/tests/mod.rs:340-369
    two = S.ntt_from_scalar(2)
    acc = LB.Commitment.zeroed(LB.KAPPA)
    for y in reversed(ys):
        acc = acc * two + y
    assert acc == cm
    # (3) planes recompose to f_coeff (digits in {-1,0,1})                LF/nifs/decomposition/utils.rs:84-195
    rec = np.zeros(wit.f_coeff.shape, dtype=np.int64)
    for k in range(DP.K):
        d = fq_to_signed(wit_s[k].f_coeff)
        assert set(np.unique(d).tolist()) <= {-1, 0, 1}
        rec += d << k
    assert np.array_equal(rec, fq_to_signed(wit.f_coeff))
    # (4) two planes against the oracle
    for k in (1, 14):
        assert np.array_equal(ys[k].as_ref(), CO.commit(A, CO.crt(wit_s[k].f_coeff)))
    # (5) linearity of the commitment at full size
    f2 = CO.fill_uniform((LB.N, 24), 81)
    s = ((wit.f.astype(object) + f2.astype(object)) % Q).astype(np.uint64)
    assert scheme.commit(s) == scheme.commit(wit.f) + scheme.commit(f2)


def test_zkvm_fold_both_sides_realistic_and_f0(zkvm):
    """SURVEY 8d case 2, realistic: right side = limbs of a steady-state step witness, left side = a folded witness
    (rounded Gaussian, sigma ~ 350, clipped to +-(2^15 - 1): mostly-empty high planes).  Both sides decomposed and
    committed at the zkVM's full size, then folded (compute_f_0 + iCRT) against the oracle."""
    A, scheme = zkvm
    rng = np.random.default_rng(5)
    left = np.clip(np.rint(rng.normal(0.0, 350.0, size=(LB.N, 24))), -(2**15 - 1), 2**15 - 1).astype(np.int64)
    left[0, :4] = [2**15 - 1, -(2**15 - 1), 2**14, -(2**14)]
    fc_left = signed_to_fq(left)
    fc_right, _ = CO.witness_from_w_ccs(steady_state_w(82), DP.B, DP.L)
    planes = []
    for side, fc in ((0, fc_left), (1, fc_right)):
        cm = scheme.commit_coeff(fc)
        _, ys = LB.LFDecompositionProver.decompose_and_commit(scheme, fc, cm, want_planes=False, side=side)
        pl = CO.decompose_planes(fc, 2, DP.K)
        for k in (1, 9, 14):  # a full plane, a sparse one, the (almost) empty top one
            assert np.array_equal(ys[k].as_ref(), CO.commit(A, CO.crt(pl[k]))), (side, k)
        acc = LB.Commitment.zeroed(LB.KAPPA)
        for y in reversed(ys):
            acc = acc * S.ntt_from_scalar(2) + y
        assert acc == cm
        planes += [CO.crt(pl[k]) for k in range(DP.K)]
    rho = CO.crt(signed_to_fq(rng.integers(-32, 32, size=(2 * DP.K, 24))))  # short challenges, CYC/rings/goldilocks.rs:32-35
    wit0 = LB.LFFoldingProver.compute_f_0(scheme, rho)
    exp = CO.compute_f0(rho, planes)
    assert np.array_equal(wit0.f, exp)
    assert np.array_equal(wit0.f_coeff, CO.icrt(exp))


def test_zkvm_fold_step_begin_finish_full_size(zkvm):
    """The GPU side of one IVC step's fold at the zkVM's full size (kappa = 32, n = 98 815; zk_latticefold.rs:37-102) through
    lat_ajtai_fold_step_begin / _finish, two dependent steps, against the oracle and the reference's own properties."""
    A, scheme = zkvm
    rng = np.random.default_rng(11)
    acc_fc = signed_to_fq(np.clip(np.rint(rng.normal(0.0, 350.0, size=(LB.N, 24))), -(2**15 - 1), 2**15 - 1).astype(np.int64))
    acc_cm = CO.commit(A, CO.crt(acc_fc))
    fs = LB.FoldStep(scheme)
    fs.set_accumulator(acc_fc, LB.Commitment(acc_cm))
    two = S.ntt_from_scalar(2)
    for step in range(2):
        w = steady_state_w(90 + step)
        rho = CO.crt(signed_to_fq(rng.integers(-32, 32, size=(2 * DP.K, 24))))
        cm, ys_acc, ys_step, d16 = fs.begin(w)
        f_coeff, f = CO.witness_from_w_ccs(w, DP.B, DP.L)
        assert np.array_equal(cm.as_ref(), CO.commit(A, f))
        assert np.array_equal(d16.astype(np.int64), fq_to_signed(f_coeff))
        # recomposed commitments equal the undecomposed ones, on both sides     LF/nifs/decomposition/tests/mod.rs:340-369
        for ys, whole in ((ys_acc, acc_cm), (ys_step, cm.as_ref())):
            acc = LB.Commitment.zeroed(LB.KAPPA)
            for y in reversed(ys):
                acc = acc * two + y
            assert np.array_equal(acc.as_ref(), whole)
        _, pf0, e_acc = CO.decompose_commit(A, acc_fc, acc_cm, 2, DP.K)
        _, pf1, e_step = CO.decompose_commit(A, f_coeff, cm.as_ref(), 2, DP.K)
        for k in (0, 1, 9, 14):
            assert np.array_equal(ys_acc[k].as_ref(), e_acc[k]) and np.array_equal(ys_step[k].as_ref(), e_step[k]), (step, k)
        cm0, f0d, f0, w0 = fs.finish(rho, want_f0=True, want_w_ccs=True)
        e_f0 = CO.compute_f0(rho, [pf0[k] for k in range(DP.K)] + [pf1[k] for k in range(DP.K)])
        e_f0c = CO.icrt(e_f0)
        assert np.array_equal(f0, e_f0) and np.array_equal(f0d.astype(np.int64), fq_to_signed(e_f0c))
        assert np.array_equal(w0, CO.gadget_recompose_ntt(e_f0, DP.B, DP.L))
        assert scheme.commit(f0) == cm0  # the folded commitment is the commitment of the folded witness
        acc_fc, acc_cm = e_f0c, cm0.as_ref().copy()


def test_sharded_config_shape_n_2_20():
    """BASELINE configs[2] / SURVEY 8d case 4 on one GPU: kappa = 32, n = 2^20 (A = 6.44 GB), uniform f in CRT form.
    Rows are generated, uploaded and checked one at a time so that the host never holds the matrix."""
    kappa, n = 32, 1 << 20
    scheme = LB.AjtaiCommitmentScheme(kappa, n)
    f = CO.fill_uniform((n, 24), 7)
    exp = np.empty((kappa, 24), np.uint64)
    for i in range(kappa):
        row = CO.fill_uniform((1, n, 24), 1000 + i)
        scheme.upload_rows(i, row)
        exp[i] = CO.commit(row, f)[0]
    cm = scheme.commit_ntt(f)
    assert np.array_equal(cm.as_ref(), exp)
    # linearity at this size
    g = CO.fill_uniform((n, 24), 8)
    s = ((f.astype(object) + g.astype(object)) % Q).astype(np.uint64)
    assert scheme.commit_ntt(s) == cm + scheme.commit_ntt(g)
    with pytest.raises(LB.WrongWitnessLength):
        scheme.commit_ntt(f[:-1])
    scheme.close()


@pytest.mark.parametrize("mont", [False, True], ids=["canonical", "montgomery"])
def test_fold_witness_compute_f0_vs_oracle(mont):
    # LF/nifs/folding.rs:258-268 (compute_f_0) + LF/arith.rs:299-313 (Witness::from_f) on the resident planes
    kappa, n = 8, 1500
    A = CO.fill_uniform((kappa, n, 24), 91)
    scheme = make_scheme(A, mont)
    sides, planes = [], []
    for side in (0, 1):
        fc, _ = small_coeffs(n, 92 + side, 2**15 - 1)
        cm = CO.commit(A, CO.crt(fc))
        LB.LFDecompositionProver.decompose_and_commit(scheme, maybe_mont(fc, mont), LB.Commitment(maybe_mont(cm, mont), mont),
                                                      want_planes=False, side=side)
        planes += list(CO.decompose_commit(A, fc, cm, 2, DP.K)[1])
    rho = CO.fill_uniform((2 * DP.K, 24), 95)
    wit = LB.LFFoldingProver.compute_f_0(scheme, maybe_mont(rho, mont))
    exp_f0 = CO.compute_f0(rho, planes)
    assert np.array_equal(unmont(wit.f, mont), exp_f0)
    assert np.array_equal(unmont(wit.f_coeff, mont), CO.icrt(exp_f0))
    scheme.close()
    # needs both sides
    s2 = make_scheme(A)
    with pytest.raises(LB.EngineError):
        LB.LFFoldingProver.compute_f_0(s2, rho)
    s2.close()


def oracle_fold_step(A, w, acc_fc, acc_cm, rho):
    """One IVC step's GPU-side work on the oracle: step commit, both decompositions, compute_f_0, from_f, cm_0."""
    f_coeff, f = CO.witness_from_w_ccs(w, DP.B, DP.L)
    cm = CO.commit(A, f)
    _, pf0, ys0 = CO.decompose_commit(A, acc_fc, acc_cm, 2, DP.K)
    _, pf1, ys1 = CO.decompose_commit(A, f_coeff, cm, 2, DP.K)
    f0 = CO.compute_f0(rho, [pf0[k] for k in range(DP.K)] + [pf1[k] for k in range(DP.K)])
    cm0 = np.zeros((A.shape[0], 24), dtype=object)
    for r, y in zip(rho, list(ys0) + list(ys1)):  # cm_0 = sum rho_i * cm_i   LF/nifs/folding/utils.rs:466-472
        cm0 = (cm0 + S._fq3_mul_scalar_vec(y, r, False).astype(object)) % Q
    return f_coeff, cm, ys0, ys1, f0, CO.icrt(f0), cm0.astype(np.uint64)


@pytest.mark.parametrize("mont", [False, True], ids=["canonical", "montgomery"])
def test_fold_step_begin_finish_vs_oracle(mont):
    # zk_latticefold.rs:37-102 as two blocking calls; three dependent steps: f_0 of step i is the accumulator of step i+1
    kappa, wl = 8, 301
    n = wl * DP.L
    A = CO.fill_uniform((kappa, n, 24), 160)
    scheme = make_scheme(A, mont)
    fs = LB.FoldStep(scheme)
    rng = np.random.default_rng(161)
    acc_fc = signed_to_fq(np.clip(np.rint(rng.normal(0.0, 350.0, size=(n, 24))), -(2**15 - 1), 2**15 - 1).astype(np.int64))
    acc_cm = CO.commit(A, CO.crt(acc_fc))
    with pytest.raises(LB.EngineError):
        fs.begin(CO.fill_uniform((wl, 24), 1))  # no accumulator yet
    fs.set_accumulator(maybe_mont(acc_fc, mont), LB.Commitment(maybe_mont(acc_cm, mont), mont))
    for step in range(3):
        w = CO.fill_uniform((wl, 24), 170 + step)
        rho = CO.crt(signed_to_fq(rng.integers(-32, 32, size=(2 * DP.K, 24))))  # short challenges
        e_fc, e_cm, e_ys0, e_ys1, e_f0, e_f0c, e_cm0 = oracle_fold_step(A, w, acc_fc, acc_cm, rho)
        # every other step hands the accumulator commitment over explicitly, the others use the resident one
        cm_arg = LB.Commitment(maybe_mont(acc_cm, mont), mont) if step % 2 == 0 else None
        cm, ys0, ys1, d16 = fs.begin(maybe_mont(w, mont), cm_arg)
        assert np.array_equal(unmont(cm.as_ref(), mont), e_cm)
        assert np.array_equal(fq_to_signed(e_fc), d16.astype(np.int64))
        assert np.array_equal(unmont(S.digits_to_fq(d16, mont), mont), e_fc)
        for k in range(DP.K):
            assert np.array_equal(unmont(ys0[k].as_ref(), mont), e_ys0[k]), (step, 0, k)
            assert np.array_equal(unmont(ys1[k].as_ref(), mont), e_ys1[k]), (step, 1, k)
        cm0, f0d, f0, w0 = fs.finish(maybe_mont(rho, mont), want_f0=True, want_w_ccs=True)
        assert np.array_equal(unmont(f0, mont), e_f0)
        assert np.array_equal(f0d.astype(np.int64), fq_to_signed(e_f0c))
        assert np.array_equal(unmont(cm0.as_ref(), mont), e_cm0)
        assert np.array_equal(unmont(w0, mont), CO.gadget_recompose_ntt(e_f0, DP.B, DP.L))
        # homomorphism: the folded commitment IS the commitment of the folded witness
        assert np.array_equal(e_cm0, CO.commit(A, e_f0))
        acc_fc, acc_cm = e_f0c, e_cm0
    # get_fhat on the device and from the digits on the host (LF/arith.rs:273-297)
    import torch
    fh = torch.empty((3, n, 24), dtype=torch.int64, device="cuda")
    assert capi.lib().lat_ajtai_get_fhat_dev(scheme._h, 1, fh.data_ptr()) == 0, capi.last_error()
    assert capi.lib().lat_ajtai_synchronize(scheme._h) == 0
    exp_fhat = S.get_fhat(maybe_mont(acc_fc, mont))
    assert np.array_equal(fh.cpu().numpy().view(np.uint64), exp_fhat)
    assert np.array_equal(S.get_fhat_from_digits(f0d, mont), exp_fhat)
    # a folded witness that breaks the norm bound 2^K is reported, not truncated; the accumulator is then invalid
    cm, ys0, ys1, _ = fs.begin(maybe_mont(CO.fill_uniform((wl, 24), 180), mont))
    with pytest.raises(LB.DigitOverflow):
        fs.finish(maybe_mont(CO.fill_uniform((2 * DP.K, 24), 181), mont))  # uniform (not short) challenges
    with pytest.raises(LB.EngineError):
        fs.begin(maybe_mont(CO.fill_uniform((wl, 24), 182), mont))
    scheme.close()


def test_fold_step_finish_in_element_ranges_vs_oracle():
    # from n = 4096 on, lat_ajtai_fold_step_finish folds, inverts and packs f_0 in four ranges of elements and sends each
    # range's digits down while the next one folds; n = 5105 makes every range boundary ragged (not a multiple of 4, of
    # the 32 elements of a fold block, or of the 64 of a planes block)
    kappa, wl, mont = 4, 1021, True
    n = wl * DP.L
    A = CO.fill_uniform((kappa, n, 24), 260)
    scheme = make_scheme(A, mont)
    fs = LB.FoldStep(scheme)
    rng = np.random.default_rng(261)
    acc_fc = signed_to_fq(np.clip(np.rint(rng.normal(0.0, 350.0, size=(n, 24))), -(2**15 - 1), 2**15 - 1).astype(np.int64))
    acc_cm = CO.commit(A, CO.crt(acc_fc))
    fs.set_accumulator(maybe_mont(acc_fc, mont), LB.Commitment(maybe_mont(acc_cm, mont), mont))
    for step in range(2):
        w = CO.fill_uniform((wl, 24), 270 + step)
        rho = CO.crt(signed_to_fq(rng.integers(-32, 32, size=(2 * DP.K, 24))))
        e_fc, e_cm, e_ys0, e_ys1, e_f0, e_f0c, e_cm0 = oracle_fold_step(A, w, acc_fc, acc_cm, rho)
        cm, ys0, ys1, d16 = fs.begin(maybe_mont(w, mont))
        assert np.array_equal(unmont(cm.as_ref(), mont), e_cm)
        assert np.array_equal(fq_to_signed(e_fc), d16.astype(np.int64))
        for k in range(DP.K):
            assert np.array_equal(unmont(ys0[k].as_ref(), mont), e_ys0[k]), (step, 0, k)
            assert np.array_equal(unmont(ys1[k].as_ref(), mont), e_ys1[k]), (step, 1, k)
        cm0, f0d, f0, w0 = fs.finish(maybe_mont(rho, mont), want_f0=True, want_w_ccs=True)
        assert np.array_equal(unmont(f0, mont), e_f0)
        assert np.array_equal(f0d.astype(np.int64), fq_to_signed(e_f0c))
        assert np.array_equal(unmont(cm0.as_ref(), mont), e_cm0)
        assert np.array_equal(unmont(w0, mont), CO.gadget_recompose_ntt(e_f0, DP.B, DP.L))
        acc_fc, acc_cm = e_f0c, e_cm0
    scheme.close()


def test_witness_from_w_ccs_compact_vs_oracle():
    kappa, wl = 9, 777
    A = CO.fill_uniform((kappa, wl * DP.L, 24), 190)
    w = CO.fill_uniform((wl, 24), 191)
    w[:5] = 0
    f_coeff, f = CO.witness_from_w_ccs(w, DP.B, DP.L)
    for mont in (False, True):
        scheme = make_scheme(A, mont)
        wit, cm = LB.Witness.from_w_ccs_compact(scheme, maybe_mont(w, mont), want_f=True)
        assert np.array_equal(wit.digits.astype(np.int64), fq_to_signed(f_coeff))
        assert np.array_equal(unmont(wit.f_coeff, mont), f_coeff) and np.array_equal(unmont(wit.f, mont), f)
        assert np.array_equal(unmont(cm.as_ref(), mont), CO.commit(A, f))
        assert np.array_equal(S.get_fhat_from_digits(wit.digits, mont), S.get_fhat(maybe_mont(f_coeff, mont)))
        scheme.close()


@pytest.mark.parametrize("log2_b,L", [(15, 5), (8, 8), (7, 10), (2, 32), (1, 32), (15, 1)])
def test_gadget_decompose_vs_oracle(log2_b, L):
    # RING/balanced_decomposition/mod.rs:163-175 for &[R]; L <= 8 goes through the kernel's shared-memory tile, L > 8
    # through the direct path; values are drawn so that exactly L balanced limbs suffice
    rng = np.random.default_rng(300 + 10 * log2_b + L)
    b = 1 << log2_b
    bound = min((b // 2) * (b**L - 1) // (b - 1), (Q - 1) // 2)  # largest magnitude with L balanced limbs
    mag = np.array([[int(rng.integers(0, bound + 1)) if bound < 2**63 else int(rng.integers(0, 2**63)) % (bound + 1)
                     for _ in range(24)] for _ in range(77)], dtype=object)
    mag[0, :] = 0
    mag[1, :] = bound
    mag[2, :3] = [min(t, bound) for t in (b // 2, b // 2 + 1, max(b // 2 - 1, 0))]  # the tie rule: |rem| == b/2 is kept
    sign = rng.integers(0, 2, size=mag.shape).astype(bool)
    v = np.array([[(Q - int(m)) % Q if s else int(m) for m, s in zip(rm, rs)] for rm, rs in zip(mag, sign)], dtype=np.uint64)
    exp = CO.gadget_decompose(v, b, L)
    assert np.array_equal(LB.gadget_decompose(v, log2_b, L), exp)
    assert np.array_equal(unmont(LB.gadget_decompose(S.to_mont(v), log2_b, L, mont=True), True), exp)
    if bound < (Q - 1) // 2:
        v2 = v.copy()
        v2[5, 7] = np.uint64(bound + 1)
        with pytest.raises(LB.DigitOverflow):
            LB.gadget_decompose(v2, log2_b, L)


def test_device_pointer_entry_points_for_matrix_and_transforms():
    """lat_ajtai_upload_rows_dev (matrix rows already on the device, with a row stride) and lat_ring_crt_dev /
    lat_ring_icrt_dev (both kernel paths: below and above the one-thread-per-element threshold, in place and out of
    place) against the host-buffer entry points and the oracle."""
    import ctypes as C
    import torch

    L = capi.lib()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
    kappa, n, stride = 5, 333, 400
    wide = CO.fill_uniform((kappa, stride, 24), 310)  # a column block [0, n) of a wider matrix
    f = CO.fill_uniform((n, 24), 311)
    for mont in (False, True):
        scheme = LB.AjtaiCommitmentScheme(kappa, n, mont=mont)
        dev = torch.from_numpy(maybe_mont(wide, mont).view(np.int64)).cuda()
        assert L.lat_ajtai_upload_rows_dev(scheme._h, 0, 2, dev.data_ptr(), stride) == 0, capi.last_error()
        assert L.lat_ajtai_upload_rows_dev(scheme._h, 2, kappa - 2, dev[2:].data_ptr(), stride) == 0, capi.last_error()
        assert np.array_equal(unmont(scheme.commit_ntt(maybe_mont(f, mont)).as_ref(), mont), CO.commit(wide[:, :n], f))
        assert L.lat_ajtai_upload_rows_dev(scheme._h, 4, 2, dev.data_ptr(), stride) == capi.LAT_E_WRONG_MATRIX_DIMENSIONS
        assert L.lat_ajtai_upload_rows_dev(scheme._h, 0, 1, dev.data_ptr(), n - 1) == capi.LAT_E_WRONG_MATRIX_DIMENSIONS
        scheme.close()
    for count in (37, (1 << 14) + 5):
        x = CO.fill_uniform((count, 24), 312 + count % 7)
        xd = torch.from_numpy(x.view(np.int64)).cuda()
        yd = torch.empty_like(xd)
        assert L.lat_ring_crt_dev(xd.data_ptr(), count, yd.data_ptr(), stream) == 0
        torch.cuda.synchronize()
        fwd = yd.cpu().numpy().view(np.uint64)
        assert np.array_equal(fwd, CO.crt(x)) and np.array_equal(fwd, gpu_crt(x))
        assert L.lat_ring_icrt_dev(yd.data_ptr(), count, yd.data_ptr(), stream) == 0  # in place
        torch.cuda.synchronize()
        assert np.array_equal(yd.cpu().numpy().view(np.uint64), x)
        assert L.lat_ring_icrt_dev(xd.data_ptr(), count, yd.data_ptr(), stream) == 0
        torch.cuda.synchronize()
        assert np.array_equal(yd.cpu().numpy().view(np.uint64), CO.icrt(x))
    assert L.lat_ring_crt_dev(None, 0, None, stream) == 0
    assert L.lat_ring_crt_dev(None, 3, None, stream) == capi.LAT_E_INVALID_ARGUMENT


def test_null_and_zero_arguments_never_crash_and_leave_the_handle_usable():
    """Every lat_ajtai_* entry point called on a live handle with NULL pointers and zero sizes returns one of the
    documented status codes; afterwards the handle still commits correctly."""
    import ctypes as C

    kappa, wl = 3, 40
    n = wl * DP.L
    A = CO.fill_uniform((kappa, n, 24), 320)
    scheme = make_scheme(A)
    L = capi.lib()
    known = {capi.LAT_OK, capi.LAT_E_INVALID_ARGUMENT, capi.LAT_E_WRONG_WITNESS_LENGTH, capi.LAT_E_WRONG_MATRIX_DIMENSIONS,
             capi.LAT_E_CUDA, capi.LAT_E_DIGIT_OVERFLOW, capi.LAT_E_MATRIX_INCOMPLETE, capi.LAT_E_WRONG_COMMITMENT_LENGTH}
    skip = {"lat_ajtai_create", "lat_ajtai_destroy", "lat_ajtai_kappa", "lat_ajtai_width", "lat_ajtai_set_stream",
            "lat_ajtai_wait"}  # set_stream(NULL) is meaningful (own stream); wait(0) has nothing to wait for
    for name, (res, args) in capi.SIGNATURES.items():
        if not name.startswith("lat_ajtai_") or name in skip:
            continue
        zero = [None if (a is C.c_void_p or hasattr(a, "contents")) else 0 for a in args]
        zero[0] = scheme._h
        got = getattr(L, name)(*zero)
        assert got in known, (name, got)
    assert L.lat_ajtai_synchronize(scheme._h) in known
    L.lat_ajtai_set_step_overlap(scheme._h, 0)
    L.lat_ajtai_set_profiling(scheme._h, 0)
    f = CO.fill_uniform((n, 24), 321)
    assert np.array_equal(scheme.commit_ntt(f).as_ref(), CO.commit(A, f))
    w = CO.fill_uniform((wl, 24), 322)
    f_coeff, fw = CO.witness_from_w_ccs(w, DP.B, DP.L)
    wit, cm = LB.Witness.from_w_ccs(scheme, w, commit=True)
    assert np.array_equal(wit.f_coeff, f_coeff) and np.array_equal(cm.as_ref(), CO.commit(A, fw))
    scheme.close()


def test_two_host_threads_with_their_own_handles():
    """Two host threads, each with its own engine handle on the same GPU, plus the handle-less transform calls (which
    share one per-device scratch behind a lock), all running at once (ctypes drops the GIL during the calls)."""
    import threading

    kappa, wl = 4, 64
    n = wl * DP.L
    errors = []

    def worker(seed):
        try:
            A = CO.fill_uniform((kappa, n, 24), seed)
            scheme = make_scheme(A, mont=bool(seed & 1))
            mont = bool(seed & 1)
            for it in range(12):
                w = CO.fill_uniform((wl, 24), seed * 100 + it)
                f_coeff, f = CO.witness_from_w_ccs(w, DP.B, DP.L)
                wit, cm = LB.Witness.from_w_ccs(scheme, maybe_mont(w, mont), commit=True)
                assert np.array_equal(unmont(cm.as_ref(), mont), CO.commit(A, f)), (seed, it)
                assert np.array_equal(unmont(wit.f_coeff, mont), f_coeff), (seed, it)
                x = CO.fill_uniform((50 + it, 24), seed * 1000 + it)
                assert np.array_equal(gpu_crt(gpu_crt(x), inverse=True), x)
                a = CO.fill_uniform((3, 64), seed * 2000 + it)
                assert np.array_equal(LB.ntt_negacyclic(LB.ntt_negacyclic(a), inverse=True), a)
            scheme.close()
        except Exception as e:  # noqa: BLE001
            errors.append((seed, repr(e)))

    threads = [threading.Thread(target=worker, args=(330 + i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert not any(t.is_alive() for t in threads)


def test_gadget_recompose_vs_oracle():
    # RING/balanced_decomposition/mod.rs:177-190 in CRT form: recompose(from_w_ccs(w).f) == w  (LF/arith.rs:516-548)
    w = CO.fill_uniform((333, 24), 96)
    f_coeff, f = CO.witness_from_w_ccs(w, DP.B, DP.L)
    got = LB.gadget_recompose(f)
    assert np.array_equal(got, CO.gadget_recompose_ntt(f, DP.B, DP.L))
    assert np.array_equal(got, w)


def test_commitment_sum_and_single_rank_exchange():
    # lat_commitment_sum (SURVEY 8e fold of column-shard partials) and the fused exchange kernel with world = 1
    # (multi-rank runs need real peers: tests/test_multigpu.py under torchrun)
    import ctypes as C

    import torch

    L = capi.lib()
    parts = CO.fill_uniform((5, 7 * 24), 123)
    out = np.empty(7 * 24, np.uint64)
    assert L.lat_commitment_sum(parts.ctypes.data, 5, 7 * 24, out.ctypes.data, 0) == 0, capi.last_error()
    assert out.tolist() == [int(v) % Q for v in parts.astype(object).sum(axis=0)]
    words = 32 * 24
    partial = torch.from_numpy(CO.fill_uniform((words,), 124).view(np.int64)).cuda()
    recv = torch.zeros(2 * words, dtype=torch.int64, device="cuda")
    flags = torch.zeros(2, dtype=torch.int64, device="cuda")
    res = torch.empty(words, dtype=torch.int64, device="cuda")
    rp, fp = (C.c_uint64 * 1)(recv.data_ptr()), (C.c_uint64 * 1)(flags.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
    for epoch in (1, 2, 3):
        st = L.lat_commitment_exchange_dev(partial.data_ptr(), words, 0, 1, rp, fp, epoch, res.data_ptr(), stream)
        assert st == 0, capi.last_error()
        torch.cuda.synchronize()
        assert torch.equal(res, partial)
    assert flags.tolist() == [2, 3]
    assert L.lat_commitment_exchange_dev(partial.data_ptr(), words, 0, 1, rp, fp, 0, res.data_ptr(), stream) == capi.LAT_E_INVALID_ARGUMENT
    # reporting form: the kernel also writes the result and then the ticket into pinned host memory
    cm_host = torch.zeros(words, dtype=torch.int64).pin_memory()
    done_host = torch.full((1,), -1, dtype=torch.int64).pin_memory()
    st = L.lat_commitment_exchange_report_dev(partial.data_ptr(), words, 0, 1, rp, fp, 4, res.data_ptr(), cm_host.data_ptr(),
                                              done_host.data_ptr(), 77, stream)
    assert st == 0, capi.last_error()
    while int(done_host[0]) != 77:
        pass
    assert torch.equal(cm_host, partial.cpu())


def test_gated_witness_call_waits_for_the_ticket():
    # lat_ajtai_witness_from_w_ccs_gated_dev: the kernel polls a device word that the uploading stream writes after the
    # data.  The copies are ENQUEUED before the call (the header's rule: a kernel must never wait for work submitted
    # after it) but held back on their stream by a few milliseconds, so the kernel really has to wait.
    import torch
    from latticeum_b200.device import DeviceScheme

    kappa, wl = 8, 500
    n = wl * DP.L
    A = CO.fill_uniform((kappa, n, 24), 130)
    w = CO.fill_uniform((wl, 24), 131)
    scheme = make_scheme(A)
    eng = DeviceScheme(scheme)
    w_pin = torch.from_numpy(w.view(np.int64)).pin_memory()
    w_dev = torch.zeros((wl, 24), dtype=torch.int64, device="cuda")
    ready = torch.full((1,), -1, dtype=torch.int64, device="cuda")
    ticket = torch.tensor([41], dtype=torch.int64).pin_memory()
    cm = eng.new_commitment()
    up = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(up):
        torch.cuda._sleep(10_000_000)  # ~5 ms on the upload stream: w_dev is still all zero when the kernel starts
        w_dev.copy_(w_pin, non_blocking=True)
        ready.copy_(ticket, non_blocking=True)
    eng.witness_commit_gated(w_dev, cm, ready, 41)
    torch.cuda.synchronize()
    _, f = CO.witness_from_w_ccs(w, DP.B, DP.L)
    assert np.array_equal(cm.cpu().numpy().view(np.uint64), CO.commit(A, f))
    scheme.close()


def test_device_side_waits_are_bounded():
    # Every in-kernel wait has a %globaltimer deadline (lat::SpinGuard): a peer that never delivers, or an upload ticket
    # that never lands, ends in LAT_E_CUDA with a message -- not in a hung GPU.
    import ctypes as C

    import torch
    from latticeum_b200.device import DeviceScheme

    L = capi.lib()
    code = C.c_uint64(0)
    assert L.lat_device_wait_status(0, C.byref(code)) == 0 and code.value == 0
    assert L.lat_set_spin_timeout_ms(50) == 0
    try:
        # (1) a 2-rank exchange in which rank 1 never shows up
        words = 32 * 24
        partial = torch.from_numpy(CO.fill_uniform((words,), 124).view(np.int64)).cuda()
        recv = torch.zeros(2 * 2 * words, dtype=torch.int64, device="cuda")
        flags = torch.zeros(2 * 2, dtype=torch.int64, device="cuda")
        res = torch.empty(words, dtype=torch.int64, device="cuda")
        rp = (C.c_uint64 * 2)(recv.data_ptr(), recv.data_ptr())  # "peer" mailboxes alias the local one: nobody raises flag 1
        fp = (C.c_uint64 * 2)(flags.data_ptr(), flags.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
        assert L.lat_commitment_exchange_dev(partial.data_ptr(), words, 0, 2, rp, fp, 1, res.data_ptr(), stream) == 0
        torch.cuda.synchronize()  # returns: the wait gave up after ~50 ms
        assert L.lat_device_wait_status(0, C.byref(code)) == capi.LAT_E_CUDA
        assert code.value & 0xFF == 2 and (code.value >> 8) & 0xFF == 1, hex(code.value)  # peer flag, rank 1
        assert "peer rank 1" in capi.last_error()
        assert L.lat_device_wait_status(0, C.byref(code)) == 0 and code.value == 0  # cleared
        # (2) a gated witness call whose ticket never arrives: the handle's synchronize reports it
        kappa, wl = 4, 64
        A = CO.fill_uniform((kappa, wl * DP.L, 24), 140)
        scheme = make_scheme(A)
        eng = DeviceScheme(scheme)
        w_dev = torch.zeros((wl, 24), dtype=torch.int64, device="cuda")
        ready = torch.full((1,), -1, dtype=torch.int64, device="cuda")
        eng.witness_commit_gated(w_dev, eng.new_commitment(), ready, 7)
        with pytest.raises(S.EngineError, match="upload ticket 7"):
            eng.synchronize()
        eng.synchronize()  # cleared, and the handle is still usable
        w = CO.fill_uniform((wl, 24), 141)
        cm = eng.witness_commit(eng.to_device(w), eng.new_commitment())
        eng.synchronize()
        assert np.array_equal(DeviceScheme.to_numpy(cm), CO.commit(A, CO.witness_from_w_ccs(w, DP.B, DP.L)[1]))
        scheme.close()
    finally:
        L.lat_set_spin_timeout_ms(5000)


def test_submit_on_the_legacy_stream_behind_a_busy_stream():
    # Regression for the round-1 N=4 hang: the first submits of a handle bound to the LEGACY stream, issued while
    # that stream is still busy.  Slot state used to be memset on first use on the legacy stream (queued behind the
    # busy work) while the ticket copy ran ahead on the copy stream; the late memset then wiped the ticket and the
    # gated kernel spun for ever.  All slot state is now initialised at creation.
    import ctypes as C

    import torch

    kappa, wl = 8, 300
    A = CO.fill_uniform((kappa, wl * DP.L, 24), 150)
    w = CO.fill_uniform((wl, 24), 151)
    scheme = make_scheme(A)
    L = capi.lib()
    assert L.lat_ajtai_set_stream(scheme._h, C.c_void_p(1)) == 0  # cudaStreamLegacy
    w_pin = S.pinned_empty((wl, 24))
    w_pin[:] = w
    cms = [np.zeros((kappa, 24), np.uint64) for _ in range(capi.LAT_PIPELINE_DEPTH)]
    torch.cuda.synchronize()
    torch.cuda._sleep(40_000_000)  # ~20 ms of work on the legacy stream (torch's default stream)
    tks = []
    for cm in cms:
        tk = C.c_uint64()
        assert L.lat_ajtai_submit_w_ccs(scheme._h, w_pin.ctypes.data, wl, cm.ctypes.data, C.byref(tk)) == 0, capi.last_error()
        tks.append(tk.value)
    exp = CO.commit(A, CO.witness_from_w_ccs(w, DP.B, DP.L)[1])
    for tk, cm in zip(tks, cms):
        assert L.lat_ajtai_wait(scheme._h, tk) == 0, capi.last_error()
        assert np.array_equal(cm, exp)
    scheme.close()


# ---- standalone power-of-two negacyclic NTT (absent from the reference, parity unpinned by construction) ----
@pytest.mark.parametrize("log2_d", [1, 2, 3, 6, 8])
def test_ntt_negacyclic_vs_direct_evaluation(log2_d):
    from oracle import lattice_oracle as PO

    d = 1 << log2_d
    batch = 5 if d > 8 else 130  # 130 small polynomials: several per block and a ragged last block
    a = CO.fill_uniform((batch, d), 140 + log2_d)
    a[0] = 0
    a[1] = Q - 1
    fwd = LB.ntt_negacyclic(a)
    for p in range(min(batch, 6)):
        assert fwd[p].tolist() == PO.ntt_negacyclic([int(v) for v in a[p]]), p
    assert np.array_equal(LB.ntt_negacyclic(fwd, inverse=True), a)
    assert LB.ntt_negacyclic(a, inverse=True)[2].tolist() == PO.ntt_negacyclic([int(v) for v in a[2]], inverse=True)


@pytest.mark.parametrize("log2_d", [9, 10, 12, 13, 14])
def test_ntt_negacyclic_large_properties(log2_d):
    from oracle import lattice_oracle as PO

    d = 1 << log2_d
    batch = 3
    rng = np.random.default_rng(log2_d)
    a = CO.fill_uniform((batch, d), 150 + log2_d)
    b = CO.fill_uniform((batch, d), 160 + log2_d)
    A, B = LB.ntt_negacyclic(a), LB.ntt_negacyclic(b)
    # spot checks against the definition (O(d) each)
    for p in range(batch):
        row = [int(v) for v in a[p]]
        for i in [0, 1, d - 1] + rng.integers(0, d, size=5).tolist():
            assert int(A[p, i]) == PO.ntt_eval_at(row, int(i)), (p, i)
    # inverse o forward = id, linearity, and the convolution theorem against sparse products (X^k * a)
    assert np.array_equal(LB.ntt_negacyclic(A, inverse=True), a)
    s = ((a.astype(object) + b.astype(object)) % Q).astype(np.uint64)
    assert np.array_equal(LB.ntt_negacyclic(s).astype(object), (A.astype(object) + B.astype(object)) % Q)
    k = int(rng.integers(1, d))
    xk = np.zeros((batch, d), np.uint64)
    xk[:, k] = 1
    prod = ((LB.ntt_negacyclic(xk).astype(object) * A.astype(object)) % Q).astype(np.uint64)
    shifted = np.concatenate([(np.uint64(Q) - a[:, d - k:]) % np.uint64(Q), a[:, : d - k]], axis=1)  # X^k * a mod X^d + 1
    assert np.array_equal(LB.ntt_negacyclic(prod, inverse=True), shifted)


@pytest.mark.parametrize("log2_d", [4, 5, 7, 11])
def test_ntt_negacyclic_more_sizes_and_unaligned_buffers(log2_d):
    """Every pass structure (radix-16 passes + a 1-3 stage remainder), non-canonical inputs, and device buffers that
    are only 8-byte aligned (the kernel then uses scalar loads and stores)."""
    import ctypes as C
    import torch
    from oracle import lattice_oracle as PO

    d = 1 << log2_d
    batch = 4099 // d + 3  # a ragged last block
    a = CO.fill_uniform((batch, d), 170 + log2_d)
    a[0, :4] = [Q, 2**64 - 1, Q + 5, 0]  # representatives >= q are accepted and reduced
    ref = a % np.uint64(Q)
    fwd = LB.ntt_negacyclic(a)
    assert fwd.max() < Q
    for p in (0, 1, batch - 1):
        assert fwd[p].tolist() == PO.ntt_negacyclic([int(v) for v in ref[p]]), p
    assert np.array_equal(LB.ntt_negacyclic(fwd, inverse=True), ref)
    assert np.array_equal(LB.ntt_negacyclic(LB.ntt_negacyclic(a, inverse=True)), ref)
    L = capi.lib()
    buf_in = torch.zeros(batch * d + 1, dtype=torch.int64, device="cuda")
    buf_out = torch.zeros(batch * d + 1, dtype=torch.int64, device="cuda")
    buf_in[1:] = torch.from_numpy(a.view(np.int64).reshape(-1)).cuda()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)
    for inverse, exp in ((0, fwd), (1, LB.ntt_negacyclic(a, inverse=True))):
        assert L.lat_ntt_negacyclic_dev(buf_in.data_ptr() + 8, batch, log2_d, inverse, buf_out.data_ptr() + 8, stream) == 0
        torch.cuda.synchronize()
        assert np.array_equal(buf_out[1:].cpu().numpy().view(np.uint64).reshape(batch, d), exp)


@pytest.mark.parametrize("log2_d", [2, 6, 9, 12, 14])
def test_ntt_negacyclic_extreme_values(log2_d):
    """Inputs that sit on the lazy arithmetic's correction paths: q - 1 everywhere, 2^64 - 1 everywhere, values whose high
    word is all ones (>= q, or just below it), zeros with a single extreme entry."""
    from oracle import lattice_oracle as PO

    d = 1 << log2_d
    rng = np.random.default_rng(200 + log2_d)
    hi_ones = (np.uint64(0xFFFFFFFF) << np.uint64(32)) | rng.integers(0, 2**32, size=d, dtype=np.uint64)
    just_below = np.uint64(Q) - rng.integers(1, 2**20, size=d, dtype=np.uint64)
    spike = np.zeros(d, np.uint64)
    spike[d - 1] = np.uint64(2**64 - 1)
    a = np.stack([np.full(d, Q - 1, np.uint64), np.full(d, 2**64 - 1, np.uint64), hi_ones, just_below, spike,
                  rng.integers(0, 2**64, size=d, dtype=np.uint64)])
    ref = [[int(v) % Q for v in row] for row in a]
    fwd = LB.ntt_negacyclic(a)
    inv = LB.ntt_negacyclic(a, inverse=True)
    assert fwd.max() < Q and inv.max() < Q
    pts = [0, 1, d // 2, d - 1] + rng.integers(0, d, size=4).tolist()
    for p in range(a.shape[0]):
        for i in pts:
            assert int(fwd[p, i]) == PO.ntt_eval_at(ref[p], int(i)), (p, i)
            assert int(inv[p, i]) == PO.ntt_eval_at(ref[p], int(i), inverse=True), (p, i)
    assert np.array_equal(LB.ntt_negacyclic(fwd, inverse=True), np.array(ref, dtype=np.uint64))
    assert np.array_equal(LB.ntt_negacyclic(inv), np.array(ref, dtype=np.uint64))


def test_ntt_negacyclic_argument_checks():
    x = np.zeros((2, 8), np.uint64)
    out = np.empty_like(x)
    L = capi.lib()
    assert L.lat_ntt_negacyclic(x.ctypes.data, 2, 0, 0, out.ctypes.data, 0) == capi.LAT_E_INVALID_ARGUMENT
    assert L.lat_ntt_negacyclic(x.ctypes.data, 2, 15, 0, out.ctypes.data, 0) == capi.LAT_E_INVALID_ARGUMENT
    assert L.lat_ntt_negacyclic(x.ctypes.data, 0, 3, 0, out.ctypes.data, 0) == 0
    with pytest.raises(ValueError):
        LB.ntt_negacyclic(np.zeros((2, 12), np.uint64))


@pytest.mark.parametrize("name,marker", [("example", "example ok"), ("example_ivc", "ivc example ok")])
def test_cpp_host_mirror_example(name, marker):
    # latticeum_b200/host/ajtai.hpp, the C++ mirror of the reference API: the reference's closed-form commit test, and
    # the zkVM's folding loop (zkvm/src/main.rs:140-182) over fold_step_begin / fold_step_finish
    import subprocess

    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "latticeum_b200", "host", name)
    if not os.path.exists(exe):
        pytest.skip("example not built (run __graft_entry__.build())")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert marker in out.stdout
