"""Pins the C oracle (the CPU baseline / GPU checker) against the reference's KATs and the Python oracle."""
import json
import os
import random

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import lattice_oracle as O

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
Q = O.Q


def u64(x):
    return np.array(x, dtype=np.uint64)


def rand_elems(rng, n):
    return u64([[rng.randrange(Q) for _ in range(24)] for _ in range(n)])


def test_roots():
    assert CO.roots().tolist() == KATS["roots_of_unity_24"]["values"]


@pytest.mark.parametrize("kat", KATS["crt_pre_homogenize"], ids=lambda k: k["name"])
def test_crt_icrt_kats(kat):
    coeffs = u64([c % Q for c in kat["coeffs"]])
    slots = u64(kat["slots_dehomogenized"])
    got = CO.crt(coeffs.reshape(1, 24))[0]
    assert CO.dehomogenize(got).tolist() == slots.tolist()
    assert CO.icrt(CO.homogenize(slots).reshape(1, 24))[0].tolist() == coeffs.tolist()


def test_crt_icrt_vs_python_and_roundtrip():
    rng = random.Random(10)
    x = rand_elems(rng, 64)
    y = CO.crt(x)
    assert y.tolist() == [O.crt(e) for e in x.tolist()]
    assert CO.icrt(y).tolist() == x.tolist()
    assert CO.icrt(x).tolist() == [O.icrt(e) for e in x.tolist()]
    assert CO.crt(x, parallel=True).tolist() == y.tolist()
    assert CO.icrt(y, parallel=True).tolist() == x.tolist()
    # edge values
    edge = u64([[0] * 24, [Q - 1] * 24, [1] * 24, [Q // 2] * 24, [Q // 2 + 1] * 24])
    assert CO.crt(edge).tolist() == [O.crt(e) for e in edge.tolist()]
    assert CO.icrt(edge).tolist() == [O.icrt(e) for e in edge.tolist()]


def test_crt_icrt_many():
    # GOLD/ntt.rs:789-806 (CRT∘iCRT = id, 10^6 times in the reference; 2*10^5 here)
    x = CO.fill_uniform((200000, 24), 5)
    assert np.array_equal(CO.icrt(CO.crt(x, parallel=True), parallel=True), x)


def test_commit_closed_form_full():
    # LF/commitment/commitment_scheme.rs:150-185 at its real size: kappa=9, n=2^15
    k = KATS["commit_ntt_closed_form"]
    kappa, n = k["kappa"], k["n"]
    A = np.zeros((kappa, n, 24), np.uint64)
    idx = (np.arange(kappa, dtype=np.uint64)[:, None] * np.uint64(n) + np.arange(n, dtype=np.uint64)[None, :])
    A[:, :, 0::3] = idx[:, :, None]
    w = np.tile(CO.scalar_elem(2), (n, 1))
    cm = CO.commit(A, w)
    for i in range(kappa):
        assert cm[i].tolist() == CO.scalar_elem(n * (2 * i * n + (n - 1))).tolist()
    with pytest.raises(CO.OracleError) as e:
        CO.commit(A, w[:-1])
    assert e.value.status == CO.E_WRONG_WITNESS_LENGTH


def test_commit_vs_python():
    rng = random.Random(11)
    kappa, n = 3, 17
    A = u64([[[rng.randrange(Q) for _ in range(24)] for _ in range(n)] for _ in range(kappa)])
    f = rand_elems(rng, n)
    assert CO.commit(A, f).tolist() == O.commit(A.tolist(), f.tolist())


def test_gadget_kat_and_vs_python():
    k = KATS["gadget_decompose_pm15"]
    vec = u64([[c % Q] * 24 for c in k["input_coeff"]])
    got = CO.gadget_decompose(vec, k["b"], k["padding"])
    assert got.tolist() == [[d % Q] * 24 for row in k["expected_digits"] for d in row]
    rng = random.Random(12)
    x = rand_elems(rng, 9)
    assert CO.gadget_decompose(x, 2**15, 5).tolist() == O.gadget_decompose(x.tolist(), 2**15, 5)
    for b in (2, 4, 8, 16, 32):
        small = u64([[rng.randrange(-40000, 40000) % Q for _ in range(24)] for _ in range(5)])
        assert CO.gadget_decompose(small, b, 32).tolist() == O.gadget_decompose(small.tolist(), b, 32)
    # tie rule and overflow status
    t = u64([[2**14, 2**14 + 1, Q - 2**14, Q - 2**14 - 1] * 6])
    assert CO.gadget_decompose(t, 2**15, 5).tolist() == O.gadget_decompose(t.tolist(), 2**15, 5)
    with pytest.raises(CO.OracleError) as e:
        CO.decompose_planes(u64([[2**15] + [0] * 23]), 2, 15)
    assert e.value.status == CO.E_DIGIT_OVERFLOW
    with pytest.raises(CO.OracleError):
        CO.decompose_planes(u64([[Q - 2**15] + [0] * 23]), 2, 15)


def test_witness_and_decompose_commit_vs_python():
    rng = random.Random(13)
    B, L, K, kappa, wl = 2**15, 5, 15, 4, 4
    n = wl * L
    A = u64([[[rng.randrange(Q) for _ in range(24)] for _ in range(n)] for _ in range(kappa)])
    w = rand_elems(rng, wl)
    f_coeff, f = CO.witness_from_w_ccs(w, B, L)
    pf_coeff, pf = O.witness_from_w_ccs(w.tolist(), B, L)
    assert f_coeff.tolist() == pf_coeff and f.tolist() == pf
    assert CO.gadget_recompose_ntt(f, B, L).tolist() == w.tolist()
    cm = CO.commit(A, f)
    pc, pff, cms = CO.decompose_commit(A, f_coeff, cm, 2, K)
    planes, planes_f = O.decompose_witness(pf_coeff, 2, K)
    assert pc.tolist() == planes and pff.tolist() == planes_f
    ys = O.commit_witnesses(A.tolist(), planes_f, cm.tolist(), 2)
    assert cms.tolist() == ys
    # the reference's own self-consistency test: homomorphic y_0 == direct commit of plane 0
    assert cms[0].tolist() == CO.commit(A, pff[0]).tolist()
    rho = rand_elems(rng, 3)
    assert CO.compute_f0(rho, [pff[0], pff[1], pff[2]]).tolist() == O.compute_f_0(rho.tolist(), planes_f[:3])


def test_mont_roundtrip():
    x = CO.fill_uniform((1000,), 3)
    m = CO.to_mont(x)
    assert m[:50].tolist() == [O.to_mont(int(v)) for v in x[:50]]
    assert np.array_equal(CO.from_mont(m), x)
    assert int(x.max()) < Q
