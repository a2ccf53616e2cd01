"""Pins the Python oracle against every known-answer vector the reference's tests hold for the path
(tests/golden/reference_kats.json) and mirrors the reference's algebraic property tests."""
import json
import os
import random

import pytest

from oracle import lattice_oracle as O

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
Q = O.Q


def test_constants():
    assert KATS["modulus"]["q"] == Q == 2**64 - 2**32 + 1
    assert KATS["nonresidue"]["value"] == O.NONRESIDUE
    # GOLD/ntt.rs:449-467 (test_roots_of_unity) and GOLD/mod.rs:194-206
    roots = KATS["roots_of_unity_24"]["values"]
    assert roots == O.W
    assert len(set(roots)) == 24 and all(pow(r, 24, Q) == 1 for r in roots)
    assert all(pow(O.NONRESIDUE, i, Q) != 1 for i in range(1, 24))
    assert KATS["kappa"]["value"] == O.KAPPA
    assert O.KAPPA * (2 * O.W[4] - 1) % Q == 1  # the literal is the INVERSE of 2*zeta-1
    assert KATS["eight_inv"]["value"] == O.EIGHT_INV == pow(8, Q - 2, Q)
    assert KATS["four_inv"]["value"] == O.FOUR_INV == pow(4, Q - 2, Q)
    # every twiddle is a power of two
    assert all(O.W[i] == pow(2, 8 * (5 * i % 24), Q) for i in range(24))
    assert O.EIGHT_INV == pow(2, 189, Q) and O.FOUR_INV == pow(2, 190, Q)


@pytest.mark.parametrize("kat", KATS["crt_pre_homogenize"], ids=lambda k: k["name"])
def test_crt_icrt_kats(kat):
    # GOLD/ntt.rs:563-787: expected arrays are the pre-homogenize layout
    coeffs, slots = kat["coeffs"], kat["slots_dehomogenized"]
    assert O.crt_raw(coeffs) == slots
    got = O.crt(coeffs)
    O.dehomogenize(got)
    assert got == slots
    ev = list(slots)
    O.homogenize(ev)
    assert O.icrt(ev) == [c % Q for c in coeffs]


def test_homogenize_inverse_and_cube():
    # GOLD/ntt.rs:471-563 (inverses, squares, extension equations)
    rng = random.Random(0)
    x = [rng.randrange(Q) for _ in range(24)]
    y = list(x)
    O.homogenize(y)
    O.dehomogenize(y)
    assert x == y
    exps = [1, 13, 7, 19, 5, 17, 11, 23]
    for s in range(8):
        e = [0] * 24
        e[3 * s + 1] = 1  # the image of X in slot s before homogenize
        e2 = [0] * 24
        e2[3 * s + 2] = 1
        O.homogenize(e)
        O.homogenize(e2)
        xs, x2 = e[3 * s : 3 * s + 3], e2[3 * s : 3 * s + 3]
        assert O.fq3_mul(xs, xs) == x2
        assert O.fq3_mul(O.fq3_mul(xs, xs), xs) == [O.W[exps[s]], 0, 0]


def test_crt_icrt_roundtrip():
    # GOLD/ntt.rs:789-806 and RING/cyclotomic_ring/crt.rs:85-147 (sized down)
    rng = random.Random(1)
    for _ in range(300):
        c = [rng.randrange(Q) for _ in range(24)]
        assert O.icrt(O.crt(c)) == c
        assert O.crt(O.icrt(c)) == c
    one = O.ntt_from_scalar(1)
    assert O.icrt(one) == [1] + [0] * 23  # GOLD/mod.rs:185-190 (test_icrt_one)


def test_mul_crt():
    # GOLD/mod.rs:231-247: crt(a)*crt(b) == crt(a*b mod X^24-X^12+1)
    rng = random.Random(2)
    for _ in range(20):
        a = [rng.randrange(Q) for _ in range(24)]
        b = [rng.randrange(Q) for _ in range(24)]
        assert O.icrt(O.ntt_mul(O.crt(a), O.crt(b))) == O.poly_mul(a, b)


def test_commit_ntt_closed_form():
    # LF/commitment/commitment_scheme.rs:150-185, on a reduced n to stay fast in pure Python; the full
    # kappa=9, n=2^15 case is checked by the C oracle (tests/test_oracle_c.py).
    kappa, n = 9, 256
    A = [[O.ntt_from_scalar(i * n + j) for j in range(n)] for i in range(kappa)]
    w = [O.ntt_from_scalar(2)] * n
    cm = O.commit(A, w)
    for i in range(kappa):
        assert cm[i] == O.ntt_from_scalar(n * (2 * i * n + (n - 1)))
    with pytest.raises(O.WrongWitnessLength):
        O.commit(A, w[:-1])


def test_gadget_decompose_kat():
    # RING/balanced_decomposition/mod.rs:469-514
    k = KATS["gadget_decompose_pm15"]
    vec = [[c % Q] * 24 for c in k["input_coeff"]]
    got = O.gadget_decompose(vec, k["b"], k["padding"])
    exp = [[d % Q] * 24 for row in k["expected_digits"] for d in row]
    assert got == exp
    back = O.gadget_recompose(got, k["b"], k["padding"], ntt_form=False)
    assert back == vec


def test_decompose_balanced_property():
    # RING/balanced_decomposition/mod.rs:405-423 (range sized down by stride), plus negatives and the
    # tie rule |rem| == b/2 is kept (mod.rs:79).
    for b in KATS["decompose_property"]["bases"]:
        for v in list(range(0, 65537, 7)) + [Q - x for x in range(1, 3000, 13)]:
            d = O.decompose_balanced(v, b, 32)
            assert all(abs(O.signed_rep(x)) <= b // 2 for x in d)
            acc = 0
            for x in reversed(d):
                acc = (acc * b + x) % Q
            assert acc == v % Q
    assert [O.signed_rep(x) for x in O.decompose_balanced(2**14, 2**15, 5)] == [2**14, 0, 0, 0, 0]
    assert [O.signed_rep(x) for x in O.decompose_balanced(2**14 + 1, 2**15, 5)] == [-(2**14) + 1, 1, 0, 0, 0]
    assert [O.signed_rep(x) for x in O.decompose_balanced(Q - 2**14, 2**15, 5)] == [-(2**14), 0, 0, 0, 0]
    with pytest.raises(O.DigitOverflow):
        O.decompose_balanced(2**15, 2, 15)
    assert [O.signed_rep(x) for x in O.decompose_balanced(Q - (2**15 - 1), 2, 15)] == [-1] * 15


def test_witness_roundtrips():
    # LF/arith.rs:516-548
    rng = random.Random(3)
    B, L = 2**15, 5
    w = [[rng.randrange(Q) for _ in range(24)] for _ in range(6)]
    f_coeff, f = O.witness_from_w_ccs(w, B, L)
    assert len(f) == 30
    assert all(abs(O.signed_rep(c)) <= B // 2 for e in f_coeff for c in e)
    fc2, w2 = O.witness_from_f(f, B, L)
    assert fc2 == f_coeff and w2 == w
    f3, w3 = O.witness_from_f_coeff(f_coeff, B, L)
    assert f3 == f and w3 == w


def test_commit_witnesses_homomorphic():
    # LF/nifs/decomposition/tests/mod.rs:203-236: homomorphic y_0 == committing every plane directly
    rng = random.Random(4)
    B, L, K, kappa, wit_len = 2**15, 5, 15, 4, 4
    n = wit_len * L
    A = [[[rng.randrange(Q) for _ in range(24)] for _ in range(n)] for _ in range(kappa)]
    w = [[rng.randrange(Q) for _ in range(24)] for _ in range(wit_len)]
    f_coeff, f = O.witness_from_w_ccs(w, B, L)
    cm = O.commit(A, f)
    planes, planes_f = O.decompose_witness(f_coeff, 2, K)
    assert len(planes) == K and all(len(p) == n for p in planes)
    assert all(O.signed_rep(c) in (-1, 0, 1) for p in planes for e in p for c in e)
    ys = O.commit_witnesses(A, planes_f, cm, 2)
    assert ys == [O.commit(A, pf) for pf in planes_f]
    # recomposed commitment equals cm (tests/mod.rs:340-369)
    acc = [[0] * 24 for _ in range(kappa)]
    for y in reversed(ys):
        acc = O.commitment_add(O.commitment_scale(acc, O.ntt_from_scalar(2)), y)
    assert acc == cm


def test_get_fhat_layout():
    # LF/arith.rs:455-502: fhat[j][i] slots carry coefficients 8j..8j+7 of f_i as base-field scalars
    f_coeff = [[(100 * i + t) for t in range(24)] for i in range(3)]
    fh = O.get_fhat(f_coeff)
    assert len(fh) == 3
    for j in range(3):
        for i in range(3):
            assert [fh[j][i][3 * t] for t in range(8)] == f_coeff[i][8 * j : 8 * j + 8]
            assert all(fh[j][i][3 * t + 1] == 0 == fh[j][i][3 * t + 2] for t in range(8))


def test_mont_helpers():
    assert O.from_mont(O.to_mont(12345)) == 12345
    assert O.MONT_R == 2**32 - 1
    assert O.MONT_RINV == pow(2, 128, Q)
