"""The Toom-3 form of the batched matrix-vector kernels (latticeum_b200/csrc/goldilocks.cuh: toom_eval, ToomAcc) restated
in Python big integers and checked against the oracle's Fq3 product: the five evaluation points, the interpolation, the
constants the header hard-codes (1/3, (q+1)/2), and the class -> word map of the table-driven planes transform
(ring_kernels.cu: plane_word).  CPU only; the kernels themselves are checked bit for bit in tests/test_gpu_parity.py."""
import os
import random
import re

from oracle import lattice_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = O.Q


def header(name):
    return open(os.path.join(ROOT, "latticeum_b200", "csrc", name)).read()


def toom_eval(x):  # goldilocks.cuh toom_eval: values at 0, infinity, 1, -1, 2
    return [x[0], x[2], (x[0] + x[1] + x[2]) % Q, (x[0] - x[1] + x[2]) % Q, (x[0] + 2 * x[1] + 4 * x[2]) % Q]


def toom_finish(v, inv3, half):  # goldilocks.cuh ToomAcc::finish
    d0, d4, r1, rm, r2 = v
    d2 = (half((r1 + rm) % Q) - d0 - d4) % Q
    s = half((r1 - rm) % Q)
    t = half((r2 - d0 - 4 * d2 - 16 * d4) % Q)
    d3 = (t - s) % Q * inv3 % Q
    d1 = (s - d3) % Q
    nr = O.NONRESIDUE
    return [(d0 + nr * d3) % Q, (d1 + nr * d4) % Q, d2]


def test_header_constants():
    src = header("goldilocks.cuh")
    inv3 = int(re.search(r"INV3 = (0x[0-9A-Fa-f]+)ull", src).group(1), 16)
    assert 3 * inv3 % Q == 1
    half_c = int(re.search(r"\(x & 1\) \? (0x[0-9A-Fa-f]+)ull", src).group(1), 16)
    assert half_c == (Q + 1) // 2 and 2 * half_c % Q == 1


def test_toom3_accumulate_then_interpolate_equals_sum_of_fq3_products():
    src = header("goldilocks.cuh")
    inv3 = int(re.search(r"INV3 = (0x[0-9A-Fa-f]+)ull", src).group(1), 16)
    half_c = (Q + 1) // 2
    half = lambda x: (x >> 1) + (half_c if x & 1 else 0)  # noqa: E731  (canonical x)
    rnd = random.Random(7)
    edge = [0, 1, Q - 1, Q - 2, (Q - 1) // 2, 2**32 - 1, 2**32, 2**63]
    for trial in range(300):
        n = rnd.randint(1, 12)
        acc = [0] * 5
        ref = [0, 0, 0]
        for _ in range(n):
            pick = (lambda: rnd.choice(edge)) if trial % 3 == 0 else (lambda: rnd.randrange(Q))
            a, y = [pick() for _ in range(3)], [pick() for _ in range(3)]
            ref = [(r + c) % Q for r, c in zip(ref, O.fq3_mul(a, y))]
            acc = [(s + ea * ey) % Q for s, ea, ey in zip(acc, toom_eval(a), toom_eval(y))]  # five point products, summed lazily
        got = toom_finish(acc, inv3, half)
        assert all(0 <= g < Q for g in got) and got == ref, trial


def test_plane_word_map_matches_the_crt_of_unit_coefficients():
    # ring_kernels.cu plane_word(c, s): coefficient X^(3i+c) is nonzero in exactly one word of every slot of its CRT image
    def plane_word(c, s):
        return 3 * s + (0 if c == 0 else (c if s < 4 else 3 - c))
    assert "return 3 * s + (c == 0 ? 0 : (s < 4 ? c : 3 - c));" in header("ring_kernels.cu")
    for t in range(24):
        e = [0] * 24
        e[t] = 1
        image = O.crt(e)
        nz = sorted(j for j, v in enumerate(image) if v)
        assert nz == sorted(plane_word(t % 3, s) for s in range(8)), t
    # ... and a ternary plane's CRT is the difference of two subset sums of those images (planes_fx_kernel)
    rnd = random.Random(11)
    basis = [O.crt([1 if j == t else 0 for j in range(24)]) for t in range(24)]
    for _ in range(50):
        d = [rnd.choice((-1, 0, 0, 1)) for _ in range(24)]
        pos = [sum(basis[t][j] for t in range(24) if d[t] > 0) % Q for j in range(24)]
        neg = [sum(basis[t][j] for t in range(24) if d[t] < 0) % Q for j in range(24)]
        assert [(p - m) % Q for p, m in zip(pos, neg)] == O.crt([x % Q for x in d])
