"""CPU oracle (Python big-int) for the Ajtai-commitment hot path of Nesquiko/Latticeum.

TEST INFRASTRUCTURE ONLY.  Nothing under ``latticeum_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may, and there only as the checker / the CPU arm, never as the thing shipped.

It is a restatement (not a copy) of the reference's algorithm, written from the maths with every
function citing the reference file:line it follows.  Paths are relative to
``/root/reference/latticeum/``; short names:

  GOLD  = crates/stark-rings/crates/ring/src/cyclotomic_ring/models/goldilocks
  RING  = crates/stark-rings/crates/ring/src
  LINALG= crates/stark-rings/crates/linear_algebra/src
  LF    = crates/latticefold/src

Parity pinning: ``tests/test_oracle_golden.py`` checks this module against every known-answer vector
the reference's own tests hold for the path (tests/golden/reference_kats.json lists them with their
file:line).  The arithmetic itself lives in ark-ff 0.5.0 (Cargo.lock:279-282, not vendored); it is
exact arithmetic in Z_q and Fq3, so it is restated from the maths and pinned by those KATs.
NOT pinned (no reference test fixes any value, the crate sources are absent and there is no Rust
toolchain here): the seeded sampler behind ``AjtaiCommitmentScheme::rand`` (ark-std 0.5.0
``test_rng`` + rand 0.8.5 StdRng + ``Fp::rand``) -- "sampler parity unpinned".

All values are canonical integers in [0, q) unless a name says ``mont``.
"""
from __future__ import annotations

from typing import List, Sequence

# ---------------------------------------------------------------------------------------------
# a1. Field Z_q and Fq3 = Fq[u]/(u^3 - 2^40)                     GOLD/mod.rs:16-54
# ---------------------------------------------------------------------------------------------
Q = 18446744069414584321  # 2^64 - 2^32 + 1                    GOLD/mod.rs:21
Q_HALF = (Q - 1) // 2
D = 24  # ring degree                                           GOLD/ntt.rs:9
NSLOT = 8  # CRT slots                                          GOLD/ntt.rs:12
NONRESIDUE = 1 << 40  # u^3 = 2^40                              GOLD/mod.rs:42
MONT_R = (1 << 64) % Q  # host in-memory form is x*2^64 mod q   GOLD/mod.rs:20-24 (MontBackend)
MONT_RINV = pow(MONT_R, Q - 2, Q)

# ROOTS_OF_UNITY_24[i] = (2^40)^i mod q                          GOLD/ntt.rs:15-40
W = [pow(NONRESIDUE, i, Q) for i in range(24)]
# KAPPA is the INVERSE of (2*zeta - 1), zeta = W[4] (the doc comment at GOLD/ntt.rs:42 says
# "2*zeta-1"; the literal at :43 is its inverse).
KAPPA = 12297829382473034411  #                                 GOLD/ntt.rs:43
EIGHT_INV = 16140901060737761281  #                             GOLD/ntt.rs:45
FOUR_INV = 13835058052060938241  #                              GOLD/ntt.rs:47


def to_mont(x: int) -> int:
    return (x * MONT_R) % Q


def from_mont(x: int) -> int:
    return (x * MONT_RINV) % Q


def fq3_mul(a: Sequence[int], b: Sequence[int]) -> List[int]:
    """Product in Fq[u]/(u^3 - NONRESIDUE) (ark-ff Fp3 semantics; value fixed by the maths)."""
    a0, a1, a2 = a
    b0, b1, b2 = b
    c0 = (a0 * b0 + NONRESIDUE * (a1 * b2 + a2 * b1)) % Q
    c1 = (a0 * b1 + a1 * b0 + NONRESIDUE * (a2 * b2)) % Q
    c2 = (a0 * b2 + a1 * b1 + a2 * b0) % Q
    return [c0, c1, c2]


# ---------------------------------------------------------------------------------------------
# a4/a5. CRT and iCRT of one ring element                        GOLD/ntt.rs:135-228, 240-319
# ---------------------------------------------------------------------------------------------
def homogenize(c: List[int]) -> None:
    """Per-slot isomorphism onto Fq[u]/(u^3 - NR).              GOLD/ntt.rs:326-334, 349-430"""
    # slot 1: NR^13
    c[4] = (-c[4]) % Q
    # slot 2: NR^7
    c[7] = c[7] * W[2] % Q
    c[8] = c[8] * W[4] % Q
    # slot 3: NR^19
    c[10] = c[10] * W[6] % Q
    c[11] = c[11] * W[12] % Q
    # slots 4..7 swap components 1 and 2 and scale
    for base, (m1, m2) in ((12, (3, 1)), (15, (11, 5)), (18, (7, 3)), (21, (15, 7))):
        c1 = c[base + 1]
        c[base + 1] = c[base + 2] * W[m1] % Q
        c[base + 2] = c1 * W[m2] % Q


def dehomogenize(c: List[int]) -> None:
    """Inverse of :func:`homogenize`.                            GOLD/ntt.rs:337-346, 355-437"""
    c[4] = (-c[4]) % Q
    c[7] = c[7] * W[22] % Q
    c[8] = c[8] * W[20] % Q
    c[10] = c[10] * W[18] % Q
    c[11] = c[11] * W[12] % Q
    for base, (m1, m2) in ((12, (23, 21)), (15, (19, 13)), (18, (21, 17)), (21, (17, 9))):
        c1 = c[base + 1]
        c[base + 1] = c[base + 2] * W[m1] % Q
        c[base + 2] = c1 * W[m2] % Q


def crt_raw(coeffs: Sequence[int]) -> List[int]:
    """The three butterfly layers, WITHOUT the final homogenize (the layout the reference's
    test_crt / test_crt2 expected arrays are written in).        GOLD/ntt.rs:135-226"""
    c = [x % Q for x in coeffs]
    assert len(c) == D
    # layer 1: mod X^12 - zeta / X^12 - zeta^5, zeta = W[4], zeta^5 = 1 - zeta      :146-152
    for i in range(12):
        a, b = c[i], c[12 + i]
        zb = W[4] * b % Q
        c[i] = (a + zb) % Q
        c[12 + i] = (a + b - zb) % Q
    # layer 2                                                                        :160-179
    for i in range(6):
        a, b = c[i], c[6 + i]
        t = W[2] * b % Q
        c[i], c[6 + i] = (a + t) % Q, (a - t) % Q
        a, b = c[12 + i], c[18 + i]
        t = W[10] * b % Q
        c[12 + i], c[18 + i] = (a + t) % Q, (a - t) % Q
    # layer 3                                                                        :186-225
    for i in range(3):
        for base, tw in ((0, 1), (6, 7), (12, 5), (18, 11)):
            a, b = c[base + i], c[base + 3 + i]
            t = W[tw] * b % Q
            c[base + i], c[base + 3 + i] = (a + t) % Q, (a - t) % Q
    return c


def crt(coeffs: Sequence[int]) -> List[int]:
    """24 coefficients -> 8 x Fq3, index = slot*3 + component.   GOLD/ntt.rs:135-228"""
    c = crt_raw(coeffs)
    homogenize(c)
    return c


def icrt_raw(c: List[int]) -> List[int]:
    """Inverse butterflies on a de-homogenized vector.            GOLD/ntt.rs:250-318"""
    c = list(c)
    for i in range(3):  #                                                            :250-283
        for base, tw in ((0, 23), (6, 17), (12, 19), (18, 13)):
            a, b = c[base + i], c[base + 3 + i]
            c[base + i] = (a + b) % Q
            c[base + 3 + i] = W[tw] * (a - b) % Q
    for i in range(6):  #                                                            :289-307
        a, b = c[i], c[6 + i]
        c[i], c[6 + i] = (a + b) % Q, W[22] * (a - b) % Q
        a, b = c[12 + i], c[18 + i]
        c[12 + i], c[18 + i] = (a + b) % Q, W[14] * (a - b) % Q
    for i in range(12):  #                                                           :310-317
        a, b = c[i], c[12 + i]
        kd = KAPPA * (a - b) % Q
        c[i] = EIGHT_INV * (a + b - kd) % Q
        c[12 + i] = FOUR_INV * kd % Q
    return c


def icrt(slots: Sequence[int]) -> List[int]:
    """8 x Fq3 -> 24 coefficients.                               GOLD/ntt.rs:240-319"""
    c = [x % Q for x in slots]
    assert len(c) == D
    dehomogenize(c)
    return icrt_raw(c)


def elementwise_crt(v: Sequence[Sequence[int]]) -> List[List[int]]:
    """RING/cyclotomic_ring/crt.rs:10-25"""
    return [crt(e) for e in v]


def elementwise_icrt(v: Sequence[Sequence[int]]) -> List[List[int]]:
    """RING/cyclotomic_ring/crt.rs:34-49"""
    return [icrt(e) for e in v]


# ---------------------------------------------------------------------------------------------
# a2/a3. Ring element helpers
# ---------------------------------------------------------------------------------------------
def ntt_from_scalar(v: int) -> List[int]:
    """All 8 slots equal (v, 0, 0).                              RING/cyclotomic_ring/ntt_form.rs:356-371"""
    out = [0] * D
    for s in range(NSLOT):
        out[3 * s] = v % Q
    return out


def ntt_mul(a: Sequence[int], b: Sequence[int]) -> List[int]:
    """Slot-wise Fq3 product.                                    RING/cyclotomic_ring/ntt_form.rs:159-175, 521-536"""
    out: List[int] = []
    for s in range(NSLOT):
        out += fq3_mul(a[3 * s : 3 * s + 3], b[3 * s : 3 * s + 3])
    return out


def ntt_add(a: Sequence[int], b: Sequence[int]) -> List[int]:
    return [(x + y) % Q for x, y in zip(a, b)]


def ntt_sub(a: Sequence[int], b: Sequence[int]) -> List[int]:
    return [(x - y) % Q for x, y in zip(a, b)]


def poly_mul(a: Sequence[int], b: Sequence[int]) -> List[int]:
    """Product in Z_q[X]/(X^24 - X^12 + 1) by schoolbook + reduction (X^24 = X^12 - 1).
    Reduction as GOLD/mod.rs:69-92 (reduce_in_place)."""
    full = [0] * (2 * D - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            full[i + j] = (full[i + j] + x * y) % Q
    for k in range(2 * D - 2, D - 1, -1):
        v = full[k]
        full[k] = 0
        full[k - 12] = (full[k - 12] + v) % Q
        full[k - 24] = (full[k - 24] - v) % Q
    return full[:D]


# ---------------------------------------------------------------------------------------------
# a7/a8. Signed representative and balanced digit decomposition
# ---------------------------------------------------------------------------------------------
def signed_rep(v: int) -> int:
    """[0,q) -> [-(q-1)/2, (q-1)/2].                             RING/balanced_decomposition/fq_convertible.rs:22-34"""
    v %= Q
    return v - Q if v > Q_HALF else v


def from_signed(v: int) -> int:
    """RING/balanced_decomposition/fq_convertible.rs:38-49"""
    return v % Q


def _trunc_div(a: int, b: int) -> int:
    qt = abs(a) // abs(b)
    return qt if (a >= 0) == (b >= 0) else -qt


def _trunc_rem(a: int, b: int) -> int:
    return a - b * _trunc_div(a, b)


def rounded_div(dividend: int, divisor: int) -> int:
    """LINALG/ops.rs:64-80"""
    if (dividend ^ divisor) >= 0:
        return _trunc_div(dividend + divisor // 2, divisor)
    return _trunc_div(dividend - divisor // 2, divisor)


class DigitOverflow(Exception):
    """The reference indexes ``out[current_i]`` unchecked (mod.rs:80,85,87): a value needing more
    than ``padding_size`` digits is an index-out-of-bounds panic there."""


def decompose_balanced(v: int, b: int, padding: int) -> List[int]:
    """One Fq -> ``padding`` balanced digits (as Fq).            RING/balanced_decomposition/mod.rs:62-103"""
    assert b >= 2 and b % 2 == 0
    cur = signed_rep(v)
    half = b // 2
    out_signed: List[int] = []
    while True:
        rem = _trunc_rem(cur, b)
        if abs(rem) <= half:
            digit = rem
            cur = _trunc_div(cur, b)
        else:
            digit = rem + b if rem < 0 else rem - b
            cur = _trunc_div(cur, b) + rounded_div(rem, b)
        if len(out_signed) >= padding:
            raise DigitOverflow(f"value {v} needs more than {padding} base-{b} digits")
        out_signed.append(digit)
        if cur == 0:
            break
    out_signed += [0] * (padding - len(out_signed))
    return [from_signed(d) for d in out_signed]


def ring_decompose(elem: Sequence[int], b: int, padding: int) -> List[List[int]]:
    """Coefficient-wise; digit l of coeff i -> out[l][i].        RING/cyclotomic_ring/coeff_form.rs:588-606"""
    out = [[0] * D for _ in range(padding)]
    for i, c in enumerate(elem):
        for l, dgt in enumerate(decompose_balanced(c, b, padding)):
            out[l][i] = dgt
    return out


def gadget_decompose(v: Sequence[Sequence[int]], b: int, padding: int) -> List[List[int]]:
    """out[i*L + l] = limb l of v[i].                            RING/balanced_decomposition/mod.rs:163-175"""
    out: List[List[int]] = []
    for e in v:
        out += ring_decompose(e, b, padding)
    return out


def decompose_to_vec(v: Sequence[Sequence[int]], b: int, padding: int) -> List[List[List[int]]]:
    """RING/balanced_decomposition/mod.rs:119-140"""
    return [ring_decompose(e, b, padding) for e in v]


def transpose(m: Sequence[Sequence]) -> List[List]:
    """LINALG/ops.rs:13-34 (rows here are equally long)."""
    if not m:
        return []
    return [[row[c] for row in m] for c in range(len(m[0]))]


def decompose_B_vec_into_k_vec(f_coeff: Sequence[Sequence[int]], b_small: int, K: int):
    """plane k, element j = digit k of f_coeff[j].               LF/nifs/decomposition/utils.rs:45-49"""
    return transpose(decompose_to_vec(f_coeff, b_small, K))


def recompose(chunk: Sequence[Sequence[int]], b_elem: Sequence[int], ntt_form: bool) -> List[int]:
    """Horner: result = result*b + v_i from the top limb.        RING/balanced_decomposition/mod.rs:105-117"""
    res = [0] * D
    for v_i in reversed(chunk):
        res = ntt_mul(res, b_elem) if ntt_form else poly_mul(res, b_elem)
        res = ntt_add(res, v_i)
    return res


def gadget_recompose(v: Sequence[Sequence[int]], b: int, padding: int, ntt_form: bool) -> List[List[int]]:
    """RING/balanced_decomposition/mod.rs:177-190"""
    if ntt_form:
        b_elem = ntt_from_scalar(b)
    else:
        b_elem = [b % Q] + [0] * (D - 1)
    return [recompose(v[i : i + padding], b_elem, ntt_form) for i in range(0, len(v) - len(v) % padding, padding)]


# ---------------------------------------------------------------------------------------------
# a13/a14. Matrix-vector product and the commitment scheme
# ---------------------------------------------------------------------------------------------
class WrongWitnessLength(Exception):
    """CommitmentError::WrongWitnessLength(got, expected).       LF/commitment.rs:13-17"""

    def __init__(self, got: int, expected: int):
        super().__init__(f"Wrong length of the witness: {got}, expected: {expected}")
        self.got, self.expected = got, expected


def checked_mul_vec(A: Sequence[Sequence[Sequence[int]]], v: Sequence[Sequence[int]]):
    """y[i] = sum_j A[i][j]*v[j], left fold from zero.           LINALG/matrix.rs:168-178"""
    ncols = len(A[0]) if A else 0
    if ncols != len(v):
        return None
    out = []
    for row in A:
        acc = [0] * D
        for a, f in zip(row, v):
            acc = ntt_add(acc, ntt_mul(a, f))
        out.append(acc)
    return out


def commit(A, f) -> List[List[int]]:
    """AjtaiCommitmentScheme::commit / commit_ntt.               LF/commitment/commitment_scheme.rs:63-80, 101-103"""
    ncols = len(A[0]) if A else 0
    if len(f) != ncols:
        raise WrongWitnessLength(len(f), ncols)
    return checked_mul_vec(A, f)


def commit_coeff(A, f_coeff):
    """LF/commitment/commitment_scheme.rs:107-112"""
    return commit(A, elementwise_crt(f_coeff))


def decompose_and_commit_coeff(A, f_coeff, B: int, L: int):
    """decompose_to_vec(B, L) flattened (element-major, limb-minor) -> CRT -> commit.
    LF/commitment/commitment_scheme.rs:116-127"""
    flat: List[List[int]] = []
    for limbs in decompose_to_vec(f_coeff, B, L):
        flat += limbs
    return commit_coeff(A, flat)


def decompose_and_commit_ntt(A, w, B: int, L: int):
    """LF/commitment/commitment_scheme.rs:132-139"""
    return decompose_and_commit_coeff(A, elementwise_icrt(w), B, L)


# ---------------------------------------------------------------------------------------------
# a17. Witness                                                   LF/arith.rs:214-362
# ---------------------------------------------------------------------------------------------
def witness_from_w_ccs(w_ccs, B: int, L: int):
    """Returns (f_coeff, f).  iCRT -> gadget_decompose(B, L) -> CRT.   LF/arith.rs:230-248"""
    w_coeff = elementwise_icrt(w_ccs)
    f_coeff = gadget_decompose(w_coeff, B, L)
    f = elementwise_crt(f_coeff)
    return f_coeff, f


def witness_from_f(f, B: int, L: int):
    """Returns (f_coeff, w_ccs).                                  LF/arith.rs:299-313"""
    f_coeff = elementwise_icrt(f)
    w_ccs = gadget_recompose(f, B, L, ntt_form=True)
    return f_coeff, w_ccs


def witness_from_f_coeff(f_coeff, B: int, L: int):
    """Returns (f, w_ccs).                                        LF/arith.rs:324-338"""
    f = elementwise_crt(f_coeff)
    w_ccs = gadget_recompose(f, B, L, ntt_form=True)
    return f, w_ccs


def get_fhat(f_coeff):
    """tau=3 tables; fhat[j][i] = slots (f_i[8j+t], 0, 0), t<8 (before truncate_lnze, which only
    trims trailing zero evaluations).                             LF/arith.rs:273-297"""
    tau = D // NSLOT
    fhat = [[[0] * D for _ in f_coeff] for _ in range(tau)]
    for i, f_i in enumerate(f_coeff):
        for j in range(tau):
            for t in range(NSLOT):
                fhat[j][i][3 * t] = f_i[NSLOT * j + t] % Q
    return fhat


# ---------------------------------------------------------------------------------------------
# a18. Decomposition prover helpers                               LF/nifs/decomposition.rs:162-201
# ---------------------------------------------------------------------------------------------
def decompose_witness(f_coeff, b_small: int, K: int):
    """K planes in coefficient form and their CRTs.               LF/nifs/decomposition.rs:162-167"""
    planes = decompose_B_vec_into_k_vec(f_coeff, b_small, K)
    return planes, [elementwise_crt(p) for p in planes]


def commitment_add(a, b):
    return [ntt_add(x, y) for x, y in zip(a, b)]


def commitment_sub(a, b):
    return [ntt_sub(x, y) for x, y in zip(a, b)]


def commitment_scale(a, r):
    return [ntt_mul(x, r) for x in a]


def commit_witnesses(A, planes_f, cm, b_small: int):
    """y_1..y_{K-1} by matrix, y_0 = cm - fold_rev((acc + y_i)*b).  LF/nifs/decomposition.rs:178-201"""
    kappa = len(A)
    b = ntt_from_scalar(b_small)
    ys = [commit(A, f) for f in planes_f[1:]]
    acc = [[0] * D for _ in range(kappa)]
    for y in reversed(ys):
        acc = commitment_scale(commitment_add(acc, y), b)
    return [commitment_sub(cm, acc)] + ys


# ---------------------------------------------------------------------------------------------
# a19. Folded witness                                             LF/nifs/folding.rs:258-268
# ---------------------------------------------------------------------------------------------
def compute_f_0(rho_s, f_s):
    n = len(f_s[0])
    acc = [[0] * D for _ in range(n)]
    for rho, f in zip(rho_s, f_s):
        acc = [ntt_add(a, ntt_mul(rho, w)) for a, w in zip(acc, f)]
    return acc


# ---------------------------------------------------------------------------------------------
# f4. Negacyclic NTT over Z_q[X]/(X^d + 1), d = 2^k            (ABSENT from the reference)
# ---------------------------------------------------------------------------------------------
# PARITY UNPINNED BY CONSTRUCTION: the reference's ring is Z_q[X]/(X^24 - X^12 + 1) and it contains no
# power-of-two transform, so no golden vector or reference output can exist.  The definition below (also in
# include/lattice_ajtai.h) is the specification; this O(d^2) restatement and the algebraic properties checked in
# tests/test_oracle_ntt.py are the only anchors.
GENERATOR = 7  # generates Z_q^* (q - 1 = 2^32 * 3 * 5 * 17 * 257 * 65537)


def ntt_psi(d: int) -> int:
    """A primitive 2d-th root of unity: 7^((q-1)/2d)."""
    assert d >= 1 and d & (d - 1) == 0 and (Q - 1) % (2 * d) == 0
    return pow(GENERATOR, (Q - 1) // (2 * d), Q)


def ntt_eval_at(a: Sequence[int], i: int, inverse: bool = False) -> int:
    """One output of the transform: forward A[i] = sum_j a[j] psi^((2i+1)j); inverse a[i] = d^-1 sum_k A[k] psi^(-(2k+1)i)."""
    d = len(a)
    psi = ntt_psi(d)
    if not inverse:
        r = pow(psi, 2 * i + 1, Q)
        acc, p = 0, 1
        for x in a:
            acc = (acc + x * p) % Q
            p = p * r % Q
        return acc
    psi_inv = pow(psi, Q - 2, Q)
    acc = 0
    for k, x in enumerate(a):
        acc = (acc + x * pow(psi_inv, (2 * k + 1) * i, Q)) % Q
    return acc * pow(d, Q - 2, Q) % Q


def ntt_negacyclic(a: Sequence[int], inverse: bool = False) -> List[int]:
    return [ntt_eval_at(a, i, inverse) for i in range(len(a))]


def negacyclic_mul(a: Sequence[int], b: Sequence[int]) -> List[int]:
    """a * b mod (X^d + 1), schoolbook."""
    d = len(a)
    out = [0] * d
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            k = i + j
            if k < d:
                out[k] = (out[k] + x * y) % Q
            else:
                out[k - d] = (out[k - d] - x * y) % Q
    return out
