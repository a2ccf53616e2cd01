"""The power-of-two negacyclic NTT has no counterpart in the reference: these checks pin the oracle's
definition against the algebra it must satisfy (parity is unpinned by construction, see oracle/lattice_oracle.py)."""
import random

import pytest

from oracle import lattice_oracle as O

Q = O.Q


@pytest.mark.parametrize("d", [1, 2, 8, 64, 1 << 14])
def test_psi_is_a_primitive_2d_th_root(d):
    psi = O.ntt_psi(d)
    assert pow(psi, d, Q) == Q - 1          # psi^d = -1: the powers psi^(2i+1) are the roots of X^d + 1
    assert pow(psi, 2 * d, Q) == 1


def test_generator_generates():
    for p in (2, 3, 5, 17, 257, 65537):      # the prime factors of q - 1
        assert pow(O.GENERATOR, (Q - 1) // p, Q) != 1


@pytest.mark.parametrize("d", [2, 8, 32])
def test_roundtrip_monomials_and_convolution(d):
    rng = random.Random(d)
    a = [rng.randrange(Q) for _ in range(d)]
    b = [rng.randrange(Q) for _ in range(d)]
    A, B = O.ntt_negacyclic(a), O.ntt_negacyclic(b)
    assert O.ntt_negacyclic(A, inverse=True) == a
    # X -> the roots themselves
    x = [0, 1] + [0] * (d - 2)
    psi = O.ntt_psi(d)
    assert O.ntt_negacyclic(x) == [pow(psi, 2 * i + 1, Q) for i in range(d)]
    # pointwise product <-> product mod X^d + 1
    assert [u * v % Q for u, v in zip(A, B)] == O.ntt_negacyclic(O.negacyclic_mul(a, b))
