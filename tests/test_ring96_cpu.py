"""The CRT / iCRT network in Z/(2^96 + 1) (latticeum_b200/csrc/ring96.cuh), built for the host with g++ (portable
fallbacks of the carry chains) and compared with the oracle: same algorithm layer the CUDA kernels instantiate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import c_oracle as CO

HERE = os.path.dirname(os.path.abspath(__file__))
Q = 2**64 - 2**32 + 1


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("r96") / "libr96.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(HERE, "native", "ring96_harness.cpp"), "-o", out])
    return C.CDLL(out)


def run(lib, fn, x):
    y = np.ascontiguousarray(x, dtype=np.uint64).copy()
    getattr(lib, fn)(C.c_void_p(y.ctypes.data), C.c_ulonglong(y.size // 24))
    return y


def edge_rows():
    vals = [0, 1, 2, Q - 1, Q - 2, Q // 2, Q // 2 + 1, 2**32 - 1, 2**32, 2**32 + 1, 2**63, 2**63 - 1, 0xFFFFFFFF00000000]
    rows = [[v] * 24 for v in vals]
    for k in range(24):  # one extreme coefficient at a time
        r = [0] * 24
        r[k] = Q - 1
        rows.append(r)
    return np.array(rows, dtype=np.uint64)


def test_crt_icrt_match_the_oracle(lib):
    x = np.concatenate([edge_rows(), CO.fill_uniform((20000, 24), 5)])
    assert np.array_equal(run(lib, "r96_crt", x), CO.crt(x))
    assert np.array_equal(run(lib, "r96_icrt", x), CO.icrt(x))
    assert np.array_equal(run(lib, "r96_icrt", run(lib, "r96_crt", x)), x)


def test_non_canonical_inputs_are_reduced(lib):
    # inputs are "any 64-bit representative": x and x + q (when it fits) transform alike
    x = CO.fill_uniform((500, 24), 6) % np.uint64(2**32 - 1)
    assert np.array_equal(run(lib, "r96_crt", x + np.uint64(Q)), CO.crt(x))
    assert np.array_equal(run(lib, "r96_icrt", x + np.uint64(Q)), CO.icrt(x))


@pytest.mark.parametrize("bound", [1, 2**14, 2**15 - 1])
def test_crt_small_matches_crt_of_the_digits(lib, bound):
    rng = np.random.default_rng(bound)
    d = rng.integers(-bound, bound + 1, size=(5000, 24)).astype(np.int32)
    d[0], d[1] = bound, -bound
    fq = np.where(d < 0, d.astype(np.int64).view(np.uint64) + np.uint64(Q), d.astype(np.int64).view(np.uint64))
    exp = CO.crt(fq)
    for mont in (0, 1):
        out = np.empty((d.shape[0], 24), np.uint64)
        lib.r96_crt_small(C.c_void_p(d.ctypes.data), C.c_ulonglong(d.shape[0]), mont, C.c_void_p(out.ctypes.data))
        assert np.array_equal(out, CO.to_mont(exp) if mont else exp)
