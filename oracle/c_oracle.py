"""ctypes/numpy front end of oracle/liblattice_oracle.so (the C restatement of the reference's CPU path).

TEST INFRASTRUCTURE ONLY -- see the header of lattice_oracle.c.  Arrays are numpy uint64, canonical,
shape (..., 24) per ring element.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblattice_oracle.so")
Q = 2**64 - 2**32 + 1
OK, E_WRONG_WITNESS_LENGTH, E_DIGIT_OVERFLOW, E_INVALID_ARG = 0, 1, 4, 5


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lattice_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liblattice_oracle.so"])
    return _SO


_lib = None
_P = C.POINTER(C.c_uint64)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64, i32 = C.c_uint64, C.c_int
        sigs = {
            "lo_roots": (None, [_P]),
            "lo_num_threads": (i32, []),
            "lo_set_num_threads": (None, [i32]),
            "lo_homogenize": (None, [_P]),
            "lo_dehomogenize": (None, [_P]),
            "lo_crt": (None, [_P, u64, _P]),
            "lo_icrt": (None, [_P, u64, _P]),
            "lo_crt_par": (None, [_P, u64, _P]),
            "lo_icrt_par": (None, [_P, u64, _P]),
            "lo_gadget_decompose": (i32, [_P, u64, u64, i32, _P]),
            "lo_decompose_planes": (i32, [_P, u64, u64, i32, _P]),
            "lo_gadget_recompose_ntt": (None, [_P, u64, u64, i32, _P]),
            "lo_commit": (i32, [_P, u64, u64, _P, u64, _P]),
            "lo_witness_from_w_ccs": (i32, [_P, u64, u64, i32, _P, _P]),
            "lo_decompose_commit": (i32, [_P, u64, u64, _P, _P, u64, i32, _P, _P, _P]),
            "lo_compute_f0": (None, [_P, C.POINTER(_P), i32, u64, _P]),
            "lo_to_mont": (None, [_P, u64, _P]),
            "lo_from_mont": (None, [_P, u64, _P]),
            "lo_fill_uniform": (None, [_P, u64, u64]),
        }
        for name, (res, args) in sigs.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P)


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


class OracleError(Exception):
    def __init__(self, status):
        super().__init__(f"oracle status {status}")
        self.status = status


def num_threads() -> int:
    return lib().lo_num_threads()


def set_num_threads(n: int) -> None:
    """OpenMP team size for the following calls (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().lo_set_num_threads(int(n))


def roots() -> np.ndarray:
    out = np.empty(24, np.uint64)
    lib().lo_roots(_p(out))
    return out


def homogenize(c):
    c = _c(c).copy()
    lib().lo_homogenize(_p(c))
    return c


def dehomogenize(c):
    c = _c(c).copy()
    lib().lo_dehomogenize(_p(c))
    return c


def crt(x, parallel=False):
    x = _c(x)
    out = np.empty_like(x)
    (lib().lo_crt_par if parallel else lib().lo_crt)(_p(x), x.size // 24, _p(out))
    return out


def icrt(x, parallel=False):
    x = _c(x)
    out = np.empty_like(x)
    (lib().lo_icrt_par if parallel else lib().lo_icrt)(_p(x), x.size // 24, _p(out))
    return out


def gadget_decompose(x, b: int, L: int):
    x = _c(x)
    cnt = x.size // 24
    out = np.empty((cnt * L, 24), np.uint64)
    st = lib().lo_gadget_decompose(_p(x), cnt, b, L, _p(out))
    if st:
        raise OracleError(st)
    return out


def decompose_planes(f_coeff, b: int, K: int):
    f_coeff = _c(f_coeff)
    n = f_coeff.size // 24
    out = np.empty((K, n, 24), np.uint64)
    st = lib().lo_decompose_planes(_p(f_coeff), n, b, K, _p(out))
    if st:
        raise OracleError(st)
    return out


def gadget_recompose_ntt(f, b: int, L: int):
    f = _c(f)
    n = f.size // 24
    out = np.empty((n // L, 24), np.uint64)
    lib().lo_gadget_recompose_ntt(_p(f), n, b, L, _p(out))
    return out


def commit(A, f):
    A, f = _c(A), _c(f)
    kappa, n = A.shape[0], A.shape[1]
    cm = np.empty((kappa, 24), np.uint64)
    st = lib().lo_commit(_p(A), kappa, n, _p(f), f.size // 24, _p(cm))
    if st:
        raise OracleError(st)
    return cm


def witness_from_w_ccs(w, B: int, L: int):
    w = _c(w)
    wl = w.size // 24
    f_coeff = np.empty((wl * L, 24), np.uint64)
    f = np.empty((wl * L, 24), np.uint64)
    st = lib().lo_witness_from_w_ccs(_p(w), wl, B, L, _p(f_coeff), _p(f))
    if st:
        raise OracleError(st)
    return f_coeff, f


def decompose_commit(A, f_coeff, cm, b: int, K: int, want_planes=True):
    A, f_coeff, cm = _c(A), _c(f_coeff), _c(cm)
    kappa, n = A.shape[0], A.shape[1]
    pc = np.empty((K, n, 24), np.uint64) if want_planes else None
    pf = np.empty((K, n, 24), np.uint64) if want_planes else None
    cms = np.empty((K, kappa, 24), np.uint64)
    st = lib().lo_decompose_commit(
        _p(A), kappa, n, _p(f_coeff), _p(cm), b, K, _p(pc) if want_planes else None, _p(pf) if want_planes else None, _p(cms)
    )
    if st:
        raise OracleError(st)
    return pc, pf, cms


def compute_f0(rho, f_s):
    rho = _c(rho)
    f_s = [_c(f) for f in f_s]
    n = f_s[0].size // 24
    ptrs = (_P * len(f_s))(*[_p(f) for f in f_s])
    out = np.empty((n, 24), np.uint64)
    lib().lo_compute_f0(_p(rho), ptrs, len(f_s), n, _p(out))
    return out


def to_mont(x):
    x = _c(x)
    out = np.empty_like(x)
    lib().lo_to_mont(_p(x), x.size, _p(out))
    return out


def from_mont(x):
    x = _c(x)
    out = np.empty_like(x)
    lib().lo_from_mont(_p(x), x.size, _p(out))
    return out


def fill_uniform(shape, seed: int) -> np.ndarray:
    out = np.empty(shape, np.uint64)
    lib().lo_fill_uniform(_p(out), out.size, seed)
    return out


def scalar_elem(v) -> np.ndarray:
    """NTT-form embedding of a base-field scalar: all slots (v,0,0).  RING/cyclotomic_ring/ntt_form.rs:356-371"""
    out = np.zeros(24, np.uint64)
    out[0::3] = np.uint64(int(v) % Q)
    return out
