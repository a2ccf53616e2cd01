"""ctypes binding of latticeum_b200/lib/liblattice_ajtai.so -- the C ABI declared in include/lattice_ajtai.h.

This is the same binding a Rust `crates/zkvm-cuda` FFI crate would generate.  There is no CPU
fallback: if the shared library is missing or cannot be loaded, importing this module's `lib()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LAT_LIB") or os.path.join(_HERE, "lib", "liblattice_ajtai.so")  # LAT_LIB: tuning builds

# lat_status (include/lattice_ajtai.h)
LAT_OK = 0
LAT_E_WRONG_WITNESS_LENGTH = 1
LAT_E_WRONG_COMMITMENT_LENGTH = 2
LAT_E_WRONG_MATRIX_DIMENSIONS = 3
LAT_E_DIGIT_OVERFLOW = 4
LAT_E_INVALID_ARGUMENT = 5
LAT_E_CUDA = 6
LAT_E_MATRIX_INCOMPLETE = 7
LAT_REPR_CANONICAL = 0
LAT_REPR_MONTGOMERY = 1
LAT_ABI_VERSION = 1
LAT_PIPELINE_DEPTH = 4

_u64p = C.c_void_p  # raw addresses (host or device); arrays are passed by address
_H = C.c_void_p     # lat_ajtai*

# name -> (restype, argtypes); must list EVERY symbol include/lattice_ajtai.h declares
SIGNATURES = {
    "lat_strerror": (C.c_char_p, [C.c_int]),
    "lat_last_error": (C.c_char_p, []),
    "lat_abi_version": (C.c_int, []),
    "lat_ajtai_create": (C.c_int, [C.POINTER(_H), C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int]),
    "lat_ajtai_destroy": (None, [_H]),
    "lat_ajtai_upload_rows": (C.c_int, [_H, C.c_uint32, C.c_uint32, _u64p, C.c_uint64]),
    "lat_ajtai_upload_rows_dev": (C.c_int, [_H, C.c_uint32, C.c_uint32, _u64p, C.c_uint64]),
    "lat_ajtai_kappa": (C.c_uint32, [_H]),
    "lat_ajtai_width": (C.c_uint64, [_H]),
    "lat_ajtai_set_stream": (C.c_int, [_H, C.c_void_p]),
    "lat_ajtai_synchronize": (C.c_int, [_H]),
    "lat_ajtai_commit_ntt": (C.c_int, [_H, _u64p, C.c_uint64, _u64p]),
    "lat_ajtai_commit_ntt_dev": (C.c_int, [_H, _u64p, C.c_uint64, _u64p]),
    "lat_ajtai_commit_ntt_batch": (C.c_int, [_H, _u64p, C.c_uint32, C.c_uint64, _u64p]),
    "lat_ajtai_commit_ntt_batch_dev": (C.c_int, [_H, _u64p, C.c_uint32, C.c_uint64, _u64p]),
    "lat_ajtai_commit_coeff": (C.c_int, [_H, _u64p, C.c_uint64, _u64p]),
    "lat_ajtai_decompose_and_commit_coeff": (C.c_int, [_H, _u64p, C.c_uint64, _u64p]),
    "lat_ajtai_decompose_and_commit_ntt": (C.c_int, [_H, _u64p, C.c_uint64, _u64p]),
    "lat_ajtai_witness_from_w_ccs": (C.c_int, [_H, _u64p, C.c_uint64, _u64p, _u64p, _u64p]),
    "lat_ajtai_witness_from_w_ccs_dev": (C.c_int, [_H, _u64p, C.c_uint64, _u64p, _u64p, _u64p]),
    "lat_ajtai_witness_from_w_ccs_compact": (C.c_int, [_H, _u64p, C.c_uint64, C.c_void_p, _u64p, _u64p]),
    "lat_ajtai_set_accumulator": (C.c_int, [_H, _u64p, C.c_uint64, _u64p]),
    "lat_ajtai_fold_step_begin": (C.c_int, [_H, _u64p, C.c_uint64, _u64p, C.c_void_p, _u64p, _u64p]),
    "lat_ajtai_fold_step_finish": (C.c_int, [_H, _u64p, C.c_void_p, _u64p, _u64p, _u64p]),
    "lat_ajtai_get_fhat_dev": (C.c_int, [_H, C.c_int, _u64p]),
    "lat_ajtai_submit_w_ccs": (C.c_int, [_H, _u64p, C.c_uint64, _u64p, C.POINTER(C.c_uint64)]),
    "lat_ajtai_wait": (C.c_int, [_H, C.c_uint64]),
    "lat_ajtai_set_peers": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint64]),
    "lat_ajtai_decompose_commit": (C.c_int, [_H, _u64p, C.c_uint64, _u64p, _u64p, _u64p, _u64p]),
    "lat_ajtai_decompose_commit_dev": (C.c_int, [_H, _u64p, C.c_uint64, _u64p, _u64p, _u64p, _u64p]),
    "lat_ajtai_decompose_commit_resident": (C.c_int, [_H, _u64p, _u64p, _u64p, _u64p]),
    "lat_ajtai_select_side": (C.c_int, [_H, C.c_int]),
    "lat_ajtai_fold_witness": (C.c_int, [_H, _u64p, _u64p, _u64p]),
    "lat_ajtai_fold_witness_dev": (C.c_int, [_H, _u64p, _u64p, _u64p]),
    "lat_ring_gadget_recompose": (C.c_int, [_u64p, C.c_uint64, C.c_uint32, C.c_uint32, _u64p, C.c_int, C.c_int]),
    "lat_ring_crt": (C.c_int, [_u64p, C.c_uint64, _u64p, C.c_int]),
    "lat_ring_icrt": (C.c_int, [_u64p, C.c_uint64, _u64p, C.c_int]),
    "lat_ring_crt_dev": (C.c_int, [_u64p, C.c_uint64, _u64p, C.c_void_p]),
    "lat_ring_icrt_dev": (C.c_int, [_u64p, C.c_uint64, _u64p, C.c_void_p]),
    "lat_ring_gadget_decompose": (C.c_int, [_u64p, C.c_uint64, C.c_uint32, C.c_uint32, _u64p, C.c_int, C.c_int]),
    "lat_commitment_y0_dev": (C.c_int, [_u64p, _u64p, C.c_uint32, C.c_uint32, C.c_void_p]),
    "lat_commitment_sum_dev": (C.c_int, [_u64p, C.c_uint32, C.c_uint64, _u64p, C.c_void_p]),
    "lat_commitment_sum": (C.c_int, [_u64p, C.c_uint32, C.c_uint64, _u64p, C.c_int]),
    "lat_commitment_exchange_dev": (C.c_int, [_u64p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                              C.c_uint64, _u64p, C.c_void_p]),
    "lat_ajtai_witness_from_w_ccs_gated_dev": (C.c_int, [_H, _u64p, C.c_uint64, _u64p, _u64p, C.c_uint64]),
    "lat_commitment_exchange_report_dev": (C.c_int, [_u64p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                                     C.c_uint64, _u64p, _u64p, _u64p, C.c_uint64, C.c_void_p]),
    "lat_ntt_negacyclic": (C.c_int, [_u64p, C.c_uint64, C.c_uint32, C.c_int, _u64p, C.c_int]),
    "lat_ntt_negacyclic_dev": (C.c_int, [_u64p, C.c_uint64, C.c_uint32, C.c_int, _u64p, C.c_void_p]),
    "lat_ajtai_set_step_overlap": (C.c_int, [_H, C.c_int]),
    "lat_ajtai_set_profiling": (C.c_int, [_H, C.c_int]),
    "lat_ajtai_mac_profile": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "lat_set_spin_timeout_ms": (C.c_int, [C.c_uint64]),
    "lat_device_wait_status": (C.c_int, [C.c_int, C.POINTER(C.c_uint64)]),
    "lat_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "lat_host_free": (None, [C.c_void_p]),
}

_lib = None


class EngineUnavailable(RuntimeError):
    """The CUDA engine's shared library is missing or unloadable.  There is deliberately no fallback."""


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineUnavailable(
                f"{LIB_PATH} not found: build it with `python -m latticeum_b200.build` "
                "(nvcc, sm_100a).  latticeum_b200 has no CPU fallback."
            )
        try:
            L = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise EngineUnavailable(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        if L.lat_abi_version() != LAT_ABI_VERSION:
            raise EngineUnavailable("ABI version mismatch between _capi.py and liblattice_ajtai.so")
        _lib = L
    return _lib


def last_error() -> str:
    return lib().lat_last_error().decode()


def strerror(status: int) -> str:
    return lib().lat_strerror(status).decode()
