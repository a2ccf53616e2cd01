"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the header declares
(and nothing is bound that the header does not declare), and the host-side mirror logic behaves like the reference.
No compute entry point is called here (there is no GPU in this container)."""
import os
import re
import subprocess

import numpy as np
import pytest

from latticeum_b200 import _capi as capi
from latticeum_b200 import scheme as S
from oracle import lattice_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from latticeum_b200 import build

    return build.build()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "lattice_ajtai.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lat_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(capi.SIGNATURES)


def test_library_exports_every_declared_symbol(built):
    out = subprocess.check_output(["nm", "-D", "--defined-only", built], text=True)
    exported = set(re.findall(r" T (lat_[a-z0-9_]+)", out))
    assert set(header_symbols()) <= exported
    L = capi.lib()  # binds every symbol; raises if one is missing
    assert L.lat_abi_version() == capi.LAT_ABI_VERSION
    assert capi.strerror(capi.LAT_E_WRONG_WITNESS_LENGTH) == "wrong length of the witness"
    assert "CPU fallback" in capi.strerror(capi.LAT_E_CUDA)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "latticeum_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU fallback", ""), f"{f} mentions the oracle"


def test_create_fails_loudly_without_gpu(built):
    import ctypes as C

    h = C.c_void_p()
    st = capi.lib().lat_ajtai_create(C.byref(h), 4, 20, 15, 5, 15, 0, 0)
    if st == capi.LAT_OK:  # running on a GPU box
        capi.lib().lat_ajtai_destroy(h)
        pytest.skip("a GPU is present")
    assert st == capi.LAT_E_CUDA and not h.value
    with pytest.raises(S.EngineError):
        S.AjtaiCommitmentScheme(4, 20)


def test_invalid_arguments(built):
    import ctypes as C

    h = C.c_void_p()
    L = capi.lib()
    assert L.lat_ajtai_create(C.byref(h), 0, 20, 15, 5, 15, 0, 0) == capi.LAT_E_WRONG_MATRIX_DIMENSIONS
    assert L.lat_ajtai_create(C.byref(h), 4, 20, 16, 5, 15, 0, 0) == capi.LAT_E_INVALID_ARGUMENT
    assert L.lat_ajtai_create(C.byref(h), 4, 20, 15, 5, 15, 7, 0) == capi.LAT_E_INVALID_ARGUMENT
    assert L.lat_ajtai_commit_ntt(None, None, 0, None) == capi.LAT_E_INVALID_ARGUMENT


def test_every_handle_entry_point_rejects_a_null_handle(built):
    """Every exported function that takes the engine handle returns LAT_E_INVALID_ARGUMENT for a NULL handle (and does
    not touch CUDA or crash); the handle-less ones accept an empty batch.  Runs without a GPU."""
    import ctypes as C

    L = capi.lib()
    checked = 0
    for name, (res, args) in capi.SIGNATURES.items():
        if not name.startswith("lat_ajtai_") or name in ("lat_ajtai_create", "lat_ajtai_destroy"):
            continue
        zero = [None if (a is C.c_void_p or hasattr(a, "contents")) else 0 for a in args]
        got = getattr(L, name)(*zero)
        if name in ("lat_ajtai_kappa", "lat_ajtai_width"):
            assert got == 0, name
        else:
            assert got == capi.LAT_E_INVALID_ARGUMENT, (name, got)
        checked += 1
    assert checked >= 30
    L.lat_ajtai_destroy(None)  # a no-op
    for name in ("lat_ring_crt", "lat_ring_icrt"):
        assert getattr(L, name)(None, 0, None, 0) == 0
    assert L.lat_ring_gadget_decompose(None, 0, 15, 5, None, 0, 0) == 0
    assert L.lat_ring_gadget_recompose(None, 0, 15, 5, None, 0, 0) == 0
    assert L.lat_ntt_negacyclic(None, 0, 3, 0, None, 0) == 0
    assert L.lat_commitment_sum(None, 0, 0, None, 0) in (0, capi.LAT_E_INVALID_ARGUMENT)


def test_commitment_ops_match_reference_semantics():
    # latticefold/src/commitment/homomorphic_commitment.rs:54-80
    rng = np.random.default_rng(0)
    a = S._uniform((4, 24), 1)
    b = S._uniform((4, 24), 2)
    r = S._uniform((24,), 3)
    ca, cb = S.Commitment(a), S.Commitment(b)
    assert (ca + cb).val.tolist() == O.commitment_add(a.tolist(), b.tolist())
    assert (ca - cb).val.tolist() == O.commitment_sub(a.tolist(), b.tolist())
    assert (ca * r).val.tolist() == O.commitment_scale(a.tolist(), r.tolist())
    assert ca == S.Commitment(a.copy()) and ca != cb
    # Montgomery representation: canonical(r) * mont(x) = mont(r * x)
    cam = S.Commitment(S.to_mont(a), mont=True)
    assert S.from_mont((cam * S.to_mont(r)).val).tolist() == O.commitment_scale(a.tolist(), r.tolist())
    assert S.from_mont((cam + S.Commitment(S.to_mont(b), True)).val).tolist() == O.commitment_add(a.tolist(), b.tolist())
    assert ca.serialize()[:8] == (4).to_bytes(8, "little") and len(ca.serialize()) == 8 + 4 * 24 * 8
    assert cam.serialize() == ca.serialize()


def test_mont_helpers_and_scalar_embedding():
    x = S._uniform((100,), 5)
    assert S.to_mont(x).tolist() == [O.to_mont(int(v)) for v in x]
    assert S.from_mont(S.to_mont(x)).tolist() == x.tolist()
    assert S.ntt_from_scalar(7).tolist() == O.ntt_from_scalar(7)
    assert S.from_mont(S.ntt_from_scalar(7, mont=True)).tolist() == O.ntt_from_scalar(7)
    edge = np.array([0, 1, S.Q - 1, 2**32 - 1, 2**32, 2**63], dtype=np.uint64)
    assert S.to_mont(edge).tolist() == [O.to_mont(int(v)) for v in edge]


def test_get_fhat_matches_reference_layout():
    fc = S._uniform((5, 24), 9)
    got = S.get_fhat(fc)
    exp = O.get_fhat(fc.tolist())
    assert got.tolist() == exp


def test_params():
    assert S.GoldiLocksDP.log2_B == 15 and S.N == 98815 and S.KAPPA == 32
    with pytest.raises(ValueError):
        S.DecompositionParams(B=10, L=5, B_SMALL=2, K=15).log2_B


def test_bench_reference_arm_prints_one_json_line():
    # bench.py --impl reference times the CPU path (the oracle port) and prints exactly one JSON line on stdout
    import json
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, env={**os.environ, "OMP_NUM_THREADS": "4"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ring elems/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["higher_is_better"] is True
