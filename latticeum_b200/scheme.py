"""Host-side mirror of the reference's commitment / witness API over the CUDA engine's C ABI.

The reference is Rust and this image has no Rust toolchain, so the host layer a `zkvm` caller would use is
mirrored here in Python with the reference's names, argument meaning and error behaviour (paths relative to
/root/reference/latticeum/crates/):

    AjtaiCommitmentScheme   latticefold/src/commitment/commitment_scheme.rs:38-140
    Commitment              latticefold/src/commitment/homomorphic_commitment.rs:12-80
    CommitmentError         latticefold/src/commitment.rs:13-26
    DecompositionParams     latticefold/src/decomposition_parameters.rs:11-20  (GoldiLocksDP: zkvm/src/ccs.rs:26-34)
    Witness                 latticefold/src/arith.rs:214-362
    LFDecompositionProver.{decompose_witness, commit_witnesses}   latticefold/src/nifs/decomposition.rs:162-201

Ring elements are numpy uint64 arrays with a trailing axis of 24 (CRT form: slot*3 + component; coefficient
form: degree).  `repr` says whether limbs are canonical integers or the Montgomery form ark-ff keeps in memory.
All heavy work happens on the GPU through liblattice_ajtai.so; only the 6 KB `Commitment` arithmetic (a15 in
SURVEY.md 8a: "stay on host") is done here.  There is no CPU fallback for the engine.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _capi as capi

Q = 2**64 - 2**32 + 1
D = 24
MONT_R = 2**32 - 1  # 2^64 mod q
MONT_RINV = pow(2, 128, Q)


# ---- errors ---------------------------------------------------------------------------------------------------
class CommitmentError(Exception):
    """latticefold/src/commitment.rs:13-26"""


class WrongWitnessLength(CommitmentError):
    def __init__(self, got: int, expected: int):
        super().__init__(f"Wrong length of the witness: {got}, expected: {expected}")
        self.got, self.expected = got, expected


class WrongCommitmentLength(CommitmentError):
    def __init__(self, got: int, expected: int):
        super().__init__(f"Wrong length of the commitment: {got}, expected: {expected}")
        self.got, self.expected = got, expected


class WrongAjtaiMatrixDimensions(CommitmentError):
    def __init__(self, rows: int, cols: int, exp_rows: int, exp_cols: int):
        super().__init__(f"Ajtai matrix has dimensions: {rows}x{cols}, expected: {exp_rows}x{exp_cols}")


class DigitOverflow(Exception):
    """A coefficient needs more digits than the padding.  The reference panics here (index out of bounds at
    stark-rings/crates/ring/src/balanced_decomposition/mod.rs:80); the engine reports LAT_E_DIGIT_OVERFLOW."""


class EngineError(RuntimeError):
    """CUDA failure or misuse of the engine (no reference counterpart)."""


def _raise(status: int, got: int = 0, expected: int = 0):
    if status == capi.LAT_OK:
        return
    msg = capi.last_error()
    if status == capi.LAT_E_WRONG_WITNESS_LENGTH:
        raise WrongWitnessLength(got, expected)
    if status == capi.LAT_E_WRONG_COMMITMENT_LENGTH:
        raise WrongCommitmentLength(got, expected)
    if status == capi.LAT_E_WRONG_MATRIX_DIMENSIONS:
        raise CommitmentError(msg)
    if status == capi.LAT_E_DIGIT_OVERFLOW:
        raise DigitOverflow(msg)
    raise EngineError(f"{capi.strerror(status)}: {msg}")


# ---- parameters -------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class DecompositionParams:
    """latticefold/src/decomposition_parameters.rs:11-20"""

    B: int
    L: int
    B_SMALL: int
    K: int

    @property
    def log2_B(self) -> int:
        lb = self.B.bit_length() - 1
        if 1 << lb != self.B:
            raise ValueError("the engine needs B to be a power of two")
        return lb


GoldiLocksDP = DecompositionParams(B=1 << 15, L=5, B_SMALL=2, K=15)  # zkvm/src/ccs.rs:26-34
KAPPA = 32  # zkvm/src/ccs.rs:43
W_SIZE = 19763  # CCSLayout::new().w_size
N = W_SIZE * GoldiLocksDP.L  # zkvm/src/ccs.rs:50


def _as_u64(a, name: str) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim < 1 or a.shape[-1] != D:
        raise ValueError(f"{name}: ring elements need a trailing axis of {D} uint64")
    return a


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


# ---- small host-side field helpers (only for the 6 KB Commitment objects) -------------------------------------------
def _fq3_mul_scalar_vec(x: np.ndarray, r: np.ndarray, mont: bool) -> np.ndarray:
    """Slot-wise Fq3 product of every element of x (k,24) with the ring element r (24,).  ntt_form.rs:159-175"""
    nr = 1 << 40
    out = np.empty_like(x)
    rr = [int(v) for v in r]
    if mont:
        rr = [v * MONT_RINV % Q for v in rr]  # canonical(r) * mont(x) = mont(r*x)
    for e in range(x.shape[0]):
        xe = [int(v) for v in x[e]]
        for s in range(8):
            a0, a1, a2 = xe[3 * s : 3 * s + 3]
            b0, b1, b2 = rr[3 * s : 3 * s + 3]
            out[e, 3 * s] = (a0 * b0 + nr * (a1 * b2 + a2 * b1)) % Q
            out[e, 3 * s + 1] = (a0 * b1 + a1 * b0 + nr * a2 * b2) % Q
            out[e, 3 * s + 2] = (a0 * b2 + a1 * b1 + a2 * b0) % Q
    return out


class Commitment:
    """Vec<R> of length kappa with element-wise +=, -= and scaling by a ring element.
    latticefold/src/commitment/homomorphic_commitment.rs:12-80"""

    def __init__(self, val, mont: bool = False):
        self.val = _as_u64(val, "commitment").reshape(-1, D).copy()
        self.mont = mont

    @classmethod
    def from_vec_raw(cls, vec, mont: bool = False) -> "Commitment":
        return cls(vec, mont)

    @classmethod
    def zeroed(cls, kappa: int, mont: bool = False) -> "Commitment":
        return cls(np.zeros((kappa, D), np.uint64), mont)

    def __len__(self) -> int:
        return self.val.shape[0]

    def is_empty(self) -> bool:
        return len(self) == 0

    def as_ref(self) -> np.ndarray:
        return self.val

    def __eq__(self, other) -> bool:
        return isinstance(other, Commitment) and self.val.shape == other.val.shape and bool(np.array_equal(self.val, other.val))

    def __hash__(self):
        return hash(self.val.tobytes())

    def _zip(self, other: "Commitment", sign: int) -> "Commitment":
        a = self.val.astype(object)
        b = other.val.astype(object)
        n = min(len(self), len(other))  # zip semantics of the reference's iterators
        out = self.val.copy()
        out[:n] = ((a[:n] + sign * b[:n]) % Q).astype(np.uint64)
        return Commitment(out, self.mont)

    def __add__(self, other: "Commitment") -> "Commitment":
        return self._zip(other, +1)

    def __sub__(self, other: "Commitment") -> "Commitment":
        return self._zip(other, -1)

    def __mul__(self, r) -> "Commitment":
        r = _as_u64(r, "scalar ring element").reshape(D)
        return Commitment(_fq3_mul_scalar_vec(self.val, r, self.mont), self.mont)

    def serialize(self) -> bytes:
        """CanonicalSerialize layout: 8-byte LE length + kappa*24 canonical LE u64."""
        v = self.val
        if self.mont:
            v = np.array([[int(x) * MONT_RINV % Q for x in row] for row in v], dtype=np.uint64)
        return int(len(self)).to_bytes(8, "little") + v.astype("<u8").tobytes()


def ntt_from_scalar(v: int, mont: bool = False) -> np.ndarray:
    """RqNTT::from(u128): all 8 slots = (v, 0, 0).  ntt_form.rs:356-371"""
    out = np.zeros(D, np.uint64)
    x = int(v) % Q
    out[0::3] = np.uint64(x * MONT_R % Q if mont else x)
    return out


# ---- the scheme --------------------------------------------------------------------------------------------------------
class AjtaiCommitmentScheme:
    """AjtaiCommitmentScheme<GoldilocksRingNTT> backed by the device-resident matrix.
    latticefold/src/commitment/commitment_scheme.rs:38-140"""

    def __init__(self, kappa: int, n: int, params: DecompositionParams = GoldiLocksDP, mont: bool = False, device: int = 0):
        if params.B_SMALL != 2:
            raise ValueError("the engine fixes B_SMALL = 2 (zkvm/src/ccs.rs:31)")
        self.params = params
        self.mont = mont
        self.device = device
        self._kappa, self._n = int(kappa), int(n)
        self._h = C.c_void_p()
        st = capi.lib().lat_ajtai_create(
            C.byref(self._h), kappa, n, params.log2_B, params.L, params.K,
            capi.LAT_REPR_MONTGOMERY if mont else capi.LAT_REPR_CANONICAL, device,
        )
        _raise(st)

    # -- construction ------------------------------------------------------------------------------------------------
    @classmethod
    def new(cls, matrix, params: DecompositionParams = GoldiLocksDP, mont: bool = False, device: int = 0) -> "AjtaiCommitmentScheme":
        """AjtaiCommitmentScheme::new(Matrix<R>)  (:49).  matrix: (kappa, n, 24) or a list of kappa rows (n, 24)."""
        rows = [_as_u64(r, "matrix row") for r in matrix]
        if not rows:
            raise WrongAjtaiMatrixDimensions(0, 0, 1, 1)
        n = rows[0].shape[0]
        if any(r.shape != (n, D) for r in rows):
            raise WrongAjtaiMatrixDimensions(len(rows), n, len(rows), n)
        s = cls(len(rows), n, params, mont, device)
        for i, r in enumerate(rows):  # the host matrix is Vec<Vec<R>>: rows are separate allocations
            s.upload_rows(i, r[None])
        return s

    @classmethod
    def rand(cls, kappa: int, n: int, seed: int = 0, **kw) -> "AjtaiCommitmentScheme":
        """AjtaiCommitmentScheme::rand (:56-58) is `vec![vec![R::rand(rng); n]; kappa]`: ONE sampled ring element
        cloned into every entry.  This mirrors that structure; the sampler itself (ark-std test_rng /
        ChaCha) is not reproducible here -- sampler parity unpinned.  Use `rand_independent` for real matrices."""
        elem = _uniform((1, D), seed)
        s = cls(kappa, n, **kw)
        row = np.broadcast_to(elem, (n, D)).copy()
        if s.mont:
            row = _to_mont_np(row)
        for i in range(kappa):
            s.upload_rows(i, row[None])
        return s

    @classmethod
    def rand_independent(cls, kappa: int, n: int, seed: int = 0, **kw) -> "AjtaiCommitmentScheme":
        """Every entry sampled independently (the semantics of Matrix::rand, linear_algebra/src/matrix.rs:93-98)."""
        s = cls(kappa, n, **kw)
        for i in range(kappa):
            row = _uniform((n, D), seed * 1000003 + i)
            if s.mont:
                row = _to_mont_np(row)
            s.upload_rows(i, row[None])
        return s

    def upload_rows(self, row0: int, rows: np.ndarray, row_stride: Optional[int] = None) -> None:
        rows = _as_u64(rows, "rows")
        if rows.ndim != 3:
            raise ValueError("rows must be (nrows, stride, 24)")
        stride = rows.shape[1] if row_stride is None else row_stride
        _raise(capi.lib().lat_ajtai_upload_rows(self._h, row0, rows.shape[0], _ptr(rows), stride))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            capi.lib().lat_ajtai_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- accessors (:85-94) ------------------------------------------------------------------------------------------------
    def kappa(self) -> int:
        return int(capi.lib().lat_ajtai_kappa(self._h))

    def width(self) -> int:
        return int(capi.lib().lat_ajtai_width(self._h))

    # -- commitments -------------------------------------------------------------------------------------------------------
    def commit(self, f) -> Commitment:
        """commit (:63-80): f.len() != ncols -> WrongWitnessLength(f.len(), ncols)."""
        f = _as_u64(f, "f").reshape(-1, D)
        cm = np.empty((self._kappa, D), np.uint64)
        _raise(capi.lib().lat_ajtai_commit_ntt(self._h, _ptr(f), f.shape[0], _ptr(cm)), f.shape[0], self._n)
        return Commitment(cm, self.mont)

    def commit_ntt(self, f) -> Commitment:
        """commit_ntt (:101-103)"""
        return self.commit(f)

    def commit_ntt_batch(self, fs) -> List[Commitment]:
        """`count` witnesses in one launch (the loop of latticefold/src/nifs/decomposition.rs:185-187, batched)."""
        fs = _as_u64(fs, "fs")
        if fs.ndim != 3:
            raise ValueError("fs must be (count, n, 24)")
        cms = np.empty((fs.shape[0], self._kappa, D), np.uint64)
        _raise(capi.lib().lat_ajtai_commit_ntt_batch(self._h, _ptr(fs), fs.shape[0], fs.shape[1], _ptr(cms)), fs.shape[1], self._n)
        return [Commitment(c, self.mont) for c in cms]

    def commit_coeff(self, f_coeff) -> Commitment:
        """commit_coeff (:107-112): CRT then commit."""
        f = _as_u64(f_coeff, "f_coeff").reshape(-1, D)
        cm = np.empty((self._kappa, D), np.uint64)
        _raise(capi.lib().lat_ajtai_commit_coeff(self._h, _ptr(f), f.shape[0], _ptr(cm)), f.shape[0], self._n)
        return Commitment(cm, self.mont)

    def decompose_and_commit_coeff(self, f_coeff) -> Commitment:
        """decompose_and_commit_coeff (:116-127)"""
        w = _as_u64(f_coeff, "f_coeff").reshape(-1, D)
        cm = np.empty((self._kappa, D), np.uint64)
        st = capi.lib().lat_ajtai_decompose_and_commit_coeff(self._h, _ptr(w), w.shape[0], _ptr(cm))
        _raise(st, w.shape[0] * self.params.L, self._n)
        return Commitment(cm, self.mont)

    def decompose_and_commit_ntt(self, w) -> Commitment:
        """decompose_and_commit_ntt (:132-139)"""
        w = _as_u64(w, "w").reshape(-1, D)
        cm = np.empty((self._kappa, D), np.uint64)
        st = capi.lib().lat_ajtai_decompose_and_commit_ntt(self._h, _ptr(w), w.shape[0], _ptr(cm))
        _raise(st, w.shape[0] * self.params.L, self._n)
        return Commitment(cm, self.mont)


class CommitPipeline:
    """A stream of per-step commitments (zkvm/src/main.rs:121-219 commits one witness per VM step, :348-367): `submit`
    queues Witness::from_w_ccs + Witness::commit for one step and returns a ticket at once, `wait` returns that step's
    Commitment.  Up to `depth` steps are in flight; the PCIe transfers of neighbouring steps overlap the kernels when
    the w_ccs buffers are page-locked (`pinned_empty`).  Buffers passed to `submit` must stay untouched until `wait`."""

    depth = capi.LAT_PIPELINE_DEPTH

    def __init__(self, scheme: "AjtaiCommitmentScheme"):
        self.scheme = scheme
        self._out = {}

    def submit(self, w_ccs) -> int:
        w = _as_u64(w_ccs, "w_ccs").reshape(-1, D)
        cm = np.empty((self.scheme._kappa, D), np.uint64)
        ticket = C.c_uint64(0)
        st = capi.lib().lat_ajtai_submit_w_ccs(self.scheme._h, _ptr(w), w.shape[0], _ptr(cm), C.byref(ticket))
        _raise(st, w.shape[0] * self.scheme.params.L, self.scheme._n)
        self._out[ticket.value] = (cm, w)  # keep both alive until the ticket is waited for
        return ticket.value

    def wait(self, ticket: int) -> Commitment:
        cm, _ = self._out.pop(ticket)
        _raise(capi.lib().lat_ajtai_wait(self.scheme._h, ticket))
        return Commitment(cm, self.scheme.mont)


def pinned_empty(shape, device: int = 0) -> np.ndarray:
    """A page-locked uint64 array (lat_host_alloc): uploads from it run on a copy engine and overlap the kernels.
    The memory is released when the array is garbage collected."""
    import weakref

    count = int(np.prod(shape))
    ptr = C.c_void_p()
    _raise(capi.lib().lat_host_alloc(C.byref(ptr), max(count, 1) * 8))
    buf = (C.c_uint64 * max(count, 1)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=np.uint64, count=count).reshape(shape)
    weakref.finalize(buf, capi.lib().lat_host_free, ptr)
    return arr


# ---- Witness --------------------------------------------------------------------------------------------------------------
class Witness:
    """latticefold/src/arith.rs:214-223.  `f_hat` (the MLE tables) stays with the host's MLE code and is derived
    from f_coeff by `get_fhat` (arith.rs:273-297); it is not on the GPU path."""

    def __init__(self, w_ccs, f, f_coeff, mont: bool = False):
        self.w_ccs, self.f, self.f_coeff, self.mont = w_ccs, f, f_coeff, mont

    @classmethod
    def from_w_ccs(cls, scheme: AjtaiCommitmentScheme, w_ccs, want_f: bool = True, want_f_coeff: bool = True,
                   commit: bool = False):
        """Witness::from_w_ccs::<P> (arith.rs:230-248), optionally fused with Witness::commit (arith.rs:357-362) as
        at the call site zkvm/src/main.rs:357-363.  Returns the witness, or (witness, commitment) if `commit`."""
        w = _as_u64(w_ccs, "w_ccs").reshape(-1, D)
        n = w.shape[0] * scheme.params.L
        f = np.empty((n, D), np.uint64) if want_f else None
        fc = np.empty((n, D), np.uint64) if want_f_coeff else None
        cm = np.empty((scheme._kappa, D), np.uint64) if commit else None
        st = capi.lib().lat_ajtai_witness_from_w_ccs(scheme._h, _ptr(w), w.shape[0], _ptr(fc), _ptr(f), _ptr(cm))
        _raise(st, n, scheme._n)
        wit = cls(w, f, fc, scheme.mont)
        return (wit, Commitment(cm, scheme.mont)) if commit else wit

    @classmethod
    def from_w_ccs_compact(cls, scheme: AjtaiCommitmentScheme, w_ccs, want_f: bool = False, commit: bool = True):
        """Witness::from_w_ccs (+ commit) fetching f_coeff as the device's int16 digits (lat_ajtai_witness_from_w_ccs_compact);
        `f_coeff` is widened on the host (digits_to_fq).  `wit.digits` keeps the int16 rows for get_fhat_from_digits."""
        w = _as_u64(w_ccs, "w_ccs").reshape(-1, D)
        n = w.shape[0] * scheme.params.L
        d16 = np.empty((n, D), np.int16)
        f = np.empty((n, D), np.uint64) if want_f else None
        cm = np.empty((scheme._kappa, D), np.uint64) if commit else None
        st = capi.lib().lat_ajtai_witness_from_w_ccs_compact(scheme._h, _ptr(w), w.shape[0], _ptr(d16), _ptr(f), _ptr(cm))
        _raise(st, n, scheme._n)
        wit = cls(w, f, digits_to_fq(d16, scheme.mont), scheme.mont)
        wit.digits = d16
        return (wit, Commitment(cm, scheme.mont)) if commit else wit

    @classmethod
    def from_f_coeff(cls, scheme: AjtaiCommitmentScheme, f_coeff):
        """Witness::from_f_coeff (arith.rs:324-338): f = CRT(f_coeff).  (w_ccs = gadget_recompose(f) is host-side MLE
        plumbing, SURVEY 8 f2, and is left None.)"""
        fc = _as_u64(f_coeff, "f_coeff").reshape(-1, D)
        f = np.empty_like(fc)
        _raise(capi.lib().lat_ring_crt(_ptr(fc), fc.shape[0], _ptr(f), scheme.device))
        return cls(None, f, fc, scheme.mont)

    @classmethod
    def from_f(cls, scheme: AjtaiCommitmentScheme, f):
        """Witness::from_f (arith.rs:299-313): f_coeff = iCRT(f)."""
        f = _as_u64(f, "f").reshape(-1, D)
        fc = np.empty_like(f)
        _raise(capi.lib().lat_ring_icrt(_ptr(f), f.shape[0], _ptr(fc), scheme.device))
        return cls(None, f, fc, scheme.mont)

    def commit(self, scheme: AjtaiCommitmentScheme) -> Commitment:
        """Witness::commit (arith.rs:357-362) = scheme.commit_ntt(&self.f)"""
        return scheme.commit_ntt(self.f)


def digits_to_fq(d16, mont: bool = False) -> np.ndarray:
    """Widen the engine's int16 digits (Witness.f_coeff as the device holds it) into Fq limbs in the caller's
    representation: d >= 0 -> d, d < 0 -> q - |d|; Montgomery: times 2^64 = 2^32 - 1 (mod q), exact for |d| < 2^31."""
    d = np.ascontiguousarray(d16, dtype=np.int16).astype(np.int64)
    m = np.abs(d).astype(np.uint64)
    if mont:
        m = (m << np.uint64(32)) - m
    return np.where((d < 0) & (m != 0), np.uint64(Q) - m, m)


def get_fhat_from_digits(d16, mont: bool = False) -> np.ndarray:
    """Witness::get_fhat (arith.rs:273-297) straight from the int16 digits: the host-side re-layout that replaces
    fetching f_coeff as 192-byte elements (4.7 MB instead of 19 MB over PCIe at the zkVM's size)."""
    d = np.ascontiguousarray(d16, dtype=np.int16).reshape(-1, D)
    out = np.zeros((3, d.shape[0], D), np.uint64)
    fq = digits_to_fq(d, mont)
    for j in range(3):
        out[j, :, 0::3] = fq[:, 8 * j : 8 * j + 8]
    return out


def get_fhat(f_coeff: np.ndarray, mont: bool = False) -> np.ndarray:
    """Witness::get_fhat re-layout (arith.rs:273-297), before truncate_lnze: (tau=3, n, 24) where
    fhat[j][i] carries coefficients 8j..8j+7 of f_coeff[i] as base-field scalars in component 0 of each slot."""
    fc = _as_u64(f_coeff, "f_coeff").reshape(-1, D)
    out = np.zeros((3, fc.shape[0], D), np.uint64)
    for j in range(3):
        out[j, :, 0::3] = fc[:, 8 * j : 8 * j + 8]
    return out


# ---- decomposition prover helpers ----------------------------------------------------------------------------------------
class LFDecompositionProver:
    """The two private helpers of latticefold/src/nifs/decomposition.rs the engine replaces."""

    @staticmethod
    def decompose_and_commit(scheme: AjtaiCommitmentScheme, wit_f_coeff, cm: Commitment, want_planes: bool = True,
                             side: int = 0):
        """decompose_witness (:162-167) + commit_witnesses (:178-201) in one engine call.
        Returns (wit_s, y_s): K witnesses (or None) and K commitments, y_0 by homomorphism.
        `side` (0 = accumulator, 1 = step witness; zk_latticefold.rs:60-71) says which resident plane buffer the
        engine fills, for LFFoldingProver.compute_f_0 later."""
        _raise(capi.lib().lat_ajtai_select_side(scheme._h, side))
        fc = _as_u64(wit_f_coeff, "f_coeff").reshape(-1, D)
        K, n, kappa = scheme.params.K, fc.shape[0], scheme._kappa
        pc = np.empty((K, n, D), np.uint64) if want_planes else None
        pf = np.empty((K, n, D), np.uint64) if want_planes else None
        cms = np.empty((K, kappa, D), np.uint64)
        cmv = _as_u64(cm.as_ref(), "cm")
        st = capi.lib().lat_ajtai_decompose_commit(scheme._h, _ptr(fc), n, _ptr(cmv), _ptr(pc), _ptr(pf), _ptr(cms))
        _raise(st, n, scheme._n)
        wit_s = [Witness(None, pf[k], pc[k], scheme.mont) for k in range(K)] if want_planes else None
        return wit_s, [Commitment(c, scheme.mont) for c in cms]

    @staticmethod
    def decompose_witness(scheme: AjtaiCommitmentScheme, wit: Witness) -> List[Witness]:
        """decompose_witness (:162-167)"""
        fc = _as_u64(wit.f_coeff, "f_coeff").reshape(-1, D)
        K, n = scheme.params.K, fc.shape[0]
        pc = np.empty((K, n, D), np.uint64)
        pf = np.empty((K, n, D), np.uint64)
        st = capi.lib().lat_ajtai_decompose_commit(scheme._h, _ptr(fc), n, None, _ptr(pc), _ptr(pf), None)
        _raise(st, n, scheme._n)
        return [Witness(None, pf[k], pc[k], scheme.mont) for k in range(K)]

    @staticmethod
    def commit_witnesses(scheme: AjtaiCommitmentScheme, wit_s: Sequence[Witness], cm: Commitment) -> List[Commitment]:
        """commit_witnesses (:178-201): matrix commits of wit_s[1..] in ONE batched launch, y_0 = cm - b_sum."""
        fs = np.stack([_as_u64(w.f, "f") for w in wit_s[1:]]) if len(wit_s) > 1 else None
        ys = scheme.commit_ntt_batch(fs) if fs is not None else []
        b = ntt_from_scalar(scheme.params.B_SMALL, scheme.mont)
        acc = Commitment.zeroed(scheme.kappa(), scheme.mont)
        for y in reversed(ys):
            acc = (acc + y) * b
        return [cm - acc] + ys


class LFFoldingProver:
    """The witness-side tail of latticefold/src/nifs/folding.rs the engine replaces."""

    @staticmethod
    def compute_f_0(scheme: AjtaiCommitmentScheme, rho_s, want_f_coeff: bool = True) -> Witness:
        """compute_f_0 (:258-268): f_0 = sum_i rho_i * f_i over the 2K planes left resident by the two
        decompose_and_commit calls (side 0 then side 1), followed by Witness::from_f's iCRT (arith.rs:299-313).
        rho_s: (2K, 24) CRT-form challenges.  Returns a Witness with f = f_0 and f_coeff = iCRT(f_0)."""
        rho = _as_u64(rho_s, "rho_s").reshape(-1, D)
        if rho.shape[0] != 2 * scheme.params.K:
            raise ValueError("need 2K challenges")
        f0 = np.empty((scheme._n, D), np.uint64)
        fc = np.empty((scheme._n, D), np.uint64) if want_f_coeff else None
        _raise(capi.lib().lat_ajtai_fold_witness(scheme._h, _ptr(rho), _ptr(f0), _ptr(fc)))
        return Witness(None, f0, fc, scheme.mont)


class FoldStep:
    """The GPU side of one IVC step's fold (zkvm/src/zk_latticefold.rs:37-102 as called from zkvm/src/main.rs:174-182),
    as the two blocking engine calls lat_ajtai_fold_step_begin / _finish.  The running accumulator witness stays on the
    device; per step only w_ccs goes up and commitments + int16 digits come down.

        fs = FoldStep(scheme); fs.set_accumulator(w_acc.f_coeff, acc_cm)        # initialize_accumulator, main.rs:306-344
        cm_i, ys_acc, ys_step, digits = fs.begin(w_ccs)      # commit + both decompositions' 2K commitments
        ... host: linearization / decomposition / folding sumchecks -> rho_s ...
        cm_0, f0_digits, f0, w_ccs0 = fs.finish(rho_s)       # compute_f_0, from_f, cm_0; f_0 is the next accumulator
    """

    def __init__(self, scheme: AjtaiCommitmentScheme):
        self.scheme = scheme

    def set_accumulator(self, f_coeff, cm_acc: Optional[Commitment] = None) -> None:
        fc = _as_u64(f_coeff, "f_coeff").reshape(-1, D)
        cm = _as_u64(cm_acc.as_ref(), "cm_acc") if cm_acc is not None else None
        _raise(capi.lib().lat_ajtai_set_accumulator(self.scheme._h, _ptr(fc), fc.shape[0], _ptr(cm)), fc.shape[0], self.scheme._n)

    def begin(self, w_ccs, cm_acc: Optional[Commitment] = None, want_digits: bool = True):
        s = self.scheme
        w = _as_u64(w_ccs, "w_ccs").reshape(-1, D)
        K, kappa, n = s.params.K, s._kappa, w.shape[0] * s.params.L
        d16 = np.empty((n, D), np.int16) if want_digits else None
        cm = np.empty((kappa, D), np.uint64)
        cms = np.empty((2, K, kappa, D), np.uint64)
        acc = _as_u64(cm_acc.as_ref(), "cm_acc") if cm_acc is not None else None
        st = capi.lib().lat_ajtai_fold_step_begin(s._h, _ptr(w), w.shape[0], _ptr(acc), _ptr(d16), _ptr(cm), _ptr(cms))
        _raise(st, n, s._n)
        ys = [[Commitment(c, s.mont) for c in side] for side in cms]
        return Commitment(cm, s.mont), ys[0], ys[1], d16

    def finish(self, rho_s, want_digits: bool = True, want_f0: bool = False, want_w_ccs: bool = False):
        s = self.scheme
        rho = _as_u64(rho_s, "rho_s").reshape(-1, D)
        if rho.shape[0] != 2 * s.params.K:
            raise ValueError("need 2K challenges")
        d16 = np.empty((s._n, D), np.int16) if want_digits else None
        f0 = np.empty((s._n, D), np.uint64) if want_f0 else None
        w0 = np.empty((s._n // s.params.L, D), np.uint64) if want_w_ccs else None
        cm0 = np.empty((s._kappa, D), np.uint64)
        _raise(capi.lib().lat_ajtai_fold_step_finish(s._h, _ptr(rho), _ptr(d16), _ptr(f0), _ptr(cm0), _ptr(w0)))
        return Commitment(cm0, s.mont), d16, f0, w0


def gadget_recompose(f, params: DecompositionParams = GoldiLocksDP, device: int = 0) -> np.ndarray:
    """GadgetRecompose for &[R] in CRT form (balanced_decomposition/mod.rs:177-190): rebuilds w_ccs from f
    (Witness::from_f / from_f_coeff, arith.rs:305,330)."""
    f = _as_u64(f, "f").reshape(-1, D)
    count = f.shape[0] // params.L
    out = np.empty((count, D), np.uint64)
    _raise(capi.lib().lat_ring_gadget_recompose(_ptr(f), count, params.log2_B, params.L, _ptr(out), 0, device))
    return out


def gadget_decompose(v, log2_b: int, L: int, mont: bool = False, device: int = 0) -> np.ndarray:
    """GadgetDecompose for &[R] in coefficient form (balanced_decomposition/mod.rs:163-175, coeff_form.rs:588-606):
    out[i*L + l] = limb l (base 2^log2_b, balanced) of v[i].  Raises DigitOverflow if a coefficient needs more than L
    limbs (the reference indexes out of bounds there, mod.rs:80)."""
    v = _as_u64(v, "v").reshape(-1, D)
    out = np.empty((v.shape[0] * L, D), np.uint64)
    _raise(capi.lib().lat_ring_gadget_decompose(_ptr(v), v.shape[0], log2_b, L, _ptr(out), 1 if mont else 0, device))
    return out


# ---- helpers ----------------------------------------------------------------------------------------------------------------
def _uniform(shape, seed: int) -> np.ndarray:
    """Uniform in [0, q) by rejection from numpy's PCG64 (not the reference's sampler; unpinned)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = rng.integers(0, 2**64, size=shape, dtype=np.uint64)
    bad = out >= np.uint64(Q)
    while bad.any():
        out[bad] = rng.integers(0, 2**64, size=int(bad.sum()), dtype=np.uint64)
        bad = out >= np.uint64(Q)
    return out


def _to_mont_np(x: np.ndarray) -> np.ndarray:
    """x * 2^64 mod q, vectorised: x*(2^32-1) = (x << 32) - x with 128-bit care via Python ints per element would be
    slow; do it with two 64-bit halves."""
    x = x.astype(np.uint64)
    lo = x & np.uint64(0xFFFFFFFF)
    hi = x >> np.uint64(32)
    # x * 2^64 = lo * 2^64 + hi * 2^96 = lo * (2^32 - 1) - hi   (mod q)
    t = lo * np.uint64(0xFFFFFFFF)  # < 2^64, exact
    # t - hi mod q
    res = np.where(t >= hi, t - hi, t + (np.uint64(Q) - hi))
    res = np.where(res >= np.uint64(Q), res - np.uint64(Q), res)
    return res


def ntt_negacyclic(a, inverse: bool = False, device: int = 0) -> np.ndarray:
    """Standalone negacyclic NTT over Z_q[X]/(X^d + 1), d = 2^k = a.shape[-1] (lat_ntt_negacyclic).  NOT on
    the drop-in path and absent from the reference -- see the header for the definition.  a: (..., d) uint64."""
    x = np.ascontiguousarray(a, dtype=np.uint64)
    if x.ndim < 1:
        raise ValueError("a: need at least one axis")
    d = x.shape[-1]
    log2_d = d.bit_length() - 1
    if d < 2 or (1 << log2_d) != d:
        raise ValueError("the last dimension must be a power of two >= 2")
    x = x.reshape(-1, d)
    out = np.empty_like(x)
    _raise(capi.lib().lat_ntt_negacyclic(_ptr(x), x.shape[0], log2_d, int(inverse), _ptr(out), device))
    return out.reshape(np.shape(a))


def to_mont(x) -> np.ndarray:
    return _to_mont_np(np.ascontiguousarray(x, dtype=np.uint64))


def from_mont(x) -> np.ndarray:
    """x * 2^-64 = x * 2^128 mod q (host helper for tests and serialisation; exact via Python ints)."""
    x = np.ascontiguousarray(x, dtype=np.uint64)
    flat = np.array([int(v) * MONT_RINV % Q for v in x.reshape(-1)], dtype=np.uint64)
    return flat.reshape(x.shape)
