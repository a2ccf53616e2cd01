"""Device-resident front end of the engine for callers that already hold their data in HBM (torch tensors).

PyTorch is plumbing here (device memory, streams, torch.distributed); the kernels are the engine's own
(liblattice_ajtai.so, `_dev` entry points of include/lattice_ajtai.h).  Tensors are torch.int64 holding the
u64 bit patterns of ring elements, shape (..., 24).  All calls are asynchronous on torch's current stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi as capi
from .scheme import AjtaiCommitmentScheme, D, _raise


class DeviceScheme:
    """Wraps an AjtaiCommitmentScheme handle; work is queued on torch's current CUDA stream."""

    def __init__(self, scheme: AjtaiCommitmentScheme):
        if not torch.cuda.is_available():
            raise RuntimeError("latticeum_b200 needs a CUDA device; there is no CPU fallback")
        self.scheme = scheme
        self.device = torch.device("cuda", scheme.device)
        self.kappa, self.n, self.L, self.K = scheme._kappa, scheme._n, scheme.params.L, scheme.params.K
        self._bound_stream = None
        self.bind_stream()

    def bind_stream(self) -> None:
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s == 0:
            s = 1  # cudaStreamLegacy: the C ABI reserves NULL for "the handle's own stream"
        if s != self._bound_stream:
            _raise(capi.lib().lat_ajtai_set_stream(self.scheme._h, C.c_void_p(s)))
            self._bound_stream = s

    def _check(self, t: torch.Tensor, name: str) -> int:
        if t.device != self.device or t.dtype != torch.int64 or not t.is_contiguous() or t.shape[-1] != D:
            raise ValueError(f"{name}: need a contiguous int64 CUDA tensor (..., 24) on {self.device}")
        return t.data_ptr()

    def to_device(self, a: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(self.device)

    @staticmethod
    def to_numpy(t: torch.Tensor) -> np.ndarray:
        return t.cpu().numpy().view(np.uint64)

    def new_commitment(self, batch: int = 1) -> torch.Tensor:
        shape = (self.kappa, D) if batch == 1 else (batch, self.kappa, D)
        return torch.empty(shape, dtype=torch.int64, device=self.device)

    # -- Witness::from_w_ccs + commit, device resident (ZKVM/main.rs:348-367) ---------------------------------------
    def witness_commit(self, w_ccs: torch.Tensor, cm: torch.Tensor, f: torch.Tensor = None,
                       f_coeff: torch.Tensor = None) -> torch.Tensor:
        self.bind_stream()
        st = capi.lib().lat_ajtai_witness_from_w_ccs_dev(
            self.scheme._h, self._check(w_ccs, "w_ccs"), w_ccs.shape[0],
            self._check(f_coeff, "f_coeff") if f_coeff is not None else None,
            self._check(f, "f") if f is not None else None, self._check(cm, "cm"))
        _raise(st, w_ccs.shape[0] * self.L, self.n)
        return cm

    def witness_commit_gated(self, w_ccs: torch.Tensor, cm: torch.Tensor, ready_flag: torch.Tensor, ready_value: int) -> torch.Tensor:
        """witness_commit whose kernel first polls `ready_flag` (a 1-element int64 CUDA tensor) until it holds
        `ready_value`: the caller uploads w_ccs on another stream and copies the value there afterwards, so the compute
        stream needs no event wait (lat_ajtai_witness_from_w_ccs_gated_dev)."""
        self.bind_stream()
        st = capi.lib().lat_ajtai_witness_from_w_ccs_gated_dev(self.scheme._h, self._check(w_ccs, "w_ccs"), w_ccs.shape[0],
                                                               self._check(cm, "cm"), ready_flag.data_ptr(), ready_value)
        _raise(st, w_ccs.shape[0] * self.L, self.n)
        return cm

    def commit_ntt(self, f: torch.Tensor, cm: torch.Tensor) -> torch.Tensor:
        self.bind_stream()
        if f.dim() == 3:
            st = capi.lib().lat_ajtai_commit_ntt_batch_dev(self.scheme._h, self._check(f, "f"), f.shape[0], f.shape[1],
                                                           self._check(cm, "cm"))
            _raise(st, f.shape[1], self.n)
        else:
            st = capi.lib().lat_ajtai_commit_ntt_dev(self.scheme._h, self._check(f, "f"), f.shape[0], self._check(cm, "cm"))
            _raise(st, f.shape[0], self.n)
        return cm

    def decompose_commit(self, f_coeff: torch.Tensor, cm, cms: torch.Tensor, planes_f: torch.Tensor = None,
                         side: int = None) -> torch.Tensor:
        """decompose_witness + commit_witnesses (latticefold/src/nifs/decomposition.rs:162-201) on device tensors.
        cms: (K, kappa, 24); cms[1:] = A * plane_k, cms[0] = y_0 by homomorphism from `cm` -- or left untouched when
        cm is None (column-sharded callers finish with `y0` after exchanging cms[1:]).  `side` (0 / 1) selects which
        resident plane buffer the call fills for `fold_witness`."""
        self.bind_stream()
        if side is not None:
            _raise(capi.lib().lat_ajtai_select_side(self.scheme._h, side))
        st = capi.lib().lat_ajtai_decompose_commit_dev(
            self.scheme._h, self._check(f_coeff, "f_coeff"), f_coeff.shape[0], self._check(cm, "cm") if cm is not None else None,
            None, self._check(planes_f, "planes_f") if planes_f is not None else None, self._check(cms, "cms"))
        _raise(st, f_coeff.shape[0], self.n)
        return cms

    def y0(self, cm: torch.Tensor, cms: torch.Tensor) -> torch.Tensor:
        """cms[0] = cm - sum_{k>=1} 2^k cms[k]  (decomposition.rs:189-197), in place on device."""
        s = torch.cuda.current_stream(self.device).cuda_stream or 1
        _raise(capi.lib().lat_commitment_y0_dev(self._check(cm, "cm"), self._check(cms, "cms"), cms.shape[0], self.kappa,
                                                C.c_void_p(s)))
        return cms

    def fold_witness(self, rho: torch.Tensor, f0: torch.Tensor = None, f0_coeff: torch.Tensor = None):
        """compute_f_0 over the 2K planes left resident by the two decompose_commit calls + Witness::from_f's iCRT
        (latticefold/src/nifs/folding.rs:258-268, arith.rs:299-313).  Returns (f0, f0_coeff) device tensors."""
        self.bind_stream()
        if f0 is None:
            f0 = torch.empty((self.n, D), dtype=torch.int64, device=self.device)
        if f0_coeff is None:
            f0_coeff = torch.empty((self.n, D), dtype=torch.int64, device=self.device)
        _raise(capi.lib().lat_ajtai_fold_witness_dev(self.scheme._h, self._check(rho, "rho"), self._check(f0, "f0"),
                                                     self._check(f0_coeff, "f0_coeff")))
        return f0, f0_coeff

    def fold_partials(self, parts: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """Sum of partial commitments mod q (parts: (world, ...), out: (...))."""
        words = out.numel()
        s = torch.cuda.current_stream(self.device).cuda_stream or 1
        _raise(capi.lib().lat_commitment_sum_dev(parts.data_ptr(), parts.shape[0], words, out.data_ptr(), C.c_void_p(s)))
        return out

    def exchange_partials(self, partial: torch.Tensor, out: torch.Tensor, peer, report=None) -> torch.Tensor:
        """Fused NVLink exchange + mod-q fold of this rank's partial commitment with all peers (`peer` is a
        latticeum_b200.sharded.PeerExchange).  Asynchronous on torch's current stream.  `report` = (cm_host, done_host,
        value): pinned host tensors the kernel also writes the result and then `value` to (the host polls done_host)."""
        peer.epoch += 1
        n = peer.world
        recv = (C.c_uint64 * n)(*peer.recv_ptrs)
        flags = (C.c_uint64 * n)(*peer.flag_ptrs)
        s = torch.cuda.current_stream(self.device).cuda_stream or 1
        cm_host, done_host, value = (report[0].data_ptr(), report[1].data_ptr(), report[2]) if report else (None, None, 0)
        _raise(capi.lib().lat_commitment_exchange_report_dev(partial.data_ptr(), partial.numel(), peer.rank, n, recv, flags,
                                                             peer.epoch, out.data_ptr(), cm_host, done_host, value, C.c_void_p(s)))
        return out

    def synchronize(self) -> None:
        _raise(capi.lib().lat_ajtai_synchronize(self.scheme._h))

    def set_step_overlap(self, on: bool) -> None:
        """Let the witness kernel of a step start while the previous step's matrix-vector kernel is draining
        (lat_ajtai_set_step_overlap).  Only for callers whose w_ccs tensors are complete before the previous
        witness_commit was enqueued and whose stream carries nothing else between two calls (see the header)."""
        _raise(capi.lib().lat_ajtai_set_step_overlap(self.scheme._h, int(on)))

    # -- diagnostics ---------------------------------------------------------------------------------------------------
    def set_profiling(self, on: bool) -> None:
        _raise(capi.lib().lat_ajtai_set_profiling(self.scheme._h, int(on)))

    def mac_profile(self):
        s, c = C.c_double(), C.c_uint64()
        _raise(capi.lib().lat_ajtai_mac_profile(self.scheme._h, C.byref(s), C.byref(c)))
        return s.value, c.value
