"""Column-sharded Ajtai commitment across the GPUs of one box.

y = sum_j A_j f_j is a sum over columns, so the witness columns (and the matching w_ccs elements: decomposition
and CRT are per element) are split into contiguous blocks, one per rank; each rank commits its block against its
own column block of A and the kappa x 24 partial commitments (6 KB each) are exchanged once -- an all-gather over
NCCL/NVLink -- and summed mod q on every rank (NCCL has no modular reduction op).  One process per GPU,
torch.distributed for the plumbing.  The reference has no counterpart (single process, rayon only).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def _capi():
    from . import _capi as capi  # deferred: the CPU (gloo) tests use this module without the CUDA library

    return capi


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` units owned by `rank`; the first total % world ranks get one extra."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _all_agree(ok: bool, device: torch.device, group) -> bool:
    """True iff `ok` on EVERY rank (all-reduce MIN).  Every rank must call this the same number of times."""
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()) == 1)


class PeerExchange:
    """NVLink peer-memory state for the fused exchange + fold kernel (lat_commitment_exchange_dev): a receive buffer
    of 2 x world x max_words u64 and 2 x world flags per rank, allocated as torch symmetric memory so that every
    rank holds addresses of every peer's buffers.  One kernel per rank per exchange, no NCCL call on the data path.

    Construction is collective.  Use `PeerExchange.create`, which makes the decision "peer memory works" COLLECTIVELY:
    a rank on which a step fails still takes part in every all-reduce, so either all ranks get an exchange or all get
    None -- never a mix of ranks spinning in the peer kernel and ranks calling NCCL."""

    def __init__(self, device: torch.device, world: int, rank: int, max_words: int, recv, flags, handles):
        self.world, self.rank, self.max_words = world, rank, max_words
        self.recv, self.flags = recv, flags
        self.recv_ptrs = [int(p) for p in handles[0].buffer_ptrs]
        self.flag_ptrs = [int(p) for p in handles[1].buffer_ptrs]
        self._handles = handles
        self.epoch = 0

    @classmethod
    def create(cls, device: torch.device, world: int, rank: int, max_words: int, group=None):
        """(exchange or None, reason).  Phase 1 (local, no collective inside): import, peer access, allocation.
        Agreement.  Phase 2 (collective rendezvous; it fails or succeeds on all ranks alike, but is checked anyway).
        Agreement, which doubles as the barrier that orders "every rank's flags are zero" before the first exchange."""
        group = group if group is not None else dist.group.WORLD
        recv = flags = symm_mem = None
        why = ""
        try:
            import torch.distributed._symmetric_memory as symm_mem

            for peer in range(torch.cuda.device_count()):
                if peer != device.index and not torch.cuda.can_device_access_peer(device.index, peer):
                    raise RuntimeError(f"no peer access {device.index} -> {peer}")
            recv = symm_mem.empty(2 * world * max_words, dtype=torch.int64, device=device)
            flags = symm_mem.empty(2 * world, dtype=torch.int64, device=device)
            recv.zero_()
            flags.zero_()
        except Exception as e:  # pragma: no cover - depends on the box
            why = f"{type(e).__name__}: {e}"
        if not _all_agree(not why, device, group):
            return None, why or "a peer rank cannot set up peer memory"
        handles = None
        try:
            handles = (symm_mem.rendezvous(recv, group), symm_mem.rendezvous(flags, group))
            torch.cuda.synchronize(device)
        except Exception as e:  # pragma: no cover
            why = f"{type(e).__name__}: {e}"
        if not _all_agree(not why, device, group):
            return None, why or "rendezvous failed on a peer rank"
        return cls(device, world, rank, max_words, recv, flags, handles), ""


class ShardedAjtaiScheme:
    """Commitment over a column-sharded matrix.  `engine` is this rank's local engine front end and must offer
    witness_commit(w_local, cm) / commit_ntt(f_local, cm) / fold_partials(parts, out) / new_commitment(batch)
    (latticeum_b200.device.DeviceScheme on a GPU).

    exchange = "nccl": all-gather of the partials over NCCL + a fold kernel;
    exchange = "p2p" : the engine's fused peer-memory kernel (needs an engine with `exchange_partials` and CUDA peers);
    exchange = "auto": "p2p" when it can be set up, else "nccl"."""

    def __init__(self, engine, world: Optional[int] = None, rank: Optional[int] = None, group=None,
                 exchange: str = "nccl", max_batch: int = 32):
        self.engine = engine
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world, self.rank = world, rank
        self._gather = None
        self.peer: Optional[PeerExchange] = None
        self.exchange = "none" if world == 1 else "nccl"
        if world > 1 and exchange in ("p2p", "auto") and hasattr(engine, "exchange_partials"):
            # the same decision on every rank (see PeerExchange.create): a per-rank fallback would leave some ranks
            # spinning in the peer kernel while others sit in an NCCL all-gather
            self.peer, why = PeerExchange.create(engine.device, world, rank, max_batch * engine.kappa * 24, group)
            if self.peer is not None:
                self.exchange = "p2p"
            elif exchange == "p2p":
                raise RuntimeError(f"peer-memory exchange unavailable: {why}")
            else:
                self.exchange = f"nccl (p2p unavailable: {why.split(':')[0]})"

    def _exchange(self, partial: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return partial
        if self.peer is not None and partial.numel() <= self.peer.max_words:
            out = torch.empty_like(partial)
            return self.engine.exchange_partials(partial, out, self.peer)
        if self._gather is None or self._gather.shape[1:] != partial.shape or self._gather.device != partial.device:
            self._gather = torch.empty((self.world,) + tuple(partial.shape), dtype=partial.dtype, device=partial.device)
        # flat views: gloo's all-gather wants a 1-D output of world * numel; NCCL accepts the same
        dist.all_gather_into_tensor(self._gather.view(-1), partial.contiguous().view(-1), group=self.group)
        out = torch.empty_like(partial)
        return self.engine.fold_partials(self._gather, out)

    def witness_commit(self, w_local: torch.Tensor, partial: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Witness::from_w_ccs + commit on this rank's block of w_ccs; returns the full commitment on every rank."""
        if partial is None:
            partial = self.engine.new_commitment()
        self.engine.witness_commit(w_local, partial)
        return self._exchange(partial)

    def commit_ntt(self, f_local: torch.Tensor, partial: Optional[torch.Tensor] = None) -> torch.Tensor:
        if partial is None:
            partial = self.engine.new_commitment(f_local.shape[0] if f_local.dim() == 3 else 1)
        self.engine.commit_ntt(f_local, partial)
        return self._exchange(partial)


    # -- the fold step, column-sharded ("the partition applies to the decompose-and-commit entry points too") --
    def decompose_commit(self, f_coeff_local: torch.Tensor, cm_full: torch.Tensor, side: int = 0,
                         cms: Optional[torch.Tensor] = None) -> torch.Tensor:
        """decompose_witness + commit_witnesses (latticefold/src/nifs/decomposition.rs:162-201) on this rank's block of
        f_coeff: the K planes of the block stay resident on this rank (for `fold_witness`), the K-1 partial commitments
        of planes 1..K-1 are exchanged in ONE go ((K-1) x kappa x 24 words) and y_0 = cm - sum 2^k y_k is derived from
        the exchanged totals and the FULL commitment `cm_full`.  Returns (K, kappa, 24) on every rank."""
        eng = self.engine
        K = eng.K
        if cms is None:
            cms = eng.new_commitment(K)
        eng.decompose_commit(f_coeff_local, None if self.world > 1 else cm_full, cms, side=side)
        if self.world == 1:
            return cms
        if K > 1:
            cms[1:].copy_(self._exchange(cms[1:]))
        return eng.y0(cm_full, cms)

    def fold_witness(self, rho: torch.Tensor):
        """compute_f_0 + iCRT on this rank's columns: per element, so no exchange (folding.rs:258-268, arith.rs:299-313).
        Returns this rank's blocks (f0_local, f0_coeff_local)."""
        return self.engine.fold_witness(rho)


class ShardedCommitPipeline:
    """A stream of per-step sharded commitments with HOST buffers: `submit(w_host)` queues, on this rank, the upload of
    its block of the step's w_ccs, witness_commit + the partial exchange and the return of the full commitment, and
    returns a ticket; `wait(ticket)` returns the pinned host tensor holding that step's commitment.  `depth` steps are
    in flight, so the PCIe transfers of neighbouring steps hide under the kernels.  Every rank must submit the same
    sequence of steps (the exchange is collective).

    With the fused peer-memory exchange the compute stream carries nothing but kernels: the upload (copy stream) is
    followed by a copy of the ticket that the witness kernel polls, and the exchange kernel writes the commitment and
    then the ticket into pinned host memory that `wait` polls -- no event waits, so consecutive steps overlap on the
    device (DeviceScheme.set_step_overlap).  With NCCL, or without an engine that offers the gated call, uploads and
    downloads are ordered with events instead.  On a CPU device (gloo tests) the same calls run synchronously."""

    def __init__(self, sharded: ShardedAjtaiScheme, w_len_local: int, depth: int = 4):
        self.sharded, self.depth = sharded, depth
        eng = sharded.engine
        dev = getattr(eng, "device", torch.device("cpu"))
        self.cuda = dev.type == "cuda"
        self.chain = self.cuda and sharded.peer is not None and hasattr(eng, "witness_commit_gated")
        # preferred: the engine's own pipelined submit/wait made sharding-aware (lat_ajtai_set_peers) -- one C call per
        # step instead of several Python-level launches, which matters once 8 ranks share the host's cores
        self.native = self.chain and hasattr(eng, "scheme") and depth <= _capi().LAT_PIPELINE_DEPTH
        self.next_ticket = 0
        self.slots = []
        for _ in range(depth):
            sl = {
                "w": torch.empty((w_len_local, 24), dtype=torch.int64, device=dev),
                "partial": eng.new_commitment(),
                "cm_host": torch.empty((eng.kappa, 24), dtype=torch.int64),
                "ticket": None,
            }
            if self.cuda:
                sl["cm_host"] = sl["cm_host"].pin_memory()
                sl["uploaded"], sl["computed"], sl["done"] = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            if self.chain:
                sl["out"] = eng.new_commitment()
                sl["ready"] = torch.full((1,), -1, dtype=torch.int64, device=dev)
                sl["ticket_host"] = torch.zeros((1,), dtype=torch.int64).pin_memory()
                sl["done_host"] = torch.full((1,), -1, dtype=torch.int64).pin_memory()
            self.slots.append(sl)
        if self.cuda:
            self.up, self.down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        if self.chain:
            eng.set_step_overlap(True)
            torch.cuda.synchronize(dev)

    def submit(self, w_host: torch.Tensor) -> int:
        sl = self.slots[self.next_ticket % self.depth]
        if sl["ticket"] is not None:
            raise RuntimeError(f"pipeline full: wait for ticket {sl['ticket']} first")
        tk = self.next_ticket
        eng = self.sharded.engine
        if not self.cuda:
            sl["w"].copy_(w_host)
            sl["cm_host"].copy_(self.sharded.witness_commit(sl["w"], sl["partial"]))
        elif self.native:
            import ctypes as C

            L, peer, h = _capi().lib(), self.sharded.peer, eng.scheme._h
            eng.bind_stream()
            n = peer.world
            st = L.lat_ajtai_set_peers(h, peer.rank, n, (C.c_uint64 * n)(*peer.recv_ptrs), (C.c_uint64 * n)(*peer.flag_ptrs),
                                       peer.epoch + 1)
            et = C.c_uint64(0)
            st = st or L.lat_ajtai_submit_w_ccs(h, w_host.data_ptr(), w_host.shape[0], sl["cm_host"].data_ptr(), C.byref(et))
            if st:
                raise RuntimeError(_capi().last_error())
            peer.epoch += 1
            sl["engine_ticket"] = et.value
        elif self.chain:
            sl["ticket_host"][0] = tk
            with torch.cuda.stream(self.up):  # the slot's previous user has been waited for: its buffers are free
                sl["w"].copy_(w_host, non_blocking=True)
                sl["ready"].copy_(sl["ticket_host"], non_blocking=True)
            eng.witness_commit_gated(sl["w"], sl["partial"], sl["ready"], tk)
            eng.exchange_partials(sl["partial"], sl["out"], self.sharded.peer, report=(sl["cm_host"], sl["done_host"], tk))
        else:
            compute = torch.cuda.current_stream()
            with torch.cuda.stream(self.up):
                sl["w"].copy_(w_host, non_blocking=True)
                sl["uploaded"].record(self.up)
            compute.wait_event(sl["uploaded"])
            cm = self.sharded.witness_commit(sl["w"], sl["partial"])
            sl["computed"].record(compute)
            self.down.wait_event(sl["computed"])
            with torch.cuda.stream(self.down):
                sl["cm_host"].copy_(cm, non_blocking=True)
                sl["done"].record(self.down)
            cm.record_stream(self.down)
        sl["ticket"] = tk
        self.next_ticket += 1
        return tk

    def wait(self, ticket: int) -> torch.Tensor:
        sl = self.slots[ticket % self.depth]
        if sl["ticket"] != ticket:
            raise KeyError(f"no such ticket in flight: {ticket}")
        if self.native:
            if _capi().lib().lat_ajtai_wait(self.sharded.engine.scheme._h, sl["engine_ticket"]):
                raise RuntimeError(_capi().last_error())
        elif self.chain:
            done, spins = sl["done_host"], 0
            while int(done[0]) != ticket:  # written by the exchange kernel after the commitment
                spins += 1
                if spins % 200000 == 0 and torch.cuda.current_stream().query() and int(done[0]) != ticket:
                    raise RuntimeError("stream idle but the step never reported completion")
        elif self.cuda:
            sl["done"].synchronize()
        sl["ticket"] = None
        return sl["cm_host"]
