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


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` units owned by `rank`; the first total % world ranks get one extra."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedAjtaiScheme:
    """Commitment over a column-sharded matrix.  `engine` is this rank's local engine front end and must offer
    witness_commit(w_local, cm) / commit_ntt(f_local, cm) / fold_partials(parts, out) / new_commitment(batch)
    (latticeum_b200.device.DeviceScheme on a GPU)."""

    def __init__(self, engine, world: Optional[int] = None, rank: Optional[int] = None, group=None):
        self.engine = engine
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world, self.rank = world, rank
        self._gather = None

    def _exchange(self, partial: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return partial
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
