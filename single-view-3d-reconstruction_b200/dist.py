"""Multi-GPU plumbing for the data-parallel path (one process per GPU, torch.distributed).

The reference has no distributed code at all; the hot path shards by scene (training)
and by point block (dense evaluation) with NO data-path collective.  The only exchange is the
gradient all-reduce of the replicated weights (2 550 881 fp32 = 10.2 MB): two buckets, the decoder
bucket is launched from a gradient hook as soon as the fused decoder backward has produced its
gradients so that it overlaps the (torch/cuDNN) encoder backward.

BatchNorm: each rank normalises with its own shard's batch statistics (like DDP without SyncBN);
parity of a DP step is therefore defined against a per-shard reference, see DESIGN.md."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class GradReducer:
    """Averages ``module``'s gradients over the process group after ``backward``."""

    def __init__(self, module: torch.nn.Module, early_prefixes: Sequence[str] = ("fc_",), group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        self.early = [p for n, p in named if n.startswith(tuple(early_prefixes))]
        self.late = [p for n, p in named if not n.startswith(tuple(early_prefixes))]
        self._pending = 0
        self._early_work = None
        self._early_flat: Optional[torch.Tensor] = None
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_early_grad) for p in self.early]

    # -- bucket helpers ---------------------------------------------------------------------------
    @staticmethod
    def _flatten(params: List[torch.Tensor]) -> torch.Tensor:
        return torch.cat([p.grad.reshape(-1) for p in params])

    @staticmethod
    def _scatter_back(flat: torch.Tensor, params: List[torch.Tensor]):
        off = 0
        for p in params:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n

    def _on_early_grad(self, _param):
        self._pending += 1
        if self._pending == len(self.early):
            self._early_flat = self._flatten(self.early)
            self._early_work = dist.all_reduce(self._early_flat, group=self.group, async_op=True)

    # -- public -----------------------------------------------------------------------------------
    def allreduce(self):
        """Call once after ``loss.backward()``: finishes the early bucket, reduces the late one."""
        late_flat = None
        if self.late and all(p.grad is not None for p in self.late):
            late_flat = self._flatten(self.late)
            late_work = dist.all_reduce(late_flat, group=self.group, async_op=True)
        if self._early_work is None and self.early and all(p.grad is not None for p in self.early):
            # hooks did not fire (e.g. gradients accumulated manually): reduce now
            self._early_flat = self._flatten(self.early)
            self._early_work = dist.all_reduce(self._early_flat, group=self.group, async_op=True)
        if self._early_work is not None:
            self._early_work.wait()
            self._scatter_back(self._early_flat.div_(self.world), self.early)
        if late_flat is not None:
            late_work.wait()
            self._scatter_back(late_flat.div_(self.world), self.late)
        self._pending, self._early_work, self._early_flat = 0, None, None


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) block of ``n_items`` for ``rank`` (scenes for training,
    first-axis lattice slabs or scenes for dense evaluation)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_scenes(batch: dict, rank: int, world: int) -> dict:
    """Slice every tensor of a batch dict along dim 0 to this rank's scenes."""
    n = next(v.shape[0] for v in batch.values() if torch.is_tensor(v))
    b, e = shard_range(n, rank, world)
    return {k: (v[b:e] if torch.is_tensor(v) and v.shape[0] == n else v) for k, v in batch.items()}
