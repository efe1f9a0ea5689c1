"""Multi-GPU plumbing for the data-parallel path (one process per GPU, torch.distributed).

The reference has no distributed code at all; the hot path shards by scene (training)
and by point block (dense evaluation) with NO data-path collective.  The only exchange is the
gradient all-reduce of the replicated weights (2 550 881 fp32 = 10.2 MB).

``GradReducer`` keeps ONE persistent flat fp32 buffer per bucket.  Buckets follow the order in which the
backward pass produces gradients (the decoder first: its all-reduce overlaps the whole encoder backward; then the
encoder); each bucket is packed with one multi-tensor copy and its all-reduce (``ReduceOp.AVG`` on NCCL) is launched
from the post-accumulate-grad hook of its last parameter.  After the reduction ``p.grad`` is re-pointed at views of the flat buffer: no
``torch.cat``, no copy back, no division kernel.

BatchNorm: each rank normalises with its own shard's batch statistics (like DDP without SyncBN);
parity of a DP step is therefore defined against a per-shard reference, see DESIGN.md."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

# bucket = parameters whose name starts with one of the prefixes; listed in backward (gradient-ready) order; everything
# else forms the last bucket.  Measured at 2 x B200 (config 2, 20 steps): decoder bucket + ONE encoder bucket 8.565 ms/step,
# decoder + three encoder buckets 8.62 ms (1 GPU: 8.52 ms) -- the encoder's 7 MB all-reduce is ~40 us, less than what the
# extra collectives cost the kernels they overlap with, so the default is two buckets.
DEFAULT_BUCKETS: Tuple[Tuple[str, ...], ...] = (("fc_",),)
FINE_BUCKETS: Tuple[Tuple[str, ...], ...] = (
    ("fc_",),
    ("ifnet_feature_extractor.conv_3", "ifnet_feature_extractor.conv3", "ifnet_feature_extractor.conv_2", "ifnet_feature_extractor.conv2"),
    ("ifnet_feature_extractor.conv_1", "ifnet_feature_extractor.conv1", "ifnet_feature_extractor.conv_0", "ifnet_feature_extractor.conv0"),
)


def _view_like(flat: torch.Tensor, off: int, p: torch.Tensor) -> torch.Tensor:
    """View of flat[off : off + p.numel()] with p's shape AND p's (dense, possibly permuted) strides."""
    n = p.numel()
    order = sorted(range(p.dim()), key=lambda d: (-p.stride(d), d))
    expect, dense = 1, True
    for d in reversed(order):
        if p.shape[d] != 1 and p.stride(d) != expect:
            dense = False
        expect *= p.shape[d]
    if not dense:
        return flat[off:off + n].view(p.shape)
    return flat.as_strided(p.shape, p.stride(), flat.storage_offset() + off)     # literally the parameter's strides


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.index = {id(p): i for i, p in enumerate(params)}
        n = sum(p.numel() for p in params)
        dev, dt = params[0].device, params[0].dtype
        self.flat = torch.zeros((n,), device=dev, dtype=dt)
        self.views, self.offsets, off = [], [], 0
        for p in params:
            # same strides as the parameter (the encoder's weights are channels_last_3d): fused optimisers require the
            # gradient's layout to match the parameter's
            self.views.append(_view_like(self.flat, off, p))
            self.offsets.append(off)
            off += p.numel()
        self.fired = set()
        self.work = None
        self.dirty = False      # a gradient changed after the bucket was packed (gradient accumulation): reduce again


class GradReducer:
    """Averages ``module``'s gradients over the process group.  Call ``allreduce()`` once after the last
    ``backward()`` of a step.  ``weight`` is this rank's share of the global batch (default 1/world: equal
    shards); with unequal shards (``shard_range`` with a remainder) pass ``n_local / n_global`` so that the result
    is the gradient of the global-batch mean."""

    def __init__(self, module: torch.nn.Module, early_prefixes: Optional[Sequence[str]] = None, group=None,
                 buckets: Optional[Sequence[Sequence[str]]] = None, weight: Optional[float] = None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.weight = weight
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if buckets is None:
            buckets = DEFAULT_BUCKETS if early_prefixes is None else (tuple(early_prefixes),)
        groups: List[List[torch.nn.Parameter]] = [[] for _ in range(len(buckets) + 1)]
        for n, p in named:
            for bi, prefixes in enumerate(buckets):
                if n.startswith(tuple(prefixes)):
                    groups[bi].append(p)
                    break
            else:
                groups[-1].append(p)     # everything else (conv_in, conv_in_bn, ...): last bucket
        self.buckets = [_Bucket(g) for g in groups if g]
        self._of: Dict[int, _Bucket] = {id(p): b for b in self.buckets for p in b.params}
        self._avg = dist.get_backend(group) == "nccl" and weight is None
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for b in self.buckets for p in b.params]
        # kept for callers written against the two-bucket version
        self.early = self.buckets[0].params
        self.late = [p for b in self.buckets[1:] for p in b.params]

    # -- bucket helpers ---------------------------------------------------------------------------
    def _launch(self, b: _Bucket):
        for i, p in enumerate(b.params):      # a module may change its parameters' memory format after construction
            if b.views[i].stride() != p.stride():      # (the encoder converts itself to channels_last_3d on first use)
                b.views[i] = _view_like(b.flat, b.offsets[i], p)
        grads = [p.grad for p in b.params]
        if any(g is None for g in grads):      # parameters without a gradient contribute zeros (every rank still
            b.flat.zero_()                     # runs the collective, so the replicas must agree on which they are)
        dst_l, src_l = [], []
        for p, g, v in zip(b.params, grads, b.views):
            if g is not None and g.data_ptr() != v.data_ptr():
                dst_l.append(v)
                src_l.append(g)
        if dst_l:
            torch._foreach_copy_(dst_l, src_l)
        if self.weight is not None:
            b.flat.mul_(float(self.weight))
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
        b.dirty = False

    def _on_grad(self, param):
        b = self._of[id(param)]
        if id(param) in b.fired:          # second backward before allreduce(): the packed copy is stale
            b.dirty = True
            return
        b.fired.add(id(param))
        if len(b.fired) == len(b.params) and b.work is None:
            # a gradient that already lives in the bucket buffer (zero_grad(set_to_none=False) after a previous step)
            # would be reduced in place while a later micro-batch may still accumulate into it: leave such buckets
            # to allreduce()
            if not any(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(b.params, b.views)):
                self._launch(b)

    # -- public -----------------------------------------------------------------------------------
    def allreduce(self):
        """Finishes the buckets launched from the hooks, reduces the rest, and points every ``p.grad`` at its
        (averaged) slice of the bucket buffer."""
        for b in self.buckets:
            if b.work is not None and b.dirty:
                b.work.wait()
                b.work = None
            if b.work is None:
                if all(p.grad is None for p in b.params):
                    b.fired.clear()
                    continue          # nothing produced a gradient for this bucket on any rank's replica of the graph
                self._launch(b)
        for b in self.buckets:
            if b.work is None:
                continue
            b.work.wait()
            if not self._avg and self.weight is None:
                b.flat.div_(self.world)
            for p, v in zip(b.params, b.views):
                p.grad = v
            b.work, b.dirty = None, False
            b.fired.clear()


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
