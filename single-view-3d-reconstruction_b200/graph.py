"""Whole-step CUDA graph.

A training step of the IF-Net path is ~160 kernel launches of 5-500 us: on one B200 the host needs 60-70 % of the step's
8 ms just to enqueue it (Python, autograd, ctypes), and with eight ranks sharing one host's cores it becomes the bound
(step 8.0 -> 9.0 ms at N = 8).  Everything the step launches -- the library's kernels (raw current stream), cuDNN, the
side-stream overlaps (forked and joined through events), the stream-ordered scratch allocations, the gradient
all-reduce -- is capturable, so the step is recorded once and replayed."""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    """``step = GraphedStep(fn, inputs)``; ``out = step(*new_inputs)`` replays ``fn`` on static copies of the inputs.

    ``fn(*tensors)`` must be a complete step on fixed shapes (forward, backward, optimiser step with
    ``capturable=True``) that does not synchronise; it is run ``warmup`` times eagerly on a side stream first
    (cuDNN autotuning, lazily created state), then captured.  The tensors it returns are static: copy them out
    (``.clone()`` / an async D2H copy) before the next replay.  Passing the very tensors given at construction skips
    the input copy."""

    def __init__(self, fn: Callable, inputs: Sequence[torch.Tensor], warmup: int = 3):
        self.static_in = list(inputs)
        dev = self.static_in[0].device
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
