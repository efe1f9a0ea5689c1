"""Host -> device input staging for the training / evaluation loops.

The reference's trainers read batches from a ``DataLoader(pin_memory=True)`` and copy them with ``.to(device)``
inside the step (trainer/trainer_ifnet.py:47-56): the 36.7 MB of a batch-4 step (128^3 grids + 50 k points + labels)
then sit on the critical path in front of the first kernel.  ``HostPrefetcher`` issues the copy of batch i+1 on a side
stream while batch i is being computed; the consumer waits on an event, not on the host.

The device side is a ring of ``depth`` persistent staging buffers (no per-step allocation: a caching-allocator block
that is handed from the copy stream to the compute stream cannot be reused while the host runs ahead, and the
resulting cudaMalloc/cudaFree calls serialise the device).  ``issue`` makes the copy stream wait for the compute work
enqueued so far, so slot k is only overwritten after the step that consumed it ``depth`` batches ago -- enqueue the
step that uses a batch BEFORE issuing the batch ``depth`` positions later (the natural loop order:
``cur = wait(h); h = issue(next_batch); step(cur)`` with depth 2)."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch


class HostPrefetcher:
    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("svr_b200: HostPrefetcher needs a CUDA device; there is no CPU path")
        if depth < 2:
            raise ValueError("svr_b200: HostPrefetcher needs at least two staging slots")
        self.stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.slots = [None] * depth          # per slot: list of device tensors
        self.n = 0

    def _buffers(self, slot: int, host_tensors: Sequence[torch.Tensor]):
        bufs = self.slots[slot]
        if (bufs is None or len(bufs) != len(host_tensors)
                or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, host_tensors))):
            bufs = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host_tensors]
            self.slots[slot] = bufs
        return bufs

    def issue(self, host_tensors: Sequence[torch.Tensor]) -> Tuple[list, torch.cuda.Event]:
        """Starts the asynchronous copies of pinned host tensors into the next staging slot; returns a handle for
        ``wait``.  The slot's previous contents must no longer be needed by work enqueued AFTER this call."""
        for t in host_tensors:
            if not t.is_pinned():
                raise RuntimeError("svr_b200: HostPrefetcher expects pinned host tensors (tensor.pin_memory())")
        slot = self.n % self.depth
        self.n += 1
        bufs = self._buffers(slot, host_tensors)
        self.stream.wait_event(torch.cuda.current_stream(self.device).record_event())   # consumers of this slot are enqueued
        with torch.cuda.stream(self.stream):
            for b, t in zip(bufs, host_tensors):
                b.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return bufs, ev

    def wait(self, handle) -> list:
        """Makes the current stream wait for the copies of ``handle``; returns the device tensors (staging buffers:
        valid until ``depth`` further batches have been issued)."""
        dev, ev = handle
        torch.cuda.current_stream(self.device).wait_event(ev)
        return list(dev)
