"""Host -> device input staging for the training / evaluation loops.

The reference's trainers read batches from a ``DataLoader(pin_memory=True)`` and copy them with ``.to(device)``
inside the step (trainer/trainer_ifnet.py:47-56): the 36.7 MB of a batch-4 step (128^3 grids + 50 k points + labels)
then sit on the critical path in front of the first kernel.  ``HostPrefetcher`` issues the copy of batch i+1 on a side
stream while batch i is being computed; the consumer waits on an event, not on the host."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch


class HostPrefetcher:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("svr_b200: HostPrefetcher needs a CUDA device; there is no CPU path")
        self.stream = torch.cuda.Stream(self.device)

    def issue(self, host_tensors: Sequence[torch.Tensor]) -> Tuple[list, torch.cuda.Event]:
        """Starts the asynchronous copies of pinned host tensors; returns a handle for ``wait``."""
        for t in host_tensors:
            if not t.is_pinned():
                raise RuntimeError("svr_b200: HostPrefetcher expects pinned host tensors (tensor.pin_memory())")
        with torch.cuda.stream(self.stream):
            dev = [t.to(self.device, non_blocking=True) for t in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def wait(self, handle) -> list:
        """Makes the current stream wait for the copies of ``handle`` and hands the device tensors over to it."""
        dev, ev = handle
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)
        return dev
