"""Double-buffered host->device staging of ray batches (the seam to the reference's data manager,
data/PreSight/my_datamanager.py:257-285, which hands the trainer one pinned CPU batch per step).

    pf = DevicePrefetcher(device, keys)
    pf.push(host_batch)                 # starts the copy of batch k+1 on a copy stream ...
    batch = pf.pop()                    # ... while the compute stream works on batch k

Two static device buffer sets are reused alternately (no allocator traffic); a set is overwritten only after the
compute stream has passed the `release` point of the step that consumed it.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor


class DevicePrefetcher:
    def __init__(self, device: torch.device, keys: Sequence[str]) -> None:
        self.device, self.keys = device, tuple(keys)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.bufs: List[Optional[Dict[str, Tensor]]] = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]       # copy of set j finished (copy stream)
        self.free = [None, None]                                    # compute done with set j (compute stream)
        self._in, self._out = 0, 0

    def push(self, host_batch: Dict[str, Tensor]) -> None:
        """Start copying `host_batch` (pinned CPU tensors) into the next buffer set."""
        j = self._in % 2
        self._in += 1
        if self.bufs[j] is None:
            self.bufs[j] = {k: torch.empty(host_batch[k].shape, dtype=host_batch[k].dtype, device=self.device)
                            for k in self.keys}
        with torch.cuda.stream(self.copy_stream):
            if self.free[j] is not None:
                self.copy_stream.wait_event(self.free[j])
            for k in self.keys:
                self.bufs[j][k].copy_(host_batch[k], non_blocking=True)
            self.ready[j].record(self.copy_stream)

    def pop(self) -> Dict[str, Tensor]:
        """Device batch of the oldest pushed copy; the current stream waits for that copy."""
        j = self._out % 2
        self._out += 1
        torch.cuda.current_stream(self.device).wait_event(self.ready[j])
        return self.bufs[j]

    def release(self) -> None:
        """Call after enqueueing the step that used the batch returned by the last pop()."""
        j = (self._out - 1) % 2
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[j] = ev
