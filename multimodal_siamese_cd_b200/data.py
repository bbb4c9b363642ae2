"""Host -> device staging for the training loops.

The reference moves every batch with a blocking `.to(device)` right before the forward pass
(train_supervised.py:68-75; DataLoader(pin_memory=True) :37-47). `DevicePrefetcher` wraps any iterable of dict batches
(what `MultimodalCDDataset` + `DataLoader` yield: utils/datasets.py:111-181) and copies batch i+1 into a second set of
device buffers on a side stream while the kernels of batch i run, so the PCIe transfer (54 MB per 16 pairs of 6-band
256x256 patches) disappears behind the step. Tensors that must stay on the host (the `is_labeled` row mask of
train_semisupervised.py:80) and non-tensor entries pass through untouched.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence

import torch


class DevicePrefetcher:
    def __init__(self, batches: Iterable[dict], device: torch.device, keep_on_host: Sequence[str] = ("is_labeled",),
                 depth: int = 2):
        if device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches into CUDA memory; there is no CPU path")
        self.batches, self.device, self.keep, self.depth = batches, device, set(keep_on_host), max(2, depth)
        self.stream = torch.cuda.Stream(device=device)
        self._bufs: list[Optional[dict]] = [None] * self.depth
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]
        self._free: list[Optional[torch.cuda.Event]] = [None] * self.depth
        self.bytes_staged = 0

    def _stage(self, slot: int, batch: dict) -> None:
        bufs = self._bufs[slot] or {}
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])   # the consumer of this slot's previous batch has finished
            out = {}
            for k, v in batch.items():
                if not torch.is_tensor(v) or k in self.keep:
                    out[k] = v
                    continue
                b = bufs.get(k)
                if b is None or b.shape != v.shape or b.dtype != v.dtype:
                    b = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                b.copy_(v, non_blocking=True)
                if not v.is_cuda:
                    self.bytes_staged += v.numel() * v.element_size()
                out[k] = b
            self._ready[slot].record(self.stream)
        self._bufs[slot] = out

    def __iter__(self) -> Iterator[dict]:
        it = iter(self.batches)
        try:
            self._stage(0, next(it))
        except StopIteration:
            return
        slot = 0
        while True:
            nxt = next(it, None)
            if nxt is not None:
                self._stage((slot + 1) % self.depth, nxt)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ready[slot])
            yield self._bufs[slot]
            ev = torch.cuda.Event()
            ev.record(cur)                                  # everything queued on the batch so far
            self._free[slot] = ev
            if nxt is None:
                return
            slot = (slot + 1) % self.depth


class LossReader:
    """Device -> host read-back of the per-step loss without stalling the launch queue.

    The reference appends `loss.item()` every step (train_supervised.py:79), which drains the GPU before the next step
    can be queued. `push(loss)` instead copies the 0-d loss into a pinned host slot asynchronously and returns the
    value of the step pushed `lag` steps earlier (None until then), so the host keeps queueing work; `drain()` returns
    the values still in flight. Every step's loss is read exactly once."""

    def __init__(self, device: torch.device, lag: int = 1):
        self.lag = max(1, lag)
        self._slots = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(self.lag + 1)]
        self._events = [torch.cuda.Event() for _ in range(self.lag + 1)]
        self._pending: list[int] = []
        self._n = 0
        self.bytes_read = 0
        self.device = device

    def _take(self, slot: int) -> float:
        self._events[slot].synchronize()
        self.bytes_read += 4
        return float(self._slots[slot])

    def push(self, loss: torch.Tensor) -> Optional[float]:
        slot = self._n % (self.lag + 1)
        self._n += 1
        self._slots[slot].copy_(loss.detach().reshape(()), non_blocking=True)
        self._events[slot].record(torch.cuda.current_stream(self.device))
        self._pending.append(slot)
        if len(self._pending) > self.lag:
            return self._take(self._pending.pop(0))
        return None

    def drain(self) -> list:
        out = [self._take(s) for s in self._pending]
        self._pending = []
        return out
