"""Host→device input pipeline helpers."""
from __future__ import annotations

from typing import Callable, Iterable, Iterator, Optional

import torch

__all__ = ["DevicePrefetcher", "LossReader"]


class DevicePrefetcher:
    """Iterate device batches while the *next* batch is already being copied from pinned host memory on a side stream.

    ``transform(device_batch)`` (e.g. cast to bf16 / channels_last) also runs on the side stream, so the compute stream
    only ever waits on an event."""

    def __init__(self, loader: Iterable, device: torch.device, transform: Optional[Callable] = None):
        self.loader = loader
        self.device = device
        self.transform = transform
        self.on_cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.on_cuda else None
        self._events = [torch.cuda.Event() for _ in range(4)] if self.on_cuda else []   # recycled: at most two batches are in flight
        self._n = 0

    def _stage(self, batch):
        with torch.cuda.stream(self.stream):
            moved = tuple(t.to(self.device, non_blocking=True) for t in batch)
            if self.transform is not None:
                moved = self.transform(*moved)
            ev = self._events[self._n % len(self._events)]
            self._n += 1
            ev.record(self.stream)
        return moved, ev

    def __iter__(self) -> Iterator:
        it = iter(self.loader)
        if not self.on_cuda:  # host-only run: nothing to overlap, same interface
            for batch in it:
                moved = tuple(t.to(self.device) for t in batch)
                yield self.transform(*moved) if self.transform is not None else moved
            return
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = self._stage(next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream().wait_event(ev)
            for t in cur:
                if isinstance(t, torch.Tensor):
                    t.record_stream(torch.cuda.current_stream())
            yield cur


class LossReader:
    """Device→host read of every step's loss without draining the launch queue: each step copies its loss into a pinned
    slot asynchronously; ``push`` returns the value of the step ``lag`` steps back (already on the host, or nearly so),
    ``flush`` reads whatever is still outstanding and returns the last value.  With ``lag=1`` the host may run at most one
    step ahead of the GPU — any hiccup on the launching thread then idles the device; ``lag=2`` keeps two steps queued."""

    def __init__(self, device: torch.device, slots: int = 8, lag: int = 2):
        assert 1 <= lag < slots
        self.on_cuda = torch.device(device).type == "cuda"
        self.buf = torch.zeros(slots, dtype=torch.float32)
        if self.on_cuda:
            self.buf = self.buf.pin_memory()
        self.events = [torch.cuda.Event() if self.on_cuda else None for _ in range(slots)]
        self.n = 0          # losses pushed
        self.read = 0       # losses already read on the host
        self.slots = slots
        self.lag = lag
        self.last: Optional[float] = None

    def _read_until(self, count: int):
        while self.read < count:
            j = self.read % self.slots
            if self.events[j] is not None:
                self.events[j].synchronize()
            self.last = float(self.buf[j])
            self.read += 1

    def push(self, loss: torch.Tensor) -> Optional[float]:
        i = self.n % self.slots
        self.buf[i : i + 1].copy_(loss.detach().float().reshape(1), non_blocking=True)
        if self.events[i] is not None:
            self.events[i].record()
        self.n += 1
        self._read_until(self.n - self.lag)
        return self.last

    def flush(self) -> Optional[float]:
        self._read_until(self.n)
        return self.last
