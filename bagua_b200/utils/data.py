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
        self.stream = torch.cuda.Stream(device=device)

    def _stage(self, batch):
        with torch.cuda.stream(self.stream):
            moved = tuple(t.to(self.device, non_blocking=True) for t in batch)
            if self.transform is not None:
                moved = self.transform(*moved)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return moved, ev

    def __iter__(self) -> Iterator:
        it = iter(self.loader)
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
    """Device→host read of every step's loss without stalling the launch queue: each step copies its loss into a pinned
    slot asynchronously; ``push`` returns the value of the *previous* step (already on the host), ``flush`` the last one."""

    def __init__(self, device: torch.device, slots: int = 4):
        self.buf = torch.zeros(slots, dtype=torch.float32).pin_memory()
        self.events = [torch.cuda.Event() for _ in range(slots)]
        self.n = 0
        self.slots = slots

    def push(self, loss: torch.Tensor) -> Optional[float]:
        i = self.n % self.slots
        self.buf[i : i + 1].copy_(loss.detach().float().reshape(1), non_blocking=True)
        self.events[i].record()
        self.n += 1
        if self.n < 2:
            return None
        j = (self.n - 2) % self.slots
        self.events[j].synchronize()
        return float(self.buf[j])

    def flush(self) -> Optional[float]:
        if self.n == 0:
            return None
        j = (self.n - 1) % self.slots
        self.events[j].synchronize()
        return float(self.buf[j])
