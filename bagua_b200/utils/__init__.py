"""Small host-side helpers (reference: bagua/torch_api/utils.py:1-244)."""
from __future__ import annotations

import logging
import math
import time
from collections import OrderedDict
from typing import List, Optional

import torch

from ..core import to_bagua_datatype  # noqa: F401  (re-export, reference utils.py:81-92)

__all__ = [
    "flatten",
    "unflatten",
    "check_contiguous",
    "get_flattened_tensor",
    "to_bagua_datatype",
    "StatisticalAverage",
    "average_by_removing_extreme_values",
    "align_size",
    "GraphedTrainStep",
    "export_chrome_trace",
]

LOGGER = logging.getLogger(__name__)   # reference utils.py:10


def flatten(tensors: List[torch.Tensor]) -> torch.Tensor:
    """Concatenate tensors into one 1-D tensor."""
    if len(tensors) == 1:
        return tensors[0].contiguous().view(-1)
    return torch.cat([t.contiguous().view(-1) for t in tensors], dim=0)


def unflatten(flat: torch.Tensor, tensors: List[torch.Tensor]) -> List[torch.Tensor]:
    """Views of ``flat`` with the shapes of ``tensors``."""
    outs, off = [], 0
    for t in tensors:
        n = t.numel()
        outs.append(flat.narrow(0, off, n).view_as(t))
        off += n
    return outs


def check_contiguous(tensors: List[torch.Tensor]) -> bool:
    """True when the tensors sit back to back in one storage, in order (reference utils.py:51-57)."""
    data_ptr = None
    for t in tensors:
        if data_ptr is not None and t.data_ptr() != data_ptr:
            return False
        data_ptr = t.data_ptr() + t.numel() * t.element_size()
    return True


def get_flattened_tensor(tensors: List[torch.Tensor]) -> Optional[torch.Tensor]:
    """A fresh flat tensor holding copies of ``tensors`` (reference utils.py:60-78)."""
    if len(tensors) == 0:
        return None
    total = sum(t.numel() for t in tensors)
    flat = torch.zeros(total, dtype=tensors[0].dtype, device=tensors[0].device)
    off = 0
    with torch.no_grad():
        for t in tensors:
            flat[off : off + t.numel()].copy_(t.reshape(-1))
            off += t.numel()
    return flat


def align_size(size: int, align: int) -> int:
    """``size`` rounded up to a multiple of ``align``."""
    return int(math.ceil(size / align)) * align


def average_by_removing_extreme_values(raw_score_list):
    """Mean of the values within one standard deviation of the mean (reference utils.py:95-124)."""
    import numpy as np

    arr = np.array(raw_score_list, dtype=float)
    if arr.size == 0:
        return float("nan"), float("nan"), []
    mean, std = arr.mean(), arr.std()
    kept = [x for x in arr if mean - std <= x <= mean + std] or list(arr)
    return float(np.mean(kept)), float(np.std(kept)), kept


class StatisticalAverage:
    """Running mean over look-back windows of 1, 2, 4, ... seconds (reference utils.py:127-244).

    ``records[k]`` is the mean of the tracked quantity over the most recent ``2**k`` seconds as of
    ``last_update_time``; ``record_tail = (extra_seconds, mean)`` is the mean over the *whole* recorded span
    ``2**(len(records)-1) + extra_seconds``.  Look-backs between the kept windows are linearly interpolated.
    """

    def __init__(self, last_update_time: Optional[float] = None, records: Optional[List[float]] = None, record_tail=(0.0, 0.0)):
        self.last_update_time: float = time.time() if last_update_time is None else last_update_time
        self.records: List[float] = list(records) if records is not None else []
        self.record_tail = (float(record_tail[0]), float(record_tail[1]))

    def record_seconds(self) -> float:
        return 2.0 ** (len(self.records) - 1) if self.records else 0.0

    def total_recording_time(self) -> float:
        return self.record_seconds() + self.record_tail[0]

    def _knots(self):
        """(window_seconds, mean) pairs, increasing in window length."""
        pts = [(2.0 ** k, v) for k, v in enumerate(self.records)]
        if self.record_tail[0] > 0 or not pts:
            pts.append((self.total_recording_time(), self.record_tail[1]))
        return pts

    def get_records_mean(self, last_n_seconds: float) -> float:
        """Mean over the ``last_n_seconds`` before ``last_update_time``."""
        if last_n_seconds <= 0.0:
            return 0.0
        pts = self._knots()
        if last_n_seconds <= pts[0][0]:
            return pts[0][1]
        for (x0, y0), (x1, y1) in zip(pts, pts[1:]):
            if last_n_seconds <= x1:
                return y0 + (y1 - y0) * (last_n_seconds - x0) / (x1 - x0) if x1 > x0 else y1
        return pts[-1][1]

    def record(self, val: float):
        """``val`` was the value of the tracked quantity since the previous call."""
        now = time.time()
        dt = max(now - self.last_update_time, 0.0)
        old_total = self.total_recording_time()
        new_total = dt + old_total
        new_records: List[float] = []
        k = 0
        while 2.0 ** k <= new_total and k < 64:
            window = 2.0 ** k
            if window <= dt:
                new_records.append(val)
            else:
                share = dt / window
                new_records.append(val * share + self.get_records_mean(window - dt) * (1.0 - share))
            k += 1
        covered = 2.0 ** (len(new_records) - 1) if new_records else 0.0
        if new_total > covered:
            share = dt / new_total if new_total > 0 else 1.0
            whole = val * share + self.get_records_mean(old_total) * (1.0 - share)
            new_tail = (new_total - covered, whole)
        else:
            new_tail = (0.0, 0.0)
        self.last_update_time = now
        self.records = new_records
        self.record_tail = new_tail

    def get(self, last_n_seconds: float) -> float:
        """Mean over the ``last_n_seconds`` before *now* (time since the last record counts as unknown → newest value)."""
        idle = time.time() - self.last_update_time
        if last_n_seconds <= idle:
            return self.records[0] if self.records else self.record_tail[1]
        return self.get_records_mean(last_n_seconds - idle)

    def __str__(self) -> str:
        return str({"last_update_time": self.last_update_time, "records": self.records, "record_tail": self.record_tail})


def apply_flattened_call(bucket: List[torch.Tensor], call, extra_args=None):
    """Run ``call`` once on the concatenation of ``bucket`` and scatter the result back (``dist.all_reduce`` results are
    averaged) — reference utils.py:16-28."""
    import torch.distributed as dist

    coalesced = flatten(bucket)
    call(coalesced, *(extra_args or ()))
    if call is dist.all_reduce:
        coalesced /= dist.get_world_size()
    for buf, synced in zip(bucket, unflatten(coalesced, bucket)):
        buf.copy_(synced)


def apply_flattened_call_all(tensors: List[torch.Tensor], call):
    """:func:`apply_flattened_call` per tensor type (dtype + device) — reference utils.py:41-48."""
    groups = OrderedDict()
    for t in tensors:
        groups.setdefault(t.type(), []).append(t)
    for group in groups.values():
        apply_flattened_call(group, call)


def __getattr__(name):  # lazy: utils.graph is only needed by scripts that capture CUDA graphs
    if name == "GraphedTrainStep":
        from .graph import GraphedTrainStep

        return GraphedTrainStep
    if name == "export_chrome_trace":
        from .trace import export_chrome_trace

        return export_chrome_trace
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")

