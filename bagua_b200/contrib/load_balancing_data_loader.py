"""Samplers that balance per-sample cost across data-parallel replicas
(reference: bagua/torch_api/contrib/load_balancing_data_loader.py:1-324).

Samples are sorted by a user supplied complexity, cut into chunks of ``num_replicas`` neighbours, and replica ``r`` takes
element ``r`` of every chunk, so all replicas process similarly expensive samples in the same step."""
from __future__ import annotations

import math
from typing import Callable, Iterator, Optional

import numpy as np
import torch
import torch.distributed as dist
from torch.utils.data.dataset import Dataset
from torch.utils.data.sampler import Sampler

__all__ = ["LoadBalancingDistributedSampler", "LoadBalancingDistributedBatchSampler"]


class LoadBalancingDistributedSampler(Sampler):
    r"""
    Args:
        dataset: dataset to sample from.
        complexity_fn: ``sample -> int`` cost estimate of a sample.
        num_replicas / rank: default to the world size / rank of the current process group.
        shuffle: shuffle chunk order (and jitter complexities by ``random_level``) each epoch.
        seed: base seed, identical on all replicas.
        drop_last: drop the tail so every replica sees the same number of samples, instead of wrapping around.
        random_level: 0.0 = strict complexity order … 1.0 = complexities jittered over their whole range.
    """

    def __init__(self, dataset: Dataset, complexity_fn: Callable[..., int], num_replicas: Optional[int] = None, rank: Optional[int] = None,
                 shuffle: bool = True, seed: int = 0, drop_last: bool = False, random_level: float = 0) -> None:
        if num_replicas is None or rank is None:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("num_replicas / rank were not given and there is no initialised process group to take them from")
            num_replicas = dist.get_world_size() if num_replicas is None else num_replicas
            rank = dist.get_rank() if rank is None else rank
        if not 0 <= rank < num_replicas:
            raise ValueError(f"Invalid rank {rank}, rank should be in the interval [0, {num_replicas - 1}]")
        if not 0.0 <= random_level <= 1.0:
            raise ValueError(f"random_level must lie in [0.0, 1.0], got {random_level}")
        self.dataset, self.num_replicas, self.rank = dataset, num_replicas, rank
        self.shuffle, self.seed, self.drop_last, self.epoch = shuffle, seed, drop_last, 0
        n = len(dataset)  # type: ignore[arg-type]
        rows = (n - num_replicas) / num_replicas if (drop_last and n % num_replicas) else n / num_replicas
        self.num_samples = math.ceil(rows)                 # steps per epoch = rows of the epoch plan
        self.total_size = self.num_samples * num_replicas
        # complexities are evaluated once; an epoch only re-sorts (jittered) keys
        self.complexities = np.fromiter((complexity_fn(dataset[i]) for i in range(n)), dtype=np.int64, count=n)
        self.item_complexity_map = dict(enumerate(self.complexities.tolist()))
        self._by_cost = np.argsort(self.complexities, kind="stable")
        spread = int(np.ptp(self.complexities)) if n else 0
        self.random_number = int(spread * random_level + 1)    # exclusive upper bound of the per-epoch jitter

    # -- the epoch plan ----------------------------------------------------------------------------------------------------
    def _chunks(self, epoch: int):
        """``(grid, row_order)``: the cost-ordered indices laid out row by row (``num_replicas`` neighbours per row; a short list wraps
        around so that the last row is full) and the order in which the rows are visited in ``epoch``."""
        rows, width = max(1, self.num_samples), self.num_replicas
        order = self._by_cost
        row_order = np.arange(rows)
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + epoch)
            if self.random_number > 0:
                jitter = torch.randint(self.random_number, (len(self.complexities),), generator=g).numpy()
                order = np.argsort(self.complexities + jitter, kind="stable")
            row_order = torch.randperm(rows, generator=g).numpy()
        return np.resize(order, rows * width).reshape(rows, width), row_order

    def epoch_plan(self, epoch: Optional[int] = None) -> np.ndarray:
        """``[num_samples, num_replicas]`` matrix of dataset indices for ``epoch`` (default: the current one): row *t* is what the
        replicas load at step *t* — ``num_replicas`` neighbours in (jittered) cost order — and replica *r* reads column *r*.
        Deterministic in ``(seed, epoch)``, so every replica derives the same matrix without communicating."""
        grid, row_order = self._chunks(self.epoch if epoch is None else epoch)
        plan = grid[row_order]
        if len(plan) < self.num_samples:                    # only when the dataset is smaller than one row per step
            plan = np.resize(plan, (self.num_samples, self.num_replicas))
        return plan[: self.num_samples]

    def shuffle_chunks(self):
        """``(index_chunks, chunk_indices)`` in the reference's shape (load_balancing_data_loader.py:148-190): the chunks of
        ``num_replicas`` similar-cost samples and the order in which this epoch visits them."""
        grid, row_order = self._chunks(self.epoch)
        return grid.tolist(), row_order.tolist()

    def __iter__(self) -> Iterator:
        return iter(self.epoch_plan()[:, self.rank].tolist())

    def __len__(self) -> int:
        return self.num_samples

    def set_epoch(self, epoch: int) -> None:
        """Make the next iteration use the ordering of ``epoch`` (all replicas must pass the same value)."""
        self.epoch = epoch


class LoadBalancingDistributedBatchSampler(Sampler):
    r"""Variable-size mini-batches on top of :class:`LoadBalancingDistributedSampler`.

    ``batch_fn(indices: List[int]) -> List[List[int]]`` groups one replica's indices into batches; the number of batches
    is equalised across replicas (padded by repeating, or truncated when ``drop_last``)."""

    def __init__(self, sampler: LoadBalancingDistributedSampler, batch_fn, drop_last: bool = False) -> None:
        if not isinstance(sampler, LoadBalancingDistributedSampler):
            raise ValueError("sampler should be of LoadBalancingDistributedSampler type.")
        if sampler.drop_last:
            raise ValueError("drop_last of sampler should be False")
        self.sampler = sampler
        self.batch_fn = batch_fn
        self.drop_last = drop_last
        self.num_replicas = sampler.num_replicas
        self.rank = sampler.rank
        self._rebuild()

    def _rebuild(self):
        """Batch every replica's column of the epoch plan, then give all replicas the same number of batches: the shortest list's
        length when ``drop_last``, otherwise the longest (shorter lists repeat their first batches)."""
        plan = self.sampler.epoch_plan()
        per_replica = [self.batch_fn(plan[:, r].tolist()) for r in range(self.num_replicas)]
        counts = [len(b) for b in per_replica]
        self.total_batch = min(counts) if self.drop_last else max(counts)
        self.padded_batches = [(b + b[: self.total_batch - len(b)])[: self.total_batch] for b in per_replica]

    def generate_batches(self):
        """Rebuild the batches of the current epoch (name of the reference's method, load_balancing_data_loader.py:285-306)."""
        self._rebuild()

    def __iter__(self):
        return iter(self.padded_batches[self.rank])

    def __len__(self):
        return self.total_batch

    def set_epoch(self, epoch: int) -> None:
        self.sampler.set_epoch(epoch)
        self._rebuild()
