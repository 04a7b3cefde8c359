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

    def __init__(
        self,
        dataset: Dataset,
        complexity_fn: Callable[..., int],
        num_replicas: Optional[int] = None,
        rank: Optional[int] = None,
        shuffle: bool = True,
        seed: int = 0,
        drop_last: bool = False,
        random_level: float = 0,
    ) -> None:
        if num_replicas is None:
            if not dist.is_available() or not dist.is_initialized():
                raise RuntimeError("Requires distributed package to be available")
            num_replicas = dist.get_world_size()
        if rank is None:
            if not dist.is_available() or not dist.is_initialized():
                raise RuntimeError("Requires distributed package to be available")
            rank = dist.get_rank()
        if rank >= num_replicas or rank < 0:
            raise ValueError(f"Invalid rank {rank}, rank should be in the interval [0, {num_replicas - 1}]")
        if random_level < 0.0 or random_level > 1.0:
            raise ValueError(f"Invalid random level {random_level}, shoule be in the range [0.0, 1.0]")
        self.dataset = dataset
        self.num_replicas = num_replicas
        self.rank = rank
        self.epoch = 0
        self.drop_last = drop_last
        n = len(dataset)  # type: ignore[arg-type]
        if self.drop_last and n % num_replicas != 0:
            self.num_samples = math.ceil((n - num_replicas) / num_replicas)
        else:
            self.num_samples = math.ceil(n / num_replicas)
        self.total_size = self.num_samples * num_replicas
        self.shuffle = shuffle
        self.seed = seed
        self.complexities = np.asarray([complexity_fn(dataset[i]) for i in range(n)], dtype=np.int64)
        self.item_complexity_map = {i: int(c) for i, c in enumerate(self.complexities)}
        self._sorted_indices = np.argsort(self.complexities, kind="stable")
        span = int(self.complexities.max() - self.complexities.min()) if n else 0
        self.random_number = int(span * random_level + 1)

    def _chunks(self, ordered: np.ndarray) -> np.ndarray:
        """[num_chunks, num_replicas] index matrix, wrapping around the ordered list when it is too short."""
        num_chunks = max(1, self.num_samples)
        need = num_chunks * self.num_replicas
        reps = math.ceil(need / len(ordered))
        return np.tile(ordered, reps)[:need].reshape(num_chunks, self.num_replicas)

    def shuffle_chunks(self):
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            if self.random_number > 0:
                jitter = torch.randint(self.random_number, (len(self.complexities),), generator=g).numpy()
                ordered = np.argsort(self.complexities + jitter, kind="stable")
            else:
                ordered = self._sorted_indices
            index_chunks = self._chunks(ordered)
            chunk_indices = torch.randperm(len(index_chunks), generator=g).tolist()
        else:
            index_chunks = self._chunks(self._sorted_indices)
            chunk_indices = list(range(len(index_chunks)))
        if not self.drop_last:
            pad = self.num_samples - len(chunk_indices)
            if pad > 0:
                chunk_indices += (chunk_indices * math.ceil(pad / len(chunk_indices)))[:pad]
        else:
            chunk_indices = chunk_indices[: self.num_samples]
        assert len(chunk_indices) == self.num_samples
        return index_chunks.tolist(), chunk_indices

    def __iter__(self) -> Iterator:
        index_chunks, chunk_indices = self.shuffle_chunks()
        indices = [index_chunks[i][self.rank] for i in chunk_indices]
        assert len(indices) == self.num_samples
        return iter(indices)

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
        self.generate_batches()

    def generate_batches(self):
        index_chunks, chunk_indices = self.sampler.shuffle_chunks()
        batches = [self.batch_fn([index_chunks[i][r] for i in chunk_indices]) for r in range(self.num_replicas)]
        lens = [len(b) for b in batches]
        self.total_batch = min(lens) if self.drop_last else max(lens)
        self.padded_batches = [(b + b[: self.total_batch - len(b)])[: self.total_batch] for b in batches]

    def __iter__(self):
        return iter(self.padded_batches[self.rank])

    def __len__(self):
        return self.total_batch

    def set_epoch(self, epoch: int) -> None:
        self.sampler.set_epoch(epoch)
        self.generate_batches()
