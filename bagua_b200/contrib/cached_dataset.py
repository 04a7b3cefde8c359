"""Dataset wrapper caching samples in a key-value store (reference: bagua/torch_api/contrib/cached_dataset.py:1-62)."""
from __future__ import annotations

from torch.utils.data.dataset import Dataset

from .cache_loader import CacheLoader

__all__ = ["CachedDataset"]


class CachedDataset(Dataset):
    r"""Wrap ``dataset`` so that ``dataset[i]`` is computed once and served from the cache afterwards — useful when
    loading/pre-processing a sample is expensive and the dataset is small enough to keep.

    Args:
        dataset: the dataset to wrap (samples must be deterministic and picklable).
        backend / dataset_name / writer_buffer_size / kwargs: see :class:`CacheLoader`.
    """

    def __init__(self, dataset: Dataset, backend: str = "redis", dataset_name: str = "", writer_buffer_size: int = 20, **kwargs):
        self.dataset = dataset
        #: the :class:`CacheLoader` behind this dataset (its ``hits`` / ``misses`` / ``store_errors`` counters tell whether the cache works)
        self.cache_loader = CacheLoader(backend, dataset_name, writer_buffer_size, **kwargs)

    def __len__(self) -> int:
        return len(self.dataset)

    def __getitem__(self, index):
        # the sample's position is its cache key; a miss falls through to the wrapped dataset and queues the result for writing
        return self.cache_loader.get(index, self.dataset.__getitem__)

    def cache_stats(self) -> dict:
        """``{"hits", "misses", "store_errors"}`` of the underlying loader since construction."""
        c = self.cache_loader
        return {"hits": c.hits, "misses": c.misses, "store_errors": c.store_errors}
