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
        self.cache_loader = CacheLoader(backend, dataset_name, writer_buffer_size, **kwargs)

    def __getitem__(self, item):
        return self.cache_loader.get(item, lambda x: self.dataset[x])

    def __len__(self):
        return len(self.dataset)
