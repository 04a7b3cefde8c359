"""Contrib utilities (reference: bagua/torch_api/contrib/__init__.py:1-7)."""
from .fuse.optimizer import fuse_optimizer, is_fused_optimizer  # noqa: F401
from .load_balancing_data_loader import LoadBalancingDistributedSampler, LoadBalancingDistributedBatchSampler  # noqa: F401
from .cache_loader import CacheLoader  # noqa: F401
from .cached_dataset import CachedDataset  # noqa: F401
from . import sync_batchnorm  # noqa: F401
