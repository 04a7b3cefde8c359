"""PyTorch-DDP-compatible front end (reference: bagua/torch_api/data_parallel/distributed.py:1-360)."""
from .distributed import DistributedDataParallel, DistributedDataParallel_V1_9_0, to_bagua_process_group  # noqa: F401
from ..bagua_distributed import BaguaDistributedDataParallel  # noqa: F401
from . import functional  # noqa: F401
