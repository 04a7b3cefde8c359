"""``DistributedDataParallel(module, ...)`` with the constructor of ``torch.nn.parallel.DistributedDataParallel`` plus
``optimizers=`` and ``algorithm=`` (reference: bagua/torch_api/data_parallel/distributed.py:93-360)."""
from __future__ import annotations

import warnings
from contextlib import contextmanager
from typing import List, Optional, Union

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as TorchDistributedDataParallel

from ... import communication as comm_mod
from ..algorithms.gradient_allreduce import GradientAllReduceAlgorithm
from ..bagua_distributed import BaguaDistributedDataParallel
from ..distributed import _name_counter

__all__ = ["DistributedDataParallel", "DistributedDataParallel_V1_9_0", "to_bagua_process_group"]


def to_bagua_process_group(process_group=None):
    """``None`` → default group; a torch ProcessGroup → its (cached) bagua wrapper; a BaguaProcessGroup → itself."""
    if process_group is None:
        return comm_mod._get_default_group()
    if isinstance(process_group, comm_mod.BaguaProcessGroup):
        return process_group
    if isinstance(process_group, dist.ProcessGroup):
        return comm_mod.from_torch_group(process_group)
    raise Exception(f"unexpect input {type(process_group)}")


class DistributedDataParallel_V1_9_0_Interface(torch.nn.Module):
    r"""The subset of PyTorch 1.9 DDP's interface that wrappers implement (reference distributed.py:19-60): ``forward``,
    ``no_sync``, ``scatter`` / ``to_kwargs`` / ``gather`` are inherited or overridden by the concrete class."""

    def no_sync(self):
        raise NotImplementedError

    def forward(self, *inputs, **kwargs):
        raise NotImplementedError


class DistributedDataParallel_V1_9_0(DistributedDataParallel_V1_9_0_Interface):
    r"""DDP-compatible module wrapper; the engine is ``self.inner``."""

    def __init__(
        self,
        module,
        device_ids=None,
        output_device=None,
        dim=0,
        broadcast_buffers=True,
        process_group=None,
        bucket_cap_mb=25,
        find_unused_parameters=False,
        check_reduction=False,
        gradient_as_bucket_view=True,
        optimizers: List[torch.optim.Optimizer] = [],
        algorithm=None,
    ) -> None:
        super().__init__()
        assert any(p.requires_grad for p in module.parameters()), (
            "DistributedDataParallel is not needed when a module doesn't have any parameter that requires a gradient."
        )
        if device_ids is not None and len(device_ids) > 1:
            raise ValueError("device_ids can only be None or contain a single element.")
        self.is_multi_device_module = len({p.device for p in module.parameters()}) > 1
        distinct = {p.device.type for p in module.parameters()}
        if len(distinct) != 1:
            raise ValueError(f"DistributedDataParallel's input module must be on the same type of devices, but input module parameters locate in {distinct}.")
        self.device_type = list(distinct)[0]
        self.static_graph = False
        self.dim = dim
        self.module = module
        self.device = next(module.parameters()).device
        assert broadcast_buffers is True, "Not yet supported"
        self.broadcast_buffers = broadcast_buffers
        self.find_unused_parameters = find_unused_parameters
        if not hasattr(module, "_bagua_module_name"):
            module._bagua_module_name = f"{self.__class__.__name__}_{next(_name_counter)}"
        self.inner = BaguaDistributedDataParallel(
            self.module,
            list(optimizers),
            algorithm if algorithm is not None else GradientAllReduceAlgorithm(),
            process_group=to_bagua_process_group(process_group),
            gradient_as_bucket_view=gradient_as_bucket_view,
            find_unused_parameters=find_unused_parameters,
            bagua_module_name=module.bagua_module_name,
        )

    @property
    def require_backward_grad_sync(self):
        """Gradient synchronisation switch, see :meth:`no_sync`."""
        return self.inner.require_backward_grad_sync

    @property
    def parameters_to_ignore(self):
        return self.inner.parameters_to_ignore

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    @contextmanager
    def no_sync(self):
        r"""Disable gradient synchronisation inside the context; gradients accumulate locally in the bucket views and are
        communicated by the first backward after leaving it."""
        old = self.require_backward_grad_sync
        self.inner.require_backward_grad_sync = False
        try:
            yield
        finally:
            self.inner.require_backward_grad_sync = old

    @property
    def bagua_algorithm(self):
        return self.inner.bagua_algorithm

    @property
    def bagua_optimizers(self):
        return self.inner.bagua_optimizers

    @property
    def bagua_buckets(self):
        return self.inner.bagua_buckets


def DistributedDataParallel(
    module: torch.nn.Module,
    device_ids: Optional[List[Union[int, torch.device]]] = None,
    output_device: Union[int, torch.device, None] = None,
    dim: int = 0,
    broadcast_buffers: bool = True,
    process_group=None,
    bucket_cap_mb: int = 25,
    find_unused_parameters: bool = False,
    check_reduction: bool = False,
    gradient_as_bucket_view: bool = True,
    optimizers: List[torch.optim.Optimizer] = [],
    algorithm=None,
):
    r"""PyTorch-DDP-compatible constructor plus ``optimizers`` and ``algorithm``.  Unsupported DDP arguments fall back to
    upstream ``torch.nn.parallel.DistributedDataParallel`` with a warning (reference distributed.py:319-346).

    Example::

        >>> bagua_b200.init_process_group()
        >>> net = bagua_b200.data_parallel.DistributedDataParallel(model, optimizers=[opt], algorithm=ByteGradAlgorithm())
    """
    supported = [device_ids is None, output_device is None, dim == 0, broadcast_buffers is True, check_reduction is False]
    if not all(supported):
        warnings.warn(
            "Some parameters passed into BaguaDistributedDataParallel have not been supported yet. "
            "Falling back to upstream PyTorch DistributedDataParallel."
        )
        return TorchDistributedDataParallel(
            module,
            device_ids=device_ids,
            output_device=output_device,
            dim=dim,
            broadcast_buffers=broadcast_buffers,
            process_group=process_group if isinstance(process_group, dist.ProcessGroup) or process_group is None else process_group.torch_group,
            bucket_cap_mb=bucket_cap_mb,
            find_unused_parameters=find_unused_parameters,
            gradient_as_bucket_view=gradient_as_bucket_view,
        )
    return DistributedDataParallel_V1_9_0(
        module,
        device_ids=device_ids,
        output_device=output_device,
        dim=dim,
        broadcast_buffers=broadcast_buffers,
        process_group=process_group,
        bucket_cap_mb=bucket_cap_mb,
        find_unused_parameters=find_unused_parameters,
        check_reduction=check_reduction,
        gradient_as_bucket_view=gradient_as_bucket_view,
        optimizers=optimizers,
        algorithm=algorithm,
    )
