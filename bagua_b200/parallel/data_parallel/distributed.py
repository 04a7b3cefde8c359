"""``DistributedDataParallel(module, ...)`` with the constructor of ``torch.nn.parallel.DistributedDataParallel`` plus
``optimizers=`` and ``algorithm=`` (capability of the reference's bagua/torch_api/data_parallel/distributed.py:93-360).

The torch-DDP compatible options travel as one ``_DDPOptions`` record: it knows which combinations this engine implements
(``engine_supported``), which ones are handed to upstream DDP (``torch_kwargs``), and validates the module once.
"""
from __future__ import annotations

import warnings
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Callable, Any, List, Optional, Sequence

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


@dataclass
class _DDPOptions:
    """The keyword arguments of torch DDP (v1.9 signature) that the wrapper accepts."""

    device_ids: Optional[Sequence[Any]] = None
    output_device: Any = None
    dim: int = 0
    broadcast_buffers: bool = True
    process_group: Any = None
    bucket_cap_mb: int = 25
    find_unused_parameters: bool = False
    check_reduction: bool = False
    gradient_as_bucket_view: bool = True

    def engine_supported(self) -> bool:
        """One device per process, batch dimension 0, buffers synchronised: what the bagua engine implements."""
        return self.device_ids is None and self.output_device is None and self.dim == 0 and self.broadcast_buffers is True and not self.check_reduction

    def torch_kwargs(self) -> dict:
        pg = self.process_group
        if pg is not None and not isinstance(pg, dist.ProcessGroup):
            pg = pg.torch_group
        return dict(device_ids=self.device_ids, output_device=self.output_device, dim=self.dim, broadcast_buffers=self.broadcast_buffers, process_group=pg,
                    bucket_cap_mb=self.bucket_cap_mb, find_unused_parameters=self.find_unused_parameters, gradient_as_bucket_view=self.gradient_as_bucket_view)

    @staticmethod
    def check_module(module: torch.nn.Module, device_ids) -> str:
        params = list(module.parameters())
        if not any(p.requires_grad for p in params):
            raise AssertionError("DistributedDataParallel is not needed when a module doesn't have any parameter that requires a gradient.")
        if device_ids is not None and len(device_ids) > 1:
            raise ValueError("device_ids can only be None or contain a single element.")
        kinds = {p.device.type for p in params}
        if len(kinds) != 1:
            raise ValueError(f"DistributedDataParallel's input module must be on the same type of devices, but input module parameters locate in {kinds}.")
        return kinds.pop()


class DistributedDataParallel_V1_9_0_Interface(torch.nn.Module):
    r"""The method set of a PyTorch-1.9 ``DistributedDataParallel`` (reference data_parallel/distributed.py:24-60); everything is
    abstract here, the concrete class below fills in what the engine supports."""

    def no_sync(self):
        raise NotImplementedError

    def forward(self, *inputs, **kwargs):
        raise NotImplementedError

    def scatter(self, inputs, kwargs, device_ids):
        raise NotImplementedError

    def to_kwargs(self, inputs, kwargs, device_id):
        raise NotImplementedError

    def gather(self, outputs, output_device):
        raise NotImplementedError

    def train(self, mode=True):
        super().train(mode)
        return self

    def join(self, divide_by_initial_world_size=True, enable=True, throw_on_early_termination=False):
        raise NotImplementedError

    def register_comm_hook(self, state: object, hook: Callable):
        raise NotImplementedError

    def will_sync_module_buffers(self):
        raise NotImplementedError


class DistributedDataParallel_V1_9_0(DistributedDataParallel_V1_9_0_Interface):
    r"""DDP-compatible module wrapper; the engine is ``self.inner``."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0, broadcast_buffers=True, process_group=None, bucket_cap_mb=25,
                 find_unused_parameters=False, check_reduction=False, gradient_as_bucket_view=True, optimizers: Sequence[torch.optim.Optimizer] = (),
                 algorithm=None) -> None:
        super().__init__()
        self.device_type = _DDPOptions.check_module(module, device_ids)
        if broadcast_buffers is not True:
            raise AssertionError("broadcast_buffers=False is not supported by the bagua engine")
        params = list(module.parameters())
        self.is_multi_device_module = len({p.device for p in params}) > 1
        self.device = params[0].device
        self.module, self.dim, self.static_graph = module, dim, False
        self.broadcast_buffers, self.find_unused_parameters = broadcast_buffers, find_unused_parameters
        if not hasattr(module, "_bagua_module_name"):
            module._bagua_module_name = f"{type(self).__name__}_{next(_name_counter)}"
        self.inner = BaguaDistributedDataParallel(
            module, list(optimizers), algorithm if algorithm is not None else GradientAllReduceAlgorithm(),
            process_group=to_bagua_process_group(process_group), gradient_as_bucket_view=gradient_as_bucket_view,
            find_unused_parameters=find_unused_parameters, bagua_module_name=module.bagua_module_name, broadcast_buffers=broadcast_buffers)

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    def zero_grad(self, set_to_none: bool = False):
        """Gradients are views of the bucket arena and stay allocated (nn.Module's default would drop them)."""
        return super().zero_grad(set_to_none=False)

    def scatter(self, inputs, kwargs, device_ids):
        """Move positional and keyword inputs to the module's (single) device, as torch's DDP does before ``forward``."""
        from torch.nn.parallel.scatter_gather import scatter_kwargs

        return scatter_kwargs(inputs, kwargs, device_ids, dim=self.dim)

    def to_kwargs(self, inputs, kwargs, device_id):
        moved_inputs, moved_kwargs = self.scatter(inputs, kwargs, [device_id])
        return moved_inputs, moved_kwargs

    def gather(self, outputs, output_device):
        from torch.nn.parallel.scatter_gather import gather

        return gather(outputs, output_device, dim=self.dim)

    def join(self, divide_by_initial_world_size=True, enable=True, throw_on_early_termination=False):
        """Uneven inputs across ranks are not supported by the bucket scheduler (every rank must mark every bucket every step) — the
        reference leaves this method abstract as well (data_parallel/distributed.py:47-54)."""
        raise NotImplementedError("join() (uneven inputs) is not supported; pad or drop the last incomplete batch instead")

    def register_comm_hook(self, state: object, hook: Callable):
        """Gradient communication is defined by the ``algorithm`` (``AlgorithmImpl.init_operations`` / ``append_python_op``), not by a
        torch comm hook — abstract in the reference too (:56-57)."""
        raise NotImplementedError("use a bagua Algorithm (e.g. a custom AlgorithmImpl with bucket.append_python_op) instead of a comm hook")

    def will_sync_module_buffers(self) -> bool:
        """Whether the next forward broadcasts the module's buffers from rank 0 (the engine does so whenever buffers exist)."""
        return self.broadcast_buffers and any(True for _ in self.module.buffers())

    @contextmanager
    def no_sync(self):
        r"""Disable gradient synchronisation inside the context; gradients accumulate locally in the bucket views and are
        communicated by the first backward after leaving it."""
        previous = self.inner.require_backward_grad_sync
        self.inner.require_backward_grad_sync = False
        try:
            yield
        finally:
            self.inner.require_backward_grad_sync = previous

    # engine state surfaced on the wrapper
    require_backward_grad_sync = property(lambda self: self.inner.require_backward_grad_sync, doc="Gradient synchronisation switch, see :meth:`no_sync`.")
    parameters_to_ignore = property(lambda self: self.inner.parameters_to_ignore)
    bagua_algorithm = property(lambda self: self.inner.bagua_algorithm)
    bagua_optimizers = property(lambda self: self.inner.bagua_optimizers)
    bagua_buckets = property(lambda self: self.inner.bagua_buckets)


def DistributedDataParallel(module: torch.nn.Module, device_ids=None, output_device=None, dim: int = 0, broadcast_buffers: bool = True, process_group=None,
                            bucket_cap_mb: int = 25, find_unused_parameters: bool = False, check_reduction: bool = False,
                            gradient_as_bucket_view: bool = True, optimizers: List[torch.optim.Optimizer] = [], algorithm=None):
    r"""PyTorch-DDP-compatible constructor plus ``optimizers`` and ``algorithm``.  Option combinations the engine does not
    implement fall back to upstream ``torch.nn.parallel.DistributedDataParallel`` with a warning.

    Example::

        >>> bagua_b200.init_process_group()
        >>> net = bagua_b200.data_parallel.DistributedDataParallel(model, optimizers=[opt], algorithm=ByteGradAlgorithm())
    """
    opts = _DDPOptions(device_ids, output_device, dim, broadcast_buffers, process_group, bucket_cap_mb, find_unused_parameters, check_reduction,
                       gradient_as_bucket_view)
    if not opts.engine_supported():
        warnings.warn("Some parameters passed into BaguaDistributedDataParallel have not been supported yet. "
                      "Falling back to upstream PyTorch DistributedDataParallel.")
        return TorchDistributedDataParallel(module, **opts.torch_kwargs())
    return DistributedDataParallel_V1_9_0(module, process_group=process_group, bucket_cap_mb=bucket_cap_mb, find_unused_parameters=find_unused_parameters,
                                          gradient_as_bucket_view=gradient_as_bucket_view, optimizers=optimizers, algorithm=algorithm)
