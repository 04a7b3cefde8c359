"""Algorithm protocol and registry (reference: bagua/torch_api/algorithms/base.py:1-263).

An :class:`Algorithm` is a user-facing description; ``reify(process_group)`` turns it into an
:class:`AlgorithmImpl` whose eight override points decide which tensors are communicated, how they are bucketed, which
ops run on each bucket and which hooks fire."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import torch

from ...bucket import BaguaBucket

__all__ = ["Algorithm", "AlgorithmImpl", "GlobalAlgorithmRegistry"]


class Algorithm:
    """Base class of all algorithms."""

    def reify(self, process_group) -> "AlgorithmImpl":
        """Create the implementation instance working on ``process_group``."""
        raise NotImplementedError

    @classmethod
    def init(cls, name: str, **kwargs) -> "Algorithm":
        """Instantiate a registered algorithm by name, e.g. ``Algorithm.init("gradient_allreduce", hierarchical=True)``."""
        return GlobalAlgorithmRegistry.get(name)(**kwargs)


class AlgorithmImpl:
    """Base class of algorithm implementations; every method may be overridden."""

    def __init__(self, process_group):
        self.process_group = process_group

    def need_reset(self) -> bool:
        """``True`` → all ``init_*`` methods are called again before the next forward (multi-stage algorithms)."""
        return False

    def init_tensors(self, bagua_ddp) -> List[torch.Tensor]:
        """Register the tensors to communicate.  Default: every parameter's gradient, in *reverse* parameter order so
        bucket 0 holds what backward produces first (reference base.py:87-102)."""
        parameters = bagua_ddp.bagua_build_params()
        tensors = []
        for name, param in reversed(parameters):
            param = param.bagua_ensure_grad().ensure_bagua_tensor(
                name,
                bagua_ddp.bagua_module_name,
                getter_closure=lambda p: p.grad,
                setter_closure=lambda p, t: setattr(p, "grad", t),
            )
            tensors.append(param)
        self._communication_tensor_names = set(name for name, _ in parameters)
        assert len(self._communication_tensor_names) == len(tensors), "tensor names should be unique"
        return tensors

    def tensors_to_buckets(self, tensors: List[List[torch.Tensor]], do_flatten: bool) -> List[BaguaBucket]:
        """Turn the bucketing suggestion into buckets (default: follow it)."""
        return [BaguaBucket(b, flatten=do_flatten, name=str(i), group=self.process_group) for i, b in enumerate(tensors)]

    def init_forward_pre_hook(self, bagua_ddp) -> Callable:
        """Returns ``hook(input)`` run before every training forward."""

        def hook(input):
            pass

        return hook

    def init_backward_hook(self, bagua_ddp) -> Callable:
        """Returns ``hook(parameter_name, parameter)`` run when a parameter's gradient has been accumulated."""
        names = self._communication_tensor_names

        def hook(parameter_name, parameter):
            if parameter_name in names:
                # the gradient must still be the bucket view registered with the scheduler (zero_grad(set_to_none=True)
                # or an optimizer that replaces .grad would silently break the aliasing)
                if parameter._bagua_backend_tensor.data_ptr() != parameter.grad.data_ptr():
                    raise AssertionError("bagua backend tensor data_ptr should match parameter grad (the gradient must stay the bucket view: "
                                         "use zero_grad(set_to_none=False) and do not assign a new tensor to .grad)")
                bagua_ddp.mark_tensor_ready(parameter)

        return hook

    def init_post_backward_hook(self, bagua_ddp) -> Callable:
        """Returns ``hook()`` run once when the whole backward pass is done."""

        def hook():
            bagua_ddp.wait_pending_comm_ops()

        return hook

    def init_post_optimizer_step_hook(self, bagua_ddp) -> Callable:
        """Returns ``hook(optimizer)`` run after every ``optimizer.step()``."""

        def hook(optimizer: torch.optim.Optimizer):
            pass

        return hook

    def init_operations(self, bagua_ddp, bucket: BaguaBucket):
        """Register the communication ops of ``bucket``."""


class _AlgorithmRegistry(dict):
    def register(self, name: str, algorithm: Callable, description: Optional[str] = None):
        if not (name is None or isinstance(name, str)):
            raise TypeError(f"`name` must be a str, found {name}")
        if name in self:
            raise ValueError(f"'{name}' is already present in the registry.")
        data: Dict[str, Any] = {"algorithm": algorithm, "description": description or ""}
        self[name] = data

    def get(self, name: str) -> Callable:
        if name in self:
            return self[name]["algorithm"]
        available = ", ".join(sorted(self.keys())) or "none"
        raise KeyError(f"'{name}' not found in registry. Available names: {available}")

    def available_algorithms(self) -> List[str]:
        return list(self.keys())

    def __str__(self) -> str:
        return "Registered Algorithms: {}".format(", ".join(self.keys()))


GlobalAlgorithmRegistry = _AlgorithmRegistry()
