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


def _noop(*_args, **_kwargs):
    """Default hook body: nothing to do at this point of the step."""


class AlgorithmImpl:
    """Base class of algorithm implementations; every method may be overridden.  The defaults describe plain gradient communication:
    gradients are the communicated tensors, the bucketing suggestion is followed, a gradient is marked ready when autograd has
    accumulated it and the step waits for the scheduler after backward."""

    def __init__(self, process_group):
        self.process_group = process_group

    def need_reset(self) -> bool:
        """``True`` → all ``init_*`` methods are called again before the next forward (multi-stage algorithms)."""
        return False

    def init_tensors(self, bagua_ddp) -> List[torch.Tensor]:
        """Register the tensors to communicate. Default: every parameter's gradient, in *reverse* parameter order so
        bucket 0 holds what backward produces first (reference base.py:87-102)."""
        named = bagua_ddp.bagua_build_params()
        self._communication_tensor_names = {name for name, _ in named}
        assert len(self._communication_tensor_names) == len(named), "tensor names should be unique"

        def grad_of(p):
            return p.grad

        def set_grad(p, t):
            p.grad = t

        return [p.bagua_ensure_grad().ensure_bagua_tensor(name, bagua_ddp.bagua_module_name, getter_closure=grad_of, setter_closure=set_grad)
                for name, p in reversed(named)]

    def tensors_to_buckets(self, tensors: List[List[torch.Tensor]], do_flatten: bool) -> List[BaguaBucket]:
        """Turn the bucketing suggestion into buckets (default: follow it)."""
        return [BaguaBucket(b, flatten=do_flatten, name=str(i), group=self.process_group) for i, b in enumerate(tensors)]

    def init_forward_pre_hook(self, bagua_ddp) -> Callable:
        """Returns ``hook(input)`` run before every training forward."""
        return _noop

    def init_backward_hook(self, bagua_ddp) -> Callable:
        """Returns ``hook(parameter_name, parameter)`` run when a parameter's gradient has been accumulated."""
        communicated = self._communication_tensor_names

        def mark_gradient_ready(parameter_name, parameter):
            if parameter_name not in communicated:
                return
            if parameter._bagua_backend_tensor.data_ptr() != parameter.grad.data_ptr():
                raise AssertionError("bagua backend tensor data_ptr should match parameter grad (the gradient must stay the bucket view: "
                                     "use zero_grad(set_to_none=False) and do not assign a new tensor to .grad)")
            bagua_ddp.mark_tensor_ready(parameter)

        return mark_gradient_ready

    def init_post_backward_hook(self, bagua_ddp) -> Callable:
        """Returns ``hook()`` run once when the whole backward pass is done."""
        return bagua_ddp.wait_pending_comm_ops

    def init_post_optimizer_step_hook(self, bagua_ddp) -> Callable:
        """Returns ``hook(optimizer)`` run after every ``optimizer.step()``."""
        return _noop

    def init_operations(self, bagua_ddp, bucket: BaguaBucket):
        """Register the communication ops of ``bucket``."""


class _AlgorithmRegistry(dict):
    """``name → {"algorithm": factory, "description": text}`` (reference base.py:211-263: the same dict-shaped registry, so code
    that iterates ``GlobalAlgorithmRegistry`` keeps working)."""

    def register(self, name: str, algorithm: Callable, description: Optional[str] = None):
        if name is not None and not isinstance(name, str):
            raise TypeError(f"`name` must be a str, found {name}")
        if name in self:
            raise ValueError(f"'{name}' is already present in the registry.")
        entry: Dict[str, Any] = {"algorithm": algorithm, "description": description or ""}
        self[name] = entry

    def get(self, name: str) -> Callable:
        try:
            return self[name]["algorithm"]
        except KeyError:
            available = ", ".join(sorted(self.keys())) or "none"
            raise KeyError(f"'{name}' not found in registry. Available names: {available}") from None

    def available_algorithms(self) -> List[str]:
        return list(self)

    def __str__(self) -> str:
        return "Registered Algorithms: " + ", ".join(self)


GlobalAlgorithmRegistry = _AlgorithmRegistry()
