"""``torch.Tensor`` extension methods used by the engine (reference: bagua/torch_api/tensor.py:1-269).

A *bagua tensor* is a torch tensor registered for communication under a unique name.  What is communicated is its
*effective* tensor, selected by an optional getter closure (e.g. ``param.grad`` for gradient algorithms, the
parameter itself for weight-averaging ones, ``exp_avg`` for QAdam).  The native record (``_C.Tensor``) caches the
effective tensor's device pointer, so the comm worker never calls back into Python to look at it.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .core import dtype_code, native

__all__ = ["BaguaTensor", "dense_strides"]


def dense_strides(t: torch.Tensor):
    """Strides to use when ``t`` is re-homed into a flat buffer: its own strides when it is a dense permutation
    (contiguous, channels_last, ...) so the memory order — and the autograd gradient-layout contract — is preserved;
    plain contiguous strides otherwise."""
    if t.numel() == 0 or t.is_contiguous():
        return t.stride()
    order = sorted(range(t.dim()), key=lambda d: (t.stride(d), t.size(d)))
    expect = 1
    for d in order:
        if t.size(d) == 1:
            continue
        if t.stride(d) != expect:
            return torch.empty(t.shape).stride()
        expect *= t.size(d)
    return t.stride()


def _device_id(t: torch.Tensor) -> int:
    return t.device.index if t.device.type == "cuda" else -1


class BaguaTensor:
    """Mixin holding the patch methods; they are installed on ``torch.Tensor`` at import time."""

    def _bagua_sanity_check(self):
        eff = self.bagua_getter_closure()
        bt = self._bagua_backend_tensor
        assert bt.data_ptr() == eff.data_ptr(), "bagua backend tensor data_ptr should match the effective tensor"
        assert bt.num_elements() == eff.numel()

    def is_bagua_tensor(self) -> bool:
        return hasattr(self, "_bagua_backend_tensor")

    def ensure_bagua_tensor(
        self,
        name: Optional[str] = None,
        module_name: Optional[str] = None,
        getter_closure: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
        setter_closure: Optional[Callable[[torch.Tensor, torch.Tensor], None]] = None,
    ):
        """Convert this tensor in place to a bagua tensor (idempotent when name/module match)."""
        if self.is_bagua_tensor():
            if name is not None:
                assert self.bagua_tensor_name == name, "assigning a different name to an existing bagua tensor is forbidden"
            if module_name is not None:
                self.bagua_module_name = module_name
            # re-registration REPLACES the closures, ``None`` meaning "the tensor itself" exactly as on first registration
            # (reference tensor.py:60-88): a module switched from a gradient algorithm (getter = p.grad) to a weight algorithm
            # (decentralized, async averaging: no closures) must communicate its weights from then on, not stale gradients
            self._bagua_getter_closure = getter_closure
            self._bagua_setter_closure = setter_closure
            self._bagua_refresh_backend_tensor()
            return self
        self.bagua_tensor_name = name if name is not None else ""
        self.bagua_module_name = module_name
        self._bagua_getter_closure = getter_closure
        self._bagua_setter_closure = setter_closure
        self._bagua_ready_event = torch.cuda.Event() if self.device.type == "cuda" else None
        self._bagua_bucket = None
        self._bagua_backend = None
        eff = self.bagua_getter_closure()
        self._bagua_backend_tensor = native().Tensor(self.bagua_tensor_name, eff.data_ptr(), eff.numel(), dtype_code(eff.dtype), _device_id(eff))
        return self

    def to_bagua_tensor(self, name=None, module_name=None, getter_closure=None, setter_closure=None):
        """Like :meth:`ensure_bagua_tensor` but returns a new view sharing storage with ``self``."""
        new = self.view(self.dtype)
        return new.ensure_bagua_tensor(name, module_name, getter_closure, setter_closure)

    def _bagua_refresh_backend_tensor(self):
        eff = self.bagua_getter_closure()
        bt = self._bagua_backend_tensor
        if bt.num_elements() != eff.numel() or bt.dtype() != dtype_code(eff.dtype) or bt.device_id() != _device_id(eff):
            self._bagua_backend_tensor = native().Tensor(self.bagua_tensor_name, eff.data_ptr(), eff.numel(), dtype_code(eff.dtype), _device_id(eff))
        elif bt.data_ptr() != eff.data_ptr():
            bt.reset_ptr(eff.data_ptr())

    def bagua_getter_closure(self) -> torch.Tensor:
        """The effective tensor (what is actually communicated)."""
        g = getattr(self, "_bagua_getter_closure", None)
        return g(self) if g is not None else self

    def bagua_setter_closure(self, tensor: torch.Tensor):
        """Replace the effective tensor."""
        s = getattr(self, "_bagua_setter_closure", None)
        assert s is not None, "this bagua tensor has no setter closure"
        s(self, tensor)
        self._bagua_refresh_backend_tensor()

    def bagua_backend_tensor(self):
        """The native record registered with the scheduler."""
        return self._bagua_backend_tensor

    def bagua_ensure_grad(self) -> torch.Tensor:
        """Make sure ``.grad`` exists (zeros) so it can be bucketed (reference tensor.py:197-212)."""
        if self.grad is None:
            with torch.no_grad():
                self.grad = torch.zeros_like(self.data)
        return self

    def bagua_mark_communication_ready(self):
        """Tell the scheduler this tensor can be communicated once the work queued so far on the current stream is
        done (reference tensor.py:214-225)."""
        be = self._bagua_backend
        assert be is not None, "tensor is not registered with a backend (call with_bagua first)"
        if self.device.type == "cuda":
            torch.cuda.current_stream().record_event(self._bagua_ready_event)
            be.mark_communication_ready(self._bagua_backend_tensor, self._bagua_ready_event.cuda_event)
        else:
            be.mark_communication_ready(self._bagua_backend_tensor, 0)

    def bagua_mark_communication_ready_without_synchronization(self):
        """Mark ready without ordering against the current stream (reference tensor.py:227-237)."""
        be = self._bagua_backend
        assert be is not None, "tensor is not registered with a backend (call with_bagua first)"
        be.mark_communication_ready(self._bagua_backend_tensor, 0)

    def bagua_set_storage(self, storage, storage_offset: int = 0):
        """Re-point the effective tensor at ``storage[storage_offset:]`` keeping shape (reference tensor.py:239-263)."""
        eff = self.bagua_getter_closure()
        strides = dense_strides(eff)
        with torch.no_grad():
            if getattr(self, "_bagua_setter_closure", None) is not None:
                new = torch.empty(0, dtype=eff.dtype, device=eff.device).set_(storage, storage_offset, eff.shape, strides)
                self.bagua_setter_closure(new)
            else:
                eff.set_(storage, storage_offset, eff.shape, strides)
                self._bagua_refresh_backend_tensor()


def _install():
    for name, attr in vars(BaguaTensor).items():
        if name.startswith("__"):
            continue
        if callable(attr):
            setattr(torch.Tensor, name, attr)


_install()
