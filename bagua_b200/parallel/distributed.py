"""``torch.nn.Module.with_bagua`` (reference: bagua/torch_api/distributed.py:53-141)."""
from __future__ import annotations

import itertools
from typing import List

import torch

from .. import communication as comm_mod
from .bagua_distributed import BaguaDistributedDataParallel

__all__ = ["BaguaModule"]

_name_counter = itertools.count()


class BaguaModule:
    """Mixin whose methods are installed on :class:`torch.nn.Module`."""

    def with_bagua(self, optimizers: List[torch.optim.Optimizer], algorithm, process_group=None, do_flatten: bool = True):
        r"""Prepare the module for data-parallel training with ``algorithm``.

        Can be called again on the same module to switch algorithms (old hooks are removed).

        Args:
            optimizers: optimizers updating this module's parameters (their state is broadcast from rank 0 and their
                ``step`` is wrapped so algorithms can run a post-step hook).
            algorithm: a :class:`bagua_b200.algorithms.Algorithm`.
            process_group: :class:`BaguaProcessGroup` (default group when ``None``).
            do_flatten: fuse each bucket's tensors into one flat storage (symmetric memory on NVSwitch).
        """
        group = process_group if process_group is not None else comm_mod._get_default_group()
        if getattr(self, "_bagua_module_name", None) is None:
            self._bagua_module_name = f"{type(self).__name__}_{next(_name_counter)}"
        # the engine object owns hooks, buckets and the per-name native scheduler; calling with_bagua again replaces it
        self.bagua_ddp = BaguaDistributedDataParallel(self, optimizers=optimizers, algorithm=algorithm, process_group=group,
                                                      bagua_module_name=self._bagua_module_name, gradient_as_bucket_view=do_flatten)
        return self

    def _get_name(self) -> str:
        return self._bagua_module_name

    def _set_name(self, name: str) -> None:
        self._bagua_module_name = name

    bagua_module_name = property(_get_name, _set_name, doc="Unique name of the module inside this process (one native scheduler per name); settable.")
    bagua_algorithm = property(lambda self: self.bagua_ddp.bagua_algorithm, doc="The reified algorithm (an :class:`AlgorithmImpl`).")
    bagua_optimizers = property(lambda self: self.bagua_ddp.bagua_optimizers, doc="Optimizers registered with :meth:`with_bagua`.")
    bagua_buckets = property(lambda self: self.bagua_ddp.bagua_buckets, doc="The communication buckets of the current algorithm.")


def _install():
    for name in ("with_bagua", "bagua_module_name", "bagua_algorithm", "bagua_optimizers", "bagua_buckets"):
        setattr(torch.nn.Module, name, BaguaModule.__dict__[name])


_install()
