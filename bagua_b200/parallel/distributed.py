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
        if process_group is None:
            process_group = comm_mod._get_default_group()
        if not hasattr(self, "_bagua_module_name"):
            self._bagua_module_name = f"{self.__class__.__name__}_{next(_name_counter)}"
        self.bagua_ddp = BaguaDistributedDataParallel(
            self,
            optimizers=optimizers,
            algorithm=algorithm,
            process_group=process_group,
            bagua_module_name=self.bagua_module_name,
            gradient_as_bucket_view=do_flatten,
        )
        return self

    @property
    def bagua_module_name(self):
        """Unique name of the module inside this process (one native scheduler per name)."""
        return self._bagua_module_name

    @bagua_module_name.setter
    def bagua_module_name(self, name: str):
        self._bagua_module_name = name

    @property
    def bagua_algorithm(self):
        """The reified algorithm (an :class:`AlgorithmImpl`)."""
        return self.bagua_ddp.bagua_algorithm

    @property
    def bagua_optimizers(self):
        return self.bagua_ddp.bagua_optimizers

    @property
    def bagua_buckets(self):
        return self.bagua_ddp.bagua_buckets


def _install():
    for name in ("with_bagua", "bagua_module_name", "bagua_algorithm", "bagua_optimizers", "bagua_buckets"):
        setattr(torch.nn.Module, name, BaguaModule.__dict__[name])


_install()
