"""PyTorch-Lightning strategy backed by this engine.

The reference is integrated into Lightning through an external ``BaguaStrategy`` (exercised by the reference's
tests/pytorch_lightning/test_bagua_strategy.py:30-107: ``Trainer(strategy=BaguaStrategy(algorithm="gradient_allreduce"))``).
Lightning ≥ 2.0 dropped that class, so an equivalent lives here.  It is built lazily: importing this module never requires
Lightning, :func:`make_bagua_strategy` / :class:`BaguaStrategy` raise a clear error when it is absent.

    >>> from bagua_b200.contrib.lightning import BaguaStrategy
    >>> trainer = Trainer(strategy=BaguaStrategy(algorithm="bytegrad"), accelerator="gpu", devices=8)
"""
from __future__ import annotations

import logging
from typing import Any, Dict, Optional, Union

import torch

logger = logging.getLogger(__name__)

__all__ = ["BaguaStrategy", "make_bagua_strategy", "lightning_available"]


def _lightning():
    try:
        import lightning.pytorch as pl  # Lightning >= 2.0
        from lightning.pytorch.strategies import DDPStrategy

        return pl, DDPStrategy
    except ImportError:
        try:
            import pytorch_lightning as pl
            from pytorch_lightning.strategies import DDPStrategy

            return pl, DDPStrategy
        except ImportError:
            return None, None


def lightning_available() -> bool:
    return _lightning()[0] is not None


class _UnwrapForward(torch.nn.Module):
    """Routes ``forward`` to the LightningModule's ``training_step`` / ``validation_step`` … the way Lightning's own DDP
    wrapper does, so the engine's forward-pre hook and autograd hooks see every step."""

    def __init__(self, lightning_module):
        super().__init__()
        self.module = lightning_module

    def forward(self, *args, **kwargs):
        trainer = getattr(self.module, "_trainer", None)
        if trainer is not None:
            if trainer.training:
                return self.module.training_step(*args, **kwargs)
            if getattr(trainer, "testing", False):
                return self.module.test_step(*args, **kwargs)
            if getattr(trainer, "sanity_checking", False) or getattr(trainer, "validating", False):
                return self.module.validation_step(*args, **kwargs)
            if getattr(trainer, "predicting", False):
                return self.module.predict_step(*args, **kwargs)
        return self.module(*args, **kwargs)


def _build_algorithm(algorithm: Union[str, Any], optimizers, kwargs: Dict[str, Any]):
    from ..parallel.algorithms import Algorithm
    from ..parallel.algorithms.q_adam import QAdamOptimizer

    if not isinstance(algorithm, str):
        return algorithm
    if algorithm == "qadam":
        if not optimizers or not isinstance(optimizers[0], QAdamOptimizer):
            raise ValueError("algorithm='qadam' needs configure_optimizers() to return a bagua QAdamOptimizer")
        kwargs = dict(kwargs, q_adam_optimizer=optimizers[0])
    return Algorithm.init(algorithm, **kwargs)


def make_bagua_strategy():
    """Returns the ``BaguaStrategy`` class (a ``DDPStrategy`` subclass of the installed Lightning)."""
    pl, DDPStrategy = _lightning()
    if pl is None:
        raise ImportError("BaguaStrategy needs `lightning` (>= 2.0) or `pytorch_lightning`; neither is installed")

    class _BaguaStrategy(DDPStrategy):
        strategy_name = "bagua"

        def __init__(self, algorithm: Union[str, Any] = "gradient_allreduce", flatten: bool = True, accelerator=None, parallel_devices=None,
                     cluster_environment=None, checkpoint_io=None, precision_plugin=None, **bagua_kwargs):
            super().__init__(accelerator=accelerator, parallel_devices=parallel_devices, cluster_environment=cluster_environment,
                             checkpoint_io=checkpoint_io, precision_plugin=precision_plugin)
            self._bagua_algorithm = algorithm
            self._bagua_flatten = flatten
            self._bagua_kwargs = bagua_kwargs

        # -- process group -----------------------------------------------------------------------------------------
        def setup_distributed(self):
            import os

            import bagua_b200 as bagua

            env = self.cluster_environment
            os.environ.setdefault("MASTER_ADDR", str(env.main_address))
            os.environ.setdefault("MASTER_PORT", str(env.main_port))
            os.environ["RANK"], os.environ["WORLD_SIZE"] = str(env.global_rank()), str(env.world_size())
            os.environ["LOCAL_RANK"], os.environ["NODE_RANK"] = str(env.local_rank()), str(env.node_rank())
            os.environ.setdefault("LOCAL_WORLD_SIZE", str(self.num_processes))
            if self.root_device.type == "cuda":
                torch.cuda.set_device(self.root_device)
            if not bagua.is_initialized():
                bagua.init_process_group()

        # -- model wrapping: optimizers first (the engine hooks optimizer.step), then with_bagua ----------------------
        def setup(self, trainer):
            assert self.accelerator is not None
            self.accelerator.setup(trainer)
            self.model_to_device()
            fitting = str(getattr(trainer.state, "fn", "")).lower().endswith("fitting")
            if fitting:
                self.setup_optimizers(trainer)
                try:
                    from lightning.pytorch.utilities.optimizer import _optimizers_to_device
                except ImportError:  # older layouts
                    from pytorch_lightning.utilities.optimizer import _optimizers_to_device
                _optimizers_to_device(self.optimizers, self.root_device)
                self._configure_bagua_model()
            self.setup_precision_plugin()

        def _configure_bagua_model(self):
            from ..parallel.data_parallel import DistributedDataParallel

            wrapped = _UnwrapForward(self.lightning_module)
            algo = _build_algorithm(self._bagua_algorithm, self.optimizers, self._bagua_kwargs)
            self.model = DistributedDataParallel(wrapped, optimizers=list(self.optimizers), algorithm=algo,
                                                 gradient_as_bucket_view=self._bagua_flatten)

        def configure_ddp(self):  # the base class would wrap with torch DDP
            pass

        def teardown(self):
            algo = getattr(getattr(self.model, "inner", None), "bagua_algorithm", None)
            if algo is not None and hasattr(algo, "abort"):
                algo.abort(self.model)  # stop the asynchronous averaging thread before the process group goes away
            super().teardown()

        @classmethod
        def register_strategies(cls, strategy_registry) -> None:
            strategy_registry.register(cls.strategy_name, cls, description="bagua_b200 data-parallel engine")

    _BaguaStrategy.__name__ = "BaguaStrategy"
    return _BaguaStrategy


class _LazyStrategy:
    """``BaguaStrategy(...)`` resolves the Lightning base class at call time."""

    def __call__(self, *args, **kwargs):
        return make_bagua_strategy()(*args, **kwargs)

    def __repr__(self):
        return "<BaguaStrategy factory (needs lightning)>"


BaguaStrategy = _LazyStrategy()
