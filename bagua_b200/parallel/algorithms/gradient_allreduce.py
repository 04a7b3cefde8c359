"""Centralized synchronous full-precision data parallelism (reference: bagua/torch_api/algorithms/gradient_allreduce.py:1-64).

On NVSwitch each bucket is reduced by ONE kernel (two-shot peer loads/stores or NVLS ``multimem``) that also applies
the 1/N average; ``fused_optimizer=True`` additionally folds the SGD update of the rank's 1/N shard and the all-gather
of the updated weights into that kernel (see ``FusedShardedSGD``)."""
from __future__ import annotations

from .base import Algorithm, AlgorithmImpl

__all__ = ["GradientAllReduceAlgorithm", "GradientAllReduceAlgorithmImpl"]


class GradientAllReduceAlgorithmImpl(AlgorithmImpl):
    def __init__(self, process_group, hierarchical: bool = False, average: bool = True, variant: str = "auto"):
        super().__init__(process_group)
        self.hierarchical = hierarchical
        self.average = average
        self.variant = variant

    def init_operations(self, bagua_ddp, bucket):
        bucket.clear_ops()
        hp = getattr(bagua_ddp, "_bagua_hyperparameters", None)
        variant = self.variant
        if variant == "auto" and hp is not None and getattr(hp, "allreduce_variant", "auto") != "auto":
            variant = hp.allreduce_variant
        bucket.append_centralized_synchronous_op(hierarchical=self.hierarchical, average=self.average, group=self.process_group, variant=variant)


class GradientAllReduceAlgorithm(Algorithm):
    def __init__(self, hierarchical: bool = False, average: bool = True, variant: str = "auto"):
        """
        Args:
            hierarchical: intra-node reduce → inter-node allreduce → intra-node broadcast when the job spans nodes.
            average: average (``True``) or sum gradients.
            variant: ``auto`` | ``one_shot`` | ``two_shot`` | ``multimem`` kernel choice (``auto`` = by size / autotune).
        """
        self.hierarchical = hierarchical
        self.average = average
        self.variant = variant

    def reify(self, process_group) -> GradientAllReduceAlgorithmImpl:
        return GradientAllReduceAlgorithmImpl(process_group, hierarchical=self.hierarchical, average=self.average, variant=self.variant)
