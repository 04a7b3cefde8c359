"""ByteGrad: 8-bit MinMaxUInt8-compressed gradient allreduce (reference: bagua/torch_api/algorithms/bytegrad.py:1-82)."""
from __future__ import annotations

from typing import List

import torch

from ...bucket import BaguaBucket
from .base import Algorithm, AlgorithmImpl

__all__ = ["ByteGradAlgorithm", "ByteGradAlgorithmImpl"]


class ByteGradAlgorithmImpl(AlgorithmImpl):
    def __init__(self, process_group, hierarchical: bool = True, average: bool = True):
        super().__init__(process_group)
        self.hierarchical = hierarchical
        self.average = average

    def tensors_to_buckets(self, tensors: List[List[torch.Tensor]], do_flatten: bool) -> List[BaguaBucket]:
        """Buckets are padded so every rank owns an equal chunk; the fused kernel additionally wants 32-element chunk
        granularity (16-byte payload vectors + 32-byte wire alignment), which is still a multiple of ``nranks``
        (reference pads to ``nranks``, bytegrad.py:33-45)."""
        n = self.process_group.size()
        return [BaguaBucket(b, flatten=do_flatten, name=str(i), alignment=32 * n, group=self.process_group) for i, b in enumerate(tensors)]

    def init_operations(self, bagua_ddp, bucket):
        bucket.clear_ops()
        bucket.append_centralized_synchronous_op(
            hierarchical=self.hierarchical, average=self.average, scattergather=True, compression="MinMaxUInt8", group=self.process_group
        )


class ByteGradAlgorithm(Algorithm):
    def __init__(self, hierarchical: bool = True, average: bool = True):
        self.hierarchical = hierarchical
        self.average = average

    def reify(self, process_group) -> ByteGradAlgorithmImpl:
        return ByteGradAlgorithmImpl(process_group, hierarchical=self.hierarchical, average=self.average)
