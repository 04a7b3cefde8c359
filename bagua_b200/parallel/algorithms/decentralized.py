"""Decentralized SGD, full and low precision (reference: bagua/torch_api/algorithms/decentralized.py:12-214).

Full precision: the *weights* are communicated as one bucket for the whole model, the exchange starts in the forward-pre
hook and overlaps forward+backward, the averaged copy is swapped in after backward and before ``optimizer.step()``.
On NVSwitch the exchange is one kernel: ``shift_one`` reads the partner's weights straight over NVLink and writes
``(mine+theirs)/2``; ``all`` is an out-of-place two-shot / multimem allreduce into the symmetric ``peer_weight`` replica.

Low precision: ring of compressed weight *differences* after each optimizer step, one fused kernel per bucket.

Both variants communicate WEIGHTS (not gradients), every ``communication_interval``-th iteration; what differs is when the
exchange is triggered and what the bucket program is — that is all the two ``*Impl`` classes below spell out, the rest lives in
:class:`_WeightExchange`."""
from __future__ import annotations

from typing import List

import torch

from ...bucket import BaguaBucket
from .base import Algorithm, AlgorithmImpl

__all__ = [
    "DecentralizedAlgorithm",
    "DecentralizedAlgorithmImpl",
    "LowPrecisionDecentralizedAlgorithm",
    "LowPrecisionDecentralizedAlgorithmImpl",
]


def _nothing(*_args, **_kwargs):
    """Hook slot that a weight-communicating algorithm leaves empty (gradients are never marked ready)."""


class _WeightExchange(AlgorithmImpl):
    """What the two decentralized implementations share: parameters (not gradients) are the communicated tensors, registered in
    reverse order like everywhere else, and an iteration takes part in communication when ``(step − 1) % interval == 0``."""

    def __init__(self, process_group, hierarchical: bool, communication_interval: int):
        super().__init__(process_group)
        self.hierarchical, self.communication_interval = hierarchical, communication_interval

    def _should_communicate(self, bagua_ddp) -> bool:
        return (bagua_ddp.bagua_train_step_counter - 1) % self.communication_interval == 0

    def _register_weights(self, bagua_ddp):
        named = bagua_ddp.bagua_build_params()
        self.tensors = [p.ensure_bagua_tensor(n, bagua_ddp.bagua_module_name) for n, p in reversed(named)]
        return named

    def init_tensors(self, bagua_ddp) -> List[torch.Tensor]:
        self._register_weights(bagua_ddp)
        return self.tensors

    def init_backward_hook(self, bagua_ddp):
        return _nothing


class DecentralizedAlgorithmImpl(_WeightExchange):
    def __init__(self, process_group, hierarchical: bool = True, peer_selection_mode: str = "all", communication_interval: int = 1):
        super().__init__(process_group, hierarchical, communication_interval)
        self.peer_selection_mode = peer_selection_mode

    def tensors_to_buckets(self, tensors: List[List[torch.Tensor]], do_flatten: bool) -> List[BaguaBucket]:
        """ONE bucket for the whole model (reference :60-66), aligned so every rank owns a 16-byte multiple of it."""
        everything = [t for suggested in tensors for t in suggested]
        align = max(1, 16 // everything[0].element_size()) * self.process_group.size()
        return [BaguaBucket(everything, flatten=do_flatten, name="0", alignment=align, group=self.process_group)]

    def init_forward_pre_hook(self, bagua_ddp):
        def start_exchange(_inputs):
            # the weights are final once the previous optimizer step has run: the exchange overlaps this forward and backward
            if self._should_communicate(bagua_ddp):
                for t in self.tensors:
                    bagua_ddp.mark_tensor_ready(t)

        return start_exchange

    def init_post_backward_hook(self, bagua_ddp):
        def swap_in_average():
            if not self._should_communicate(bagua_ddp):
                return
            bagua_ddp.wait_pending_comm_ops()
            for b in bagua_ddp.bagua_buckets:    # averaged replica → live weights, before optimizer.step() applies the local gradients
                b._decentralized_op.copy_back_peer_weight(b)

        return swap_in_average

    def _init_states(self, bucket: BaguaBucket):
        replica = bucket.new_companion(init_from_bucket=True, symmetric=True, group=self.process_group)
        bucket._peer_weight = replica.ensure_bagua_tensor("peer_weight", bucket.bagua_module_name)

    def init_operations(self, bagua_ddp, bucket: BaguaBucket):
        bucket.clear_ops()
        self._init_states(bucket)
        bucket._decentralized_op = bucket.append_decentralized_synchronous_op(peer_weight=bucket._peer_weight, hierarchical=self.hierarchical,
                                                                              peer_selection_mode=self.peer_selection_mode, group=self.process_group)


class LowPrecisionDecentralizedAlgorithmImpl(_WeightExchange):
    def __init__(self, process_group, hierarchical: bool = True, communication_interval: int = 1):
        super().__init__(process_group, hierarchical, communication_interval)

    def init_tensors(self, bagua_ddp) -> List[torch.Tensor]:
        """The ring exchanges what the optimizer has just written, so every communicated parameter must belong to an optimizer."""
        named = self._register_weights(bagua_ddp)
        stepped = {id(p) for opt in bagua_ddp.bagua_optimizers for g in opt.param_groups for p in g["params"]}
        orphans = [n for n, p in named if id(p) not in stepped]
        if orphans:
            raise RuntimeError(f"Module parameter {orphans[0]} is not used by your optimizer(s), need to exclude it by adding the parameter name to the "
                               "`List` attribute `_bagua_params_and_buffers_to_ignore` of your module.")
        return self.tensors

    def tensors_to_buckets(self, tensors: List[List[torch.Tensor]], do_flatten: bool) -> List[BaguaBucket]:
        return [BaguaBucket(b, flatten=do_flatten, name=str(i), alignment=32, group=self.process_group) for i, b in enumerate(tensors)]

    def init_post_backward_hook(self, bagua_ddp):
        return _nothing

    def init_post_optimizer_step_hook(self, bagua_ddp):
        from ...contrib.fuse.optimizer import is_fused_optimizer

        def exchange_after_step(optimizer: torch.optim.Optimizer):
            assert not is_fused_optimizer(optimizer), "Low decentralized algorithm can not work with fused optimizer at present."
            if not self._should_communicate(bagua_ddp):
                return
            for p in (p for g in optimizer.param_groups for p in g["params"]):
                if p.is_bagua_tensor():
                    bagua_ddp.mark_tensor_ready(p)
            bagua_ddp.wait_pending_comm_ops()

        return exchange_after_step

    def _init_states(self, bucket: BaguaBucket):
        owner = bucket.bagua_module_name
        for attr, name in (("_weight", "weight"), ("_left_peer_weight", "left_peer_weight"), ("_right_peer_weight", "right_peer_weight")):
            setattr(bucket, attr, bucket.new_companion().ensure_bagua_tensor(name, owner))

    def init_operations(self, bagua_ddp, bucket: BaguaBucket):
        bucket.clear_ops()
        self._init_states(bucket)
        bucket.append_low_precision_decentralized_synchronous_op(weight=bucket._weight, left_peer_weight=bucket._left_peer_weight,
                                                                 right_peer_weight=bucket._right_peer_weight, hierarchical=self.hierarchical,
                                                                 compression="MinMaxUInt8", group=self.process_group)


class DecentralizedAlgorithm(Algorithm):
    def __init__(self, hierarchical: bool = True, peer_selection_mode: str = "all", communication_interval: int = 1):
        """
        Args:
            hierarchical: enable hierarchical communication across nodes.
            peer_selection_mode: ``"all"`` (average over every worker) or ``"shift_one"`` (pairwise, partner changes each step).
            communication_interval: iterations between two communication steps.
        """
        self.hierarchical, self.peer_selection_mode, self.communication_interval = hierarchical, peer_selection_mode, communication_interval

    def reify(self, process_group) -> DecentralizedAlgorithmImpl:
        return DecentralizedAlgorithmImpl(process_group, self.hierarchical, self.peer_selection_mode, self.communication_interval)


class LowPrecisionDecentralizedAlgorithm(Algorithm):
    def __init__(self, hierarchical: bool = True, communication_interval: int = 1):
        """
        Args:
            hierarchical: enable hierarchical communication across nodes.
            communication_interval: iterations between two ring exchanges.
        """
        self.hierarchical, self.communication_interval = hierarchical, communication_interval

    def reify(self, process_group) -> LowPrecisionDecentralizedAlgorithmImpl:
        return LowPrecisionDecentralizedAlgorithmImpl(process_group, self.hierarchical, self.communication_interval)
