"""Decentralized SGD, full and low precision (reference: bagua/torch_api/algorithms/decentralized.py:12-214).

Full precision: the *weights* are communicated as one bucket for the whole model, the exchange starts in the forward-pre
hook and overlaps forward+backward, the averaged copy is swapped in after backward and before ``optimizer.step()``.
On NVSwitch the exchange is one kernel: ``shift_one`` reads the partner's weights straight over NVLink and writes
``(mine+theirs)/2``; ``all`` is an out-of-place two-shot / multimem allreduce into the symmetric ``peer_weight`` replica.

Low precision: ring of compressed weight *differences* after each optimizer step, one fused kernel per bucket."""
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


class DecentralizedAlgorithmImpl(AlgorithmImpl):
    def __init__(self, process_group, hierarchical: bool = True, peer_selection_mode: str = "all", communication_interval: int = 1):
        super().__init__(process_group)
        self.hierarchical = hierarchical
        self.peer_selection_mode = peer_selection_mode
        self.communication_interval = communication_interval

    def _should_communicate(self, bagua_ddp) -> bool:
        cur_step = bagua_ddp.bagua_train_step_counter - 1
        return cur_step % self.communication_interval == 0

    def init_tensors(self, bagua_ddp) -> List[torch.Tensor]:
        parameters = bagua_ddp.bagua_build_params()
        self.tensors = [param.ensure_bagua_tensor(name, bagua_ddp.bagua_module_name) for name, param in reversed(parameters)]
        return self.tensors

    def tensors_to_buckets(self, tensors: List[List[torch.Tensor]], do_flatten: bool) -> List[BaguaBucket]:
        # one bucket for the whole model (reference decentralized.py:52-61); 16-byte granularity for the peer kernels
        all_tensors = [t for b in tensors for t in b]
        align = max(1, 16 // all_tensors[0].element_size()) * self.process_group.size()
        return [BaguaBucket(all_tensors, flatten=do_flatten, name="0", alignment=align, group=self.process_group)]

    def init_forward_pre_hook(self, bagua_ddp):
        def hook(input):
            if self._should_communicate(bagua_ddp):
                for t in self.tensors:
                    bagua_ddp.mark_tensor_ready(t)

        return hook

    def init_backward_hook(self, bagua_ddp):
        def hook(parameter_name, parameter):
            return

        return hook

    def init_post_backward_hook(self, bagua_ddp):
        def hook():
            if self._should_communicate(bagua_ddp):
                bagua_ddp.wait_pending_comm_ops()
                # stream-ordered: the copy-back queues behind the wait on the compute stream, no host sync
                for bucket in bagua_ddp.bagua_buckets:
                    bucket._decentralized_op.copy_back_peer_weight(bucket)

        return hook

    def _init_states(self, bucket: BaguaBucket):
        bucket._peer_weight = bucket.new_companion(init_from_bucket=True, symmetric=True, group=self.process_group).ensure_bagua_tensor(
            "peer_weight", bucket.bagua_module_name
        )

    def init_operations(self, bagua_ddp, bucket: BaguaBucket):
        bucket.clear_ops()
        self._init_states(bucket)
        bucket._decentralized_op = bucket.append_decentralized_synchronous_op(
            peer_weight=bucket._peer_weight,
            hierarchical=self.hierarchical,
            peer_selection_mode=self.peer_selection_mode,
            group=self.process_group,
        )


class LowPrecisionDecentralizedAlgorithmImpl(AlgorithmImpl):
    def __init__(self, process_group, hierarchical: bool = True, communication_interval: int = 1):
        super().__init__(process_group)
        self.hierarchical = hierarchical
        self.communication_interval = communication_interval

    def _should_communicate(self, bagua_ddp) -> bool:
        cur_step = bagua_ddp.bagua_train_step_counter - 1
        return cur_step % self.communication_interval == 0

    def init_tensors(self, bagua_ddp) -> List[torch.Tensor]:
        parameters = bagua_ddp.bagua_build_params()
        self.tensors = [param.ensure_bagua_tensor(name, bagua_ddp.bagua_module_name) for name, param in reversed(parameters)]
        optimizer_param_ids = {id(p) for opt in bagua_ddp.bagua_optimizers for g in opt.param_groups for p in g["params"]}
        for name, param in parameters:
            if id(param) not in optimizer_param_ids:
                raise RuntimeError(
                    f"Module parameter {name} is not used by your optimizer(s), need to exclude it "
                    "by adding the parameter name to the `List` attribute `_bagua_params_and_buffers_to_ignore` "
                    "of your module."
                )
        return self.tensors

    def tensors_to_buckets(self, tensors: List[List[torch.Tensor]], do_flatten: bool) -> List[BaguaBucket]:
        return [BaguaBucket(b, flatten=do_flatten, name=str(i), alignment=32, group=self.process_group) for i, b in enumerate(tensors)]

    def init_backward_hook(self, bagua_ddp):
        def hook(parameter_name, parameter):
            pass

        return hook

    def init_post_backward_hook(self, bagua_ddp):
        def hook():
            pass

        return hook

    def init_post_optimizer_step_hook(self, bagua_ddp):
        from ...contrib.fuse.optimizer import is_fused_optimizer

        def hook(optimizer: torch.optim.Optimizer):
            assert not is_fused_optimizer(optimizer), "Low decentralized algorithm can not work with fused optimizer at present."
            if self._should_communicate(bagua_ddp):
                for group in optimizer.param_groups:
                    for param in group["params"]:
                        if param.is_bagua_tensor():
                            bagua_ddp.mark_tensor_ready(param)
                bagua_ddp.wait_pending_comm_ops()

        return hook

    def _init_states(self, bucket: BaguaBucket):
        name = bucket.bagua_module_name
        bucket._weight = bucket.new_companion().ensure_bagua_tensor("weight", name)
        bucket._left_peer_weight = bucket.new_companion().ensure_bagua_tensor("left_peer_weight", name)
        bucket._right_peer_weight = bucket.new_companion().ensure_bagua_tensor("right_peer_weight", name)

    def init_operations(self, bagua_ddp, bucket: BaguaBucket):
        bucket.clear_ops()
        self._init_states(bucket)
        bucket.append_low_precision_decentralized_synchronous_op(
            weight=bucket._weight,
            left_peer_weight=bucket._left_peer_weight,
            right_peer_weight=bucket._right_peer_weight,
            hierarchical=self.hierarchical,
            compression="MinMaxUInt8",
            group=self.process_group,
        )


class DecentralizedAlgorithm(Algorithm):
    def __init__(self, hierarchical: bool = True, peer_selection_mode: str = "all", communication_interval: int = 1):
        """
        Args:
            hierarchical: enable hierarchical communication across nodes.
            peer_selection_mode: ``"all"`` (average over every worker) or ``"shift_one"`` (pairwise, partner changes each step).
            communication_interval: iterations between two communication steps.
        """
        self.hierarchical = hierarchical
        self.peer_selection_mode = peer_selection_mode
        self.communication_interval = communication_interval

    def reify(self, process_group) -> DecentralizedAlgorithmImpl:
        return DecentralizedAlgorithmImpl(
            process_group,
            hierarchical=self.hierarchical,
            peer_selection_mode=self.peer_selection_mode,
            communication_interval=self.communication_interval,
        )


class LowPrecisionDecentralizedAlgorithm(Algorithm):
    def __init__(self, hierarchical: bool = True, communication_interval: int = 1):
        self.hierarchical = hierarchical
        self.communication_interval = communication_interval

    def reify(self, process_group) -> LowPrecisionDecentralizedAlgorithmImpl:
        return LowPrecisionDecentralizedAlgorithmImpl(
            process_group, hierarchical=self.hierarchical, communication_interval=self.communication_interval
        )
