"""User-facing MoE layer (reference: bagua/torch_api/model_parallel/moe/layer.py:22-108)."""
from __future__ import annotations

import logging
import typing

import torch
import torch.distributed as dist

from ... import env
from .experts import Experts
from .sharded_moe import MOELayer, TopKGate


class MoE(torch.nn.Module):
    def __init__(self, hidden_size, expert, num_local_experts=1, k=1, output_dropout_prob=0.0, capacity_factor=1.0, eval_capacity_factor=1.0,
                 min_capacity=4, noisy_gate_policy: typing.Optional[str] = None):
        """
        Args:
            hidden_size: model dimension (input and output of the layer).
            expert: module defining one expert (e.g. an MLP); deep-copied ``num_local_experts`` times.
            num_local_experts: experts hosted by each rank; total experts = ``num_local_experts * world_size``.
            k: top-k gating (1 or 2).
            output_dropout_prob: dropout on the layer output.
            capacity_factor / eval_capacity_factor: expert capacity multiplier at training / evaluation time.
            min_capacity: lower bound of the per-expert capacity.
            noisy_gate_policy: ``None`` | ``'Jitter'`` | ``'RSample'``.
        """
        super().__init__()
        assert noisy_gate_policy is None or noisy_gate_policy in ["None", "Jitter", "RSample"], "Unsupported noisy_gate_policy: " + str(noisy_gate_policy)
        world = dist.get_world_size() if dist.is_initialized() else env.get_world_size()
        self.num_experts = num_local_experts * world
        logging.info(f"num_experts: {self.num_experts} | num_local_experts: {num_local_experts} | world_size: {world}")
        experts = Experts(expert, num_local_experts)
        self.bagua_moe = MOELayer(
            TopKGate(hidden_size, self.num_experts, k, capacity_factor, eval_capacity_factor, min_capacity, noisy_gate_policy),
            experts,
            num_local_experts,
            group=dist.group.WORLD if dist.is_initialized() else None,
        )
        self.dropout = torch.nn.Dropout(output_dropout_prob)

    def forward(self, hidden_states, used_token=None):
        """Returns ``(output, l_aux, exp_counts)``."""
        output = self.bagua_moe(hidden_states, used_token)
        output = self.dropout(output)
        return output, self.bagua_moe.l_aux, self.bagua_moe.exp_counts
