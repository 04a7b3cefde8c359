"""Container of a rank's local experts (reference: bagua/torch_api/model_parallel/moe/experts.py:1-41)."""
from __future__ import annotations

import copy

import torch


class Experts(torch.nn.Module):
    def __init__(self, expert: torch.nn.Module, num_local_experts: int = 1):
        super().__init__()
        self.bagua_experts = torch.nn.ModuleList([copy.deepcopy(expert) for _ in range(num_local_experts)])
        self.num_local_experts = num_local_experts
        for e in self.bagua_experts:
            for _, p in e.named_parameters():
                p.expert = True  # skipped by the data-parallel engine (bagua_build_params)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        """``inputs``: ``[world, num_local_experts, capacity, model]``; expert ``i`` processes ``inputs[:, i]``."""
        outs = []
        for chunk, expert in zip(inputs.chunk(self.num_local_experts, dim=1), self.bagua_experts):
            out = expert(chunk)
            if isinstance(out, tuple):
                out = out[0]
            outs.append(out)
        return torch.cat(outs, dim=1)
