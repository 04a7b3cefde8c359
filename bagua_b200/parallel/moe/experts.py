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

    def _grouped_mlp(self):
        """(fc1, fc2, activation) lists when every local expert is a 2-layer MLP ``fc2(act(fc1(x)))`` — the shape the
        tcgen05 grouped GEMM accelerates; ``None`` otherwise."""
        cached = getattr(self, "_grouped_cache", None)
        if cached is not None:
            return cached or None
        ok = True
        for e in self.bagua_experts:
            fc1, fc2 = getattr(e, "fc1", None), getattr(e, "fc2", None)
            if not (isinstance(fc1, torch.nn.Linear) and isinstance(fc2, torch.nn.Linear) and getattr(e, "grouped_gemm_compatible", False)):
                ok = False
        self._grouped_cache = ok
        return ok or None

    def _forward_grouped(self, inputs: torch.Tensor) -> torch.Tensor:
        from ...ops.gemm import grouped_linear

        W, E, C, M = inputs.shape
        x = inputs.permute(1, 0, 2, 3).reshape(E, W * C, M)  # tokens of every source rank, per local expert
        w1 = torch.stack([e.fc1.weight for e in self.bagua_experts])
        b1 = torch.stack([e.fc1.bias for e in self.bagua_experts]) if self.bagua_experts[0].fc1.bias is not None else None
        w2 = torch.stack([e.fc2.weight for e in self.bagua_experts])
        b2 = torch.stack([e.fc2.bias for e in self.bagua_experts]) if self.bagua_experts[0].fc2.bias is not None else None
        h = torch.nn.functional.gelu(grouped_linear(x, w1, b1), approximate="tanh")
        y = grouped_linear(h, w2, b2)
        return y.reshape(E, W, C, -1).permute(1, 0, 2, 3)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        """``inputs``: ``[world, num_local_experts, capacity, model]``; expert ``i`` processes ``inputs[:, i]``."""
        if inputs.is_cuda and inputs.dtype == torch.bfloat16 and self._grouped_mlp():
            return self._forward_grouped(inputs)
        outs = []
        for chunk, expert in zip(inputs.chunk(self.num_local_experts, dim=1), self.bagua_experts):
            out = expert(chunk)
            if isinstance(out, tuple):
                out = out[0]
            outs.append(out)
        return torch.cat(outs, dim=1)
