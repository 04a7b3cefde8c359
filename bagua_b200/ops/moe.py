"""MoE token dispatch / combine: autograd functions over row scatter/gather + the expert all-to-all.

``dispatch``: rows of ``tokens[S, M]`` → ``[world(src), E_local, C, M]`` on the rank owning each expert.
``combine``: ``out[s] = Σ_k w[s,k] · expert_out[expert(s,k), slot(s,k)]`` pulled back from the owning ranks.

Paths: (1) NVSwitch peer kernels (``csrc/moe_kernels.cu``): rows are stored/loaded straight in peer symmetric memory;
(2) ``torch.distributed.all_to_all_single`` around local index scatter/gather (CPU/gloo, multi-node, no engine).
The reference does dense one-hot einsums + ``all_to_all_single`` (sharded_moe.py:352-374)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

__all__ = ["dispatch", "combine"]


def _all_to_all(x: torch.Tensor, group, world: int) -> torch.Tensor:
    if world == 1 or not dist.is_initialized():
        return x
    out = torch.empty_like(x)
    dist.all_to_all_single(out, x.contiguous(), group=group)
    return out


class _AllToAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group, world):
        ctx.group, ctx.world = group, world
        return _all_to_all(x, group, world)

    @staticmethod
    def backward(ctx, grad):
        return _all_to_all(grad.contiguous(), ctx.group, ctx.world), None, None


def _flat_slots(expert_idx, slot_idx, capacity):
    """Flat row index e*C + c for every (token, k); dropped choices map to -1."""
    flat = expert_idx * capacity + slot_idx
    return torch.where(slot_idx >= 0, flat, torch.full_like(flat, -1))


class _ScatterRows(torch.autograd.Function):
    """rows[flat[s,k]] = tokens[s] (each destination row receives at most one token)."""

    @staticmethod
    def forward(ctx, tokens, flat, num_rows):
        S, K = flat.shape
        out = torch.zeros(num_rows, tokens.shape[1], dtype=tokens.dtype, device=tokens.device)
        valid = flat >= 0
        src = torch.arange(S, device=tokens.device).unsqueeze(1).expand(S, K)[valid]
        out.index_copy_(0, flat[valid], tokens.index_select(0, src))
        ctx.save_for_backward(flat)
        ctx.S = S
        return out

    @staticmethod
    def backward(ctx, grad_rows):
        (flat,) = ctx.saved_tensors
        S, K = flat.shape
        valid = flat >= 0
        g = torch.zeros(S, grad_rows.shape[1], dtype=grad_rows.dtype, device=grad_rows.device)
        src = torch.arange(S, device=grad_rows.device).unsqueeze(1).expand(S, K)[valid]
        g.index_add_(0, src, grad_rows.index_select(0, flat[valid]))
        return g, None, None


class _GatherRows(torch.autograd.Function):
    """out[s] = Σ_k w[s,k] · rows[flat[s,k]]."""

    @staticmethod
    def forward(ctx, rows, flat, weights):
        S, K = flat.shape
        safe = flat.clamp(min=0)
        picked = rows.index_select(0, safe.reshape(-1)).view(S, K, -1)
        w = (weights * (flat >= 0).to(weights.dtype)).unsqueeze(-1)
        ctx.save_for_backward(rows, flat, weights)
        return (picked * w).sum(dim=1)

    @staticmethod
    def backward(ctx, grad_out):
        rows, flat, weights = ctx.saved_tensors
        S, K = flat.shape
        valid = (flat >= 0)
        safe = flat.clamp(min=0)
        picked = rows.index_select(0, safe.reshape(-1)).view(S, K, -1)
        vw = valid.to(weights.dtype)
        grad_w = (picked * grad_out.unsqueeze(1)).sum(-1).to(weights.dtype) * vw
        grad_rows = torch.zeros_like(rows)
        contrib = (grad_out.unsqueeze(1) * (weights * vw).unsqueeze(-1)).reshape(S * K, -1)
        grad_rows.index_add_(0, safe.reshape(-1)[valid.reshape(-1)], contrib[valid.reshape(-1)].to(rows.dtype))
        return grad_rows, None, grad_w


def _peer_ctx(tokens: torch.Tensor, group, world: int):
    """The NVSwitch MoE context for this group, or None."""
    if not tokens.is_cuda or (world < 2 and os.environ.get("BAGUA_SELF_PEER", "0") != "1"):   # world 1 only in self-peer mode
        return None
    from . import moe_peer

    return moe_peer.get_context(group, world)


def dispatch(tokens, expert_idx, slot_idx, num_experts: int, capacity: int, group, world: int, num_local_experts: int, key=None) -> torch.Tensor:
    """Route token rows to their experts' capacity slots: ``tokens [S, M]`` → ``[world, E_local, C, M]`` on the expert owners (peer scatter
    kernel on one NVSwitch node, dense one-hot + ``all_to_all_single`` otherwise).  Differentiable."""
    ctx = _peer_ctx(tokens, group, world)
    if ctx is not None:
        # key (one per MoE layer): the layer owns its receive buffer and uses the rows in place — no staging copy (ops/moe_peer.py)
        return ctx.dispatch(tokens, expert_idx, slot_idx, num_experts, capacity, num_local_experts, key=key)
    flat = _flat_slots(expert_idx, slot_idx, capacity)
    rows = _ScatterRows.apply(tokens, flat, num_experts * capacity)            # [E_total*C, M]
    rows = _AllToAll.apply(rows.view(num_experts, capacity, -1), group, world)  # chunks of E_local experts per rank
    return rows.reshape(world, num_local_experts, capacity, -1)


def combine(expert_out, expert_idx, slot_idx, weights, num_experts: int, capacity: int, group, world: int, num_local_experts: int, key=None) -> torch.Tensor:
    """Inverse of :func:`dispatch`: fetch every token's expert outputs and sum them with the gate ``weights`` → ``[S, M]``.  Differentiable
    in ``expert_out`` and ``weights``."""
    ctx = _peer_ctx(expert_out, group, world)
    if ctx is not None:
        return ctx.combine(expert_out, expert_idx, slot_idx, weights, num_experts, capacity, num_local_experts, key=key)
    rows = _AllToAll.apply(expert_out.reshape(num_experts, capacity, -1), group, world)
    flat = _flat_slots(expert_idx, slot_idx, capacity)
    return _GatherRows.apply(rows.reshape(num_experts * capacity, -1), flat, weights)
