"""Autograd-aware collectives (reference: bagua/torch_api/data_parallel/functional.py:1-79)."""
from __future__ import annotations

import torch

from ... import communication as comm_mod
from ...communication import ReduceOp

__all__ = ["all_reduce", "torch_reduce_op_to_bagua"]


def torch_reduce_op_to_bagua(op) -> ReduceOp:
    """``torch.distributed.ReduceOp`` → :class:`bagua_b200.ReduceOp` (reference functional.py:22-28, extended beyond SUM/MAX)."""
    import torch.distributed as dist

    table = {dist.ReduceOp.SUM: ReduceOp.SUM, dist.ReduceOp.MAX: ReduceOp.MAX, dist.ReduceOp.MIN: ReduceOp.MIN, dist.ReduceOp.PRODUCT: ReduceOp.PRODUCT}
    if hasattr(dist.ReduceOp, "AVG"):
        table[dist.ReduceOp.AVG] = ReduceOp.AVG
    for k, v in table.items():
        if op == k:
            return v
    raise ValueError(f"Unexpect input={op}")


class _AllReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, op, group, tensor):
        ctx.group = group
        ctx.op = op
        out = tensor.clone().contiguous()
        comm = group.get_global_communicator() if group is not None else None
        comm_mod.allreduce_inplace(out, op=op, comm=comm)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        return (None, None) + (_AllReduce.apply(ctx.op, ctx.group, grad_output.contiguous()),)


def all_reduce(tensor: torch.Tensor, op: ReduceOp = ReduceOp.SUM, group=None) -> torch.Tensor:
    """All-reduce ``tensor`` across ``group``; differentiable (the gradient is all-reduced with the same op)."""
    return _AllReduce.apply(op, group, tensor)
