import torch


def is_moe_param(param: torch.Tensor) -> bool:
    """Expert parameters are local to their rank and excluded from data-parallel communication."""
    return bool(getattr(param, "expert", False))
