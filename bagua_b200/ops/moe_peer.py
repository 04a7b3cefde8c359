"""NVSwitch path of MoE dispatch/combine (kernels in csrc/moe_kernels.cu).  Filled in by the peer-kernel milestone;
until the kernels are registered ``get_context`` reports that the fused path is unavailable."""
from __future__ import annotations

_contexts = {}


def get_context(group, world: int):
    return None
