"""Type definitions shared by the engine and the autotune service
(reference: bagua/bagua_define.py:1-58).  ``bf16`` is new: the reference rejects it
(bagua/torch_api/utils.py:81-92)."""
from __future__ import annotations

import enum
from typing import List

try:  # python ≥ 3.8
    from typing import TypedDict
except ImportError:  # pragma: no cover
    from typing_extensions import TypedDict

from pydantic import BaseModel


class TensorDtype(str, enum.Enum):
    F32 = "f32"
    F16 = "f16"
    BF16 = "bf16"
    U8 = "u8"
    I64 = "i64"


class TensorDeclaration(TypedDict):
    name: str
    num_elements: int
    dtype: TensorDtype


_DTYPE_BYTES = {TensorDtype.F32: 4, TensorDtype.F16: 2, TensorDtype.BF16: 2, TensorDtype.U8: 1, TensorDtype.I64: 8}


def get_tensor_declaration_bytes(td: TensorDeclaration) -> int:
    """Size in bytes of the tensor a declaration describes (reference bagua_define.py:24-35; bf16 added)."""
    dtype = td["dtype"]
    if not isinstance(dtype, TensorDtype):
        dtype = TensorDtype(dtype)
    return td["num_elements"] * _DTYPE_BYTES[dtype]


def _as_declaration(td) -> TensorDeclaration:
    d = dict(td)
    if not isinstance(d.get("dtype"), TensorDtype):
        d["dtype"] = TensorDtype(d["dtype"])
    return TensorDeclaration(**d)


class BaguaCoreTelemetrySpan(BaseModel):
    trace_id: int
    action: str
    tensor_name: str
    start_time: int
    end_time: int


class BaguaHyperparameter(BaseModel):
    """What the autotune service hands to the engine.  ``allreduce_variant`` and ``comm_blocks`` extend the
    reference's (buckets, bucket_size, is_hierarchical_reduce) with the NVSwitch kernel choice."""

    buckets: List[List[TensorDeclaration]] = []
    bucket_size: int = 0
    is_hierarchical_reduce: bool = False
    allreduce_variant: str = "auto"
    comm_blocks: int = 0
    # Per bucket (same order as ``buckets``): the kernel variant / CTA count the service looked up, for that bucket's message
    # size, in the bus-bandwidth table the workers MEASURED on their fabric at start-up (PeerEngine.calibrate). Empty = none measured.
    bucket_variants: List[str] = []
    bucket_blocks: List[int] = []

    def update(self, param_dict: dict) -> "BaguaHyperparameter":
        tmp = self.model_dump()
        for key, value in param_dict.items():
            if key in tmp:
                self.__dict__[key] = value
        for key, value in param_dict.items():
            if key in tmp and key == "buckets":
                self.buckets = [[_as_declaration(td) for td in b] for b in value]
        return self

    # pydantic v1 spelling used by the reference's call sites
    def dict(self, *a, **k):  # type: ignore[override]
        return self.model_dump(*a, **k)
