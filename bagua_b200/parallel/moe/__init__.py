"""Expert parallelism (reference: bagua/torch_api/model_parallel/moe/)."""
from .layer import MoE  # noqa: F401
from .utils import is_moe_param  # noqa: F401
from .experts import Experts  # noqa: F401
from .sharded_moe import MOELayer, TopKGate, top1gating, top2gating  # noqa: F401
