from .optimizer import fuse_optimizer, is_fused_optimizer  # noqa: F401
