from .autotune_service import AutotuneClient, AutotuneService  # noqa: F401
