import sys as _sys

from bagua_b200.service import AutotuneClient, AutotuneService  # noqa: F401
from bagua_b200.service import autotune_service, autotune_system, autotune_task_manager, bayesian_optimizer  # noqa: F401

for _n in ("autotune_service", "autotune_system", "autotune_task_manager", "bayesian_optimizer"):
    _sys.modules[f"{__name__}.{_n}"] = globals()[_n]
