"""``python -m bagua.script.bagua_sys_perf`` — alias of :mod:`bagua_b200.script.bagua_sys_perf` (reference: bagua/script/bagua_sys_perf:19-60)."""
from bagua_b200.script.bagua_sys_perf import *  # noqa: F401,F403
from bagua_b200.script.bagua_sys_perf import main  # noqa: F401

if __name__ == "__main__":
    import sys

    sys.exit(main())
