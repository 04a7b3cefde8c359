"""``python -m bagua.script.baguarun`` — alias of :mod:`bagua_b200.script.baguarun` (reference: bagua/script/baguarun.py:115-209)."""
from bagua_b200.script.baguarun import *  # noqa: F401,F403
from bagua_b200.script.baguarun import main  # noqa: F401

if __name__ == "__main__":
    import sys

    sys.exit(main())
