"""Launchers: ``python -m bagua_b200.distributed.launch`` (static) and ``python -m bagua_b200.distributed.run`` (elastic)."""
