"""Version of the bagua-compatible API surface (reference: bagua/version.py)."""
from bagua_b200 import __version__  # noqa: F401
