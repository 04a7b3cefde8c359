"""Compatibility alias: user code written for the reference (``import bagua.torch_api as bagua``) runs on bagua_b200."""
from bagua_b200 import __version__  # noqa: F401
from . import bagua_define  # noqa: F401
