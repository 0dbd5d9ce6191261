"""Graph transforms (parity: /root/reference/python/dgl/transforms/__init__.py)."""
from .functional import *  # noqa: F401,F403
from .module import *  # noqa: F401,F403
