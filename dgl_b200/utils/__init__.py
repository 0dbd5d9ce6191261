"""Utilities (parity: /root/reference/python/dgl/utils/)."""
from .internal import *  # noqa: F401,F403
from .pin_memory import gather_pinned_tensor_rows, pin_memory_inplace, scatter_pinned_tensor_rows, unpin_memory_inplace  # noqa: F401
from . import trace  # noqa: F401,E402
