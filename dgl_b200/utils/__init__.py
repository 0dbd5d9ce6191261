"""Utilities (parity: /root/reference/python/dgl/utils/ — internal, checks, data, filter, exception, pin_memory,
shared_mem)."""
from .internal import *  # noqa: F401,F403
from .containers import *  # noqa: F401,F403
from .checks import *  # noqa: F401,F403
from .data import *  # noqa: F401,F403
from .exception import ExceptionWrapper, KeyErrorMessage  # noqa: F401
from .filter import Filter  # noqa: F401
from .pin_memory import gather_pinned_tensor_rows, pin_memory_inplace, scatter_pinned_tensor_rows, unpin_memory_inplace  # noqa: F401
from .shared_mem import *  # noqa: F401,F403
from . import trace  # noqa: F401,E402
