"""Neural-network modules (``dgl.nn``, PyTorch only).
Parity: /root/reference/python/dgl/nn/pytorch/__init__.py."""
from .conv import *  # noqa: F401,F403
from . import functional  # noqa: F401
