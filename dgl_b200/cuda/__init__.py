"""CUDA helpers (parity: /root/reference/python/dgl/cuda/__init__.py, gpu_cache.py, nccl.py)."""
from ..graphbolt.impl.caches import GPUCache  # noqa: F401
from . import nccl  # noqa: F401
