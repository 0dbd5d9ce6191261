"""User-defined-function argument types (parity: /root/reference/python/dgl/udf.py): :class:`EdgeBatch` is what a
message / edge UDF receives, :class:`NodeBatch` what a reduce / apply-node UDF receives. Both are defined with the
dispatcher in :mod:`dgl_b200.core`; this module is their reference import path."""
from .core import EdgeBatch, NodeBatch  # noqa: F401

__all__ = ["EdgeBatch", "NodeBatch"]
