"""Graph convolution layers (parity: /root/reference/python/dgl/nn/pytorch/conv/__init__.py)."""
from .graphconv import EdgeWeightNorm, GraphConv
from .sageconv import SAGEConv

__all__ = ["GraphConv", "EdgeWeightNorm", "SAGEConv"]
