"""reference nn/pytorch/hetero.py: HeteroGraphConv, HeteroLinear, HeteroEmbedding."""
from .. import HeteroEmbedding, HeteroGraphConv, HeteroLinear  # noqa: F401
