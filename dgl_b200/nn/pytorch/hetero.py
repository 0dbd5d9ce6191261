"""reference nn/pytorch/hetero.py: HeteroGraphConv, HeteroLinear, HeteroEmbedding."""
from .. import HeteroEmbedding, HeteroGraphConv, HeteroLinear  # noqa: F401
from ..conv.relational import get_aggregate_fn  # noqa: F401,E402
