"""Distributed sparse optimizers for ``DistEmbedding`` (parity: python/dgl/distributed/optim/pytorch/sparse_optim.py)."""
from ..dist_graph import DistSparseAdagrad as SparseAdagrad  # noqa: F401
from ..dist_graph import DistSparseAdam as SparseAdam  # noqa: F401
from ..dist_graph import DistSparseGradOptimizer  # noqa: F401
