"""Sparse optimizers for NodeEmbedding (parity: /root/reference/python/dgl/optim/pytorch/sparse_optim.py)."""
from .sparse_optim import SparseAdagrad, SparseAdam, SparseGradOptimizer  # noqa: F401
