"""Reference module path ``dgl.optim.pytorch.sparse_optim``: the names below are implemented in
dgl_b200.optim, dgl_b200.optim.sparse_optim, dgl_b200 ... and re-exported lazily."""
from ..._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.optim', 'dgl_b200.optim.sparse_optim', 'dgl_b200'],
                                          ['SparseAdagrad', 'SparseAdam', 'SparseGradOptimizer'])
