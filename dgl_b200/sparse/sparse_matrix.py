"""Reference module path ``dgl.sparse.sparse_matrix``: the names below are implemented in
dgl_b200.sparse, dgl_b200 ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.sparse', 'dgl_b200'],
                                          ['SparseMatrix', 'diag', 'from_coo', 'from_csc', 'from_csr', 'from_torch_sparse', 'identity', 'spmatrix', 'to_torch_sparse_coo', 'to_torch_sparse_csc', 'to_torch_sparse_csr', 'val_like'])
