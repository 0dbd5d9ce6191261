"""Reference module path ``dgl.sparse.elementwise_op``: the names below are implemented in
dgl_b200.sparse, dgl_b200 ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.sparse', 'dgl_b200'],
                                          ['add', 'div', 'mul', 'power', 'sub'])
