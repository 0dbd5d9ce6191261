"""Reference module path ``dgl.sparse.elementwise_op_sp``: the names below are implemented in
dgl_b200.sparse, dgl_b200 ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.sparse', 'dgl_b200'],
                                          ['sp_add', 'sp_div', 'sp_mul', 'sp_power', 'sp_sub', 'spsp_add', 'spsp_div', 'spsp_mul'])
