"""Reference module path ``dgl.ops.sddmm``: the names below are implemented in
dgl_b200.ops, dgl_b200 ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.ops', 'dgl_b200'],
                                          ['copy_e', 'copy_u', 'copy_v', 'gsddmm', 'reshape_lhs_rhs'])
