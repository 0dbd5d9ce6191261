"""Reference module path ``dgl.function.message``: the names below are implemented in
dgl_b200.function, dgl_b200 ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.function', 'dgl_b200'],
                                          ['BinaryMessageFunction', 'CopyMessageFunction', 'MessageFunction', 'copy_e', 'copy_u'])
