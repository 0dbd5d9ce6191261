"""Reference module path ``dgl.data.tensor_serialize``: the names below are implemented in
dgl_b200.data, dgl_b200 ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.data', 'dgl_b200'],
                                          ['load_tensors', 'save_tensors'])
