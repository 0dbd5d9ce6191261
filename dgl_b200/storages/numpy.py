"""Reference module path ``dgl.storages.numpy``: the names below are implemented in
dgl_b200.storages, dgl_b200 ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.storages', 'dgl_b200'],
                                          ['NumpyStorage'])
