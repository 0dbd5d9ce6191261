"""Reference module path ``dgl.multiprocessing.pytorch``: the names below are implemented in
dgl_b200.multiprocessing, dgl_b200 ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.multiprocessing', 'dgl_b200'],
                                          ['Process', 'call_once_and_share', 'shared_tensor', 'thread_wrapped_func'])
