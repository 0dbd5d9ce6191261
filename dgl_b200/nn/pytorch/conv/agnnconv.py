"""Reference module path ``dgl.nn.pytorch.conv.agnnconv``: the names below are implemented in
dgl_b200.nn.pytorch.conv, dgl_b200.nn.pytorch, dgl_b200.nn ... and re-exported lazily."""
from ...._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.nn.pytorch.conv', 'dgl_b200.nn.pytorch', 'dgl_b200.nn', 'dgl_b200.nn.conv', 'dgl_b200'],
                                          ['AGNNConv'])
