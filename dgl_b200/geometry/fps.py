"""Reference module path ``dgl.geometry.fps``: the names below are implemented in
dgl_b200.geometry, dgl_b200 ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.geometry', 'dgl_b200'],
                                          ['farthest_point_sampler'])
