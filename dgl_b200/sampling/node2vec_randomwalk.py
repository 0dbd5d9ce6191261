"""Reference module path ``dgl.sampling.node2vec_randomwalk``: the names below are implemented in
dgl_b200.sampling, dgl_b200 ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.sampling', 'dgl_b200'],
                                          ['node2vec_random_walk'])
