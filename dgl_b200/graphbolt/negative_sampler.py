"""Reference module path ``dgl.graphbolt.negative_sampler``: the names below are implemented in
dgl_b200.graphbolt, dgl_b200.graphbolt.compat, dgl_b200.graphbolt.impl ... and re-exported lazily."""
from .._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.graphbolt', 'dgl_b200.graphbolt.compat', 'dgl_b200.graphbolt.impl', 'dgl_b200.graphbolt.utils', 'dgl_b200.graphbolt.internal', 'dgl_b200.graphbolt.datapipes', 'dgl_b200.graphbolt.minibatch', 'dgl_b200.graphbolt.itemset', 'dgl_b200.graphbolt.base', 'dgl_b200'],
                                          ['NegativeSampler'])
