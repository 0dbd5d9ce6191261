"""``dgl.nn.pytorch`` import path of the reference (``import dgl.nn.pytorch as dglnn``): the layers live in
``dgl_b200.nn``; this package re-exports them under the reference's module layout (conv, glob, hetero, linear, softmax,
factory, utils, sparse_emb, network_emb, link, gt, explain)."""
from .. import *  # noqa: F401,F403
from .. import conv  # noqa: F401
from . import explain, factory, glob, gt, hetero, linear, link, network_emb, softmax, sparse_emb, utils  # noqa: F401
from .softmax import edge_softmax  # noqa: F401


def __getattr__(name):
    # names dgl_b200.nn resolves lazily (NodeEmbedding, DeepWalk, GNNExplainer, graph-transformer blocks, ...)
    if name.startswith("__"):
        raise AttributeError(name)
    from .. import __getattr__ as _parent_getattr
    return _parent_getattr(name)
from ...transforms.functional import knn_graph, radius_graph, segmented_knn_graph  # noqa: F401,E402  (reference nn/pytorch/factory.py)
from .factory import pairwise_squared_distance  # noqa: F401,E402
