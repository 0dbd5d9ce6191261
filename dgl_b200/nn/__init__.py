"""Neural-network modules (``dgl.nn``, PyTorch only).
Parity: /root/reference/python/dgl/nn/pytorch/__init__.py."""
from .conv import *  # noqa: F401,F403
from . import functional  # noqa: F401
from .glob import *  # noqa: F401,F403
from .linear import HeteroEmbedding, HeteroLinear, TypedLinear  # noqa: F401
from .link import EdgePredictor, TransE, TransR  # noqa: F401
from .utils import (Identity, JumpingKnowledge, KNNGraph, LabelPropagation, RadiusGraph, SegmentedKNNGraph,  # noqa: F401
                    Sequential, WeightBasis)


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    import importlib
    for modname in ("sparse_emb", "network_emb", "gt", "explain"):
        mod = importlib.import_module(f"{__name__}.{modname}")
        if hasattr(mod, name):
            return getattr(mod, name)
    raise AttributeError(f"module '{__name__}' has no attribute '{name}'")
