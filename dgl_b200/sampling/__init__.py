"""Sampling operators (parity: /root/reference/python/dgl/sampling/__init__.py)."""
from .neighbor import (sample_etype_neighbors, sample_neighbors, sample_neighbors_biased, sample_neighbors_fused,  # noqa: F401
                       select_topk)
from .labor import sample_labors  # noqa: F401
from .negative import global_uniform_negative_sampling  # noqa: F401
from .randomwalks import node2vec_random_walk, pack_traces, random_walk  # noqa: F401
from .pinsage import PinSAGESampler, RandomWalkNeighborSampler  # noqa: F401

from .utils import EidExcluder  # noqa: F401,E402
