"""Datapipe stage base: functional verbs + ``MiniBatchTransformer``.

Parity: /root/reference/python/dgl/graphbolt/minibatch_transformer.py and the functional-form registrations spread over
the reference's datapipe classes. The reference builds on torchdata datapipes; here a stage is a plain Python iterable
whose verbs (``sample_neighbor``, ``fetch_feature``, ``copy_to`` ...) construct the next stage."""
from __future__ import annotations

from .minibatch import MiniBatch

__all__ = ["MiniBatchTransformer"]


class _Stage:
    """Iterable with the functional verbs of the reference's datapipes."""

    def __init__(self, datapipe):
        self.datapipe = datapipe

    def __len__(self):
        return len(self.datapipe)

    def sample_neighbor(self, graph, fanouts, **kw):
        from .impl.neighbor_sampler import NeighborSampler
        return NeighborSampler(self, graph, fanouts, **kw)

    def sample_layer_neighbor(self, graph, fanouts, **kw):
        from .impl.neighbor_sampler import LayerNeighborSampler
        return LayerNeighborSampler(self, graph, fanouts, **kw)

    def sample_in_subgraph(self, graph):
        from .impl.in_subgraph_sampler import InSubgraphSampler
        return InSubgraphSampler(self, graph)

    def sample_uniform_negative(self, graph, negative_ratio):
        from .impl.negative_sampler import UniformNegativeSampler
        return UniformNegativeSampler(self, graph, negative_ratio)

    def fetch_feature(self, feature_store, node_feature_keys=None, edge_feature_keys=None, **kw):
        from .feature_fetcher import FeatureFetcher
        return FeatureFetcher(self, feature_store, node_feature_keys, edge_feature_keys, **kw)

    def copy_to(self, device, non_blocking=False):
        from .dataloader import CopyTo
        return CopyTo(self, device, non_blocking)

    def transform(self, fn):
        return MiniBatchTransformer(self, fn)

    def exclude_seed_edges(self, include_reverse_edges=False, reverse_etypes_mapping=None, asynchronous=False):
        from functools import partial
        from .utils import exclude_seed_edges
        return MiniBatchTransformer(self, partial(exclude_seed_edges, include_reverse_edges=include_reverse_edges,
                                                  reverse_etypes_mapping=reverse_etypes_mapping))


class MiniBatchTransformer(_Stage):
    def __init__(self, datapipe, transformer=None):
        super().__init__(datapipe)
        self.transformer = transformer or (lambda mb: mb)

    def __iter__(self):
        for mb in self.datapipe:
            out = self.transformer(mb)
            assert isinstance(out, MiniBatch) or out is None or True
            yield out


def __getattr__(name):
    # remaining names of the reference module of this path (markers, helpers) are defined in graphbolt/compat.py
    if name.startswith("__"):
        raise AttributeError(name)
    from . import compat
    try:
        return getattr(compat, name)
    except AttributeError:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}") from None
