"""``FeatureFetcher``: reads the node features of ``input_nodes`` and the edge features of every sampled layer from a
feature store (device tables, pinned host tables through UVA, GPU / CPU caches, disk).

Parity: /root/reference/python/dgl/graphbolt/feature_fetcher.py:40-190."""
from __future__ import annotations

from .minibatch_transformer import MiniBatchTransformer

__all__ = ["FeatureFetcher"]


class FeatureFetcher(MiniBatchTransformer):
    """Read node features of ``input_nodes`` and edge features of every sampled layer."""

    def __init__(self, datapipe, feature_store, node_feature_keys=None, edge_feature_keys=None, overlap_fetch=True,
                 cooperative=False):
        super().__init__(datapipe, self._read)
        self.feature_store = feature_store
        self.node_feature_keys = node_feature_keys
        self.edge_feature_keys = edge_feature_keys

    def _read(self, data):
        node_features, edge_features = {}, []
        input_nodes = data.input_nodes
        if self.node_feature_keys and input_nodes is not None:
            if isinstance(self.node_feature_keys, dict):
                for type_name, names in self.node_feature_keys.items():
                    nodes = input_nodes[type_name] if isinstance(input_nodes, dict) else input_nodes
                    if nodes is None:
                        continue
                    for name in names:
                        node_features[(type_name, name)] = self.feature_store.read("node", type_name, name, nodes)
            else:
                for name in self.node_feature_keys:
                    node_features[name] = self.feature_store.read("node", None, name, input_nodes)
        if self.edge_feature_keys and data.sampled_subgraphs:
            for sg in data.sampled_subgraphs:
                ef = {}
                eids = sg.original_edge_ids
                if eids is not None:
                    if isinstance(self.edge_feature_keys, dict):
                        for type_name, names in self.edge_feature_keys.items():
                            e = eids.get(type_name) if isinstance(eids, dict) else eids
                            if e is None:
                                continue
                            for name in names:
                                ef[(type_name, name)] = self.feature_store.read("edge", type_name, name, e)
                    else:
                        for name in self.edge_feature_keys:
                            ef[name] = self.feature_store.read("edge", None, name, eids)
                edge_features.append(ef)
        data.set_node_features(node_features)
        data.set_edge_features(edge_features)
        return data


def __getattr__(name):
    # remaining names of the reference module of this path (markers, helpers) are defined in graphbolt/compat.py
    if name.startswith("__"):
        raise AttributeError(name)
    from . import compat
    try:
        return getattr(compat, name)
    except AttributeError:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}") from None
