"""Node/edge views and data views (``g.nodes``, ``g.edges``, ``g.ndata`` ...).

Parity: /root/reference/python/dgl/view.py.
"""
from __future__ import annotations

from collections import namedtuple
from collections.abc import MutableMapping

import torch

from .base import ALL, DGLError, is_all

NodeSpace = namedtuple("NodeSpace", ["data"])
EdgeSpace = namedtuple("EdgeSpace", ["data"])


class HeteroNodeView:
    def __init__(self, graph, typeid_getter):
        self._graph = graph
        self._typeid_getter = typeid_getter

    def __getitem__(self, key):
        if isinstance(key, slice):
            if not is_all(key):
                raise DGLError('Currently only full slice ":" is supported')
            nodes, ntype = ALL, None
        elif isinstance(key, tuple):
            nodes, ntype = key
        elif key is None or isinstance(key, str):
            nodes, ntype = ALL, key
        else:
            nodes, ntype = key, None
        ntid = self._typeid_getter(ntype)
        return NodeSpace(data=HeteroNodeDataView(self._graph, ntype, ntid, nodes))

    def __call__(self, ntype=None):
        ntid = self._typeid_getter(ntype)
        return torch.arange(self._graph._num_nodes[ntid], dtype=self._graph.idtype, device=self._graph.device)


class HeteroNodeDataView(MutableMapping):
    def __init__(self, graph, ntype, ntid, nodes):
        self._graph = graph
        self._ntype = ntype
        self._ntid = ntid
        self._nodes = nodes

    def _frames(self):
        if isinstance(self._ntid, list):
            return [self._graph._node_frames[i] for i in self._ntid]
        return self._graph._node_frames[self._ntid]

    def __getitem__(self, key):
        if isinstance(self._ntid, list):
            ret = {}
            for i in self._ntid:
                fr = self._graph._node_frames[i]
                if key in fr:
                    ret[self._graph._ntypes[i]] = fr[key]
            return ret
        fr = self._graph._node_frames[self._ntid]
        if is_all(self._nodes):
            return fr[key]
        return fr[key][_as_long(self._nodes, fr[key].device)]

    def __setitem__(self, key, val):
        if isinstance(self._ntid, list):
            if not isinstance(val, dict):
                raise DGLError("Current HeteroNodeDataView has multiple node types, please pass the node type "
                               "and the corresponding data through a dict.")
            for ntype, data in val.items():
                self._graph._set_n_repr(self._graph.get_ntype_id(ntype), self._nodes, {key: data})
        else:
            if isinstance(val, dict):
                raise DGLError("The HeteroNodeDataView has only one node type; pass a tensor directly.")
            self._graph._set_n_repr(self._ntid, self._nodes, {key: val})

    def __delitem__(self, key):
        if isinstance(self._ntid, list):
            for i in self._ntid:
                self._graph._node_frames[i].pop(key, None)
        else:
            del self._graph._node_frames[self._ntid][key]

    def _as_dict(self):
        if isinstance(self._ntid, list):
            keys = set()
            for i in self._ntid:
                keys.update(self._graph._node_frames[i].keys())
            return {k: self[k] for k in keys}
        fr = self._graph._node_frames[self._ntid]
        return {k: self[k] for k in fr.keys()}

    def __len__(self):
        return len(self._as_dict())

    def __iter__(self):
        return iter(self._as_dict())

    def keys(self):
        return self._as_dict().keys()

    def values(self):
        return self._as_dict().values()

    def items(self):
        return self._as_dict().items()

    def __contains__(self, key):
        return key in self._as_dict()

    def __repr__(self):
        return repr(self._as_dict())


class HeteroEdgeView:
    def __init__(self, graph):
        self._graph = graph

    def __getitem__(self, key):
        if isinstance(key, slice):
            if not is_all(key):
                raise DGLError('Currently only full slice ":" is supported')
            edges, etype = ALL, None
        elif key is None:
            edges, etype = ALL, None
        elif isinstance(key, tuple):
            if len(key) == 3 and all(isinstance(k, str) for k in key):
                edges, etype = ALL, key
            else:
                edges, etype = key, None
        elif isinstance(key, str):
            edges, etype = ALL, key
        else:
            edges, etype = key, None
        return EdgeSpace(data=HeteroEdgeDataView(self._graph, etype, edges))

    def __call__(self, *args, **kwargs):
        return self._graph.all_edges(*args, **kwargs)


class HeteroEdgeDataView(MutableMapping):
    def __init__(self, graph, etype, edges):
        self._graph = graph
        self._etype = etype
        self._etid = list(range(len(graph.canonical_etypes))) if (etype is None and len(graph.canonical_etypes) > 1) \
            else graph.get_etype_id(etype)
        self._edges = edges

    def __getitem__(self, key):
        if isinstance(self._etid, list):
            ret = {}
            for i in self._etid:
                fr = self._graph._edge_frames[i]
                if key in fr:
                    ret[self._graph.canonical_etypes[i]] = fr[key]
            return ret
        fr = self._graph._edge_frames[self._etid]
        if is_all(self._edges):
            return fr[key]
        eid = self._graph._parse_edges(self._edges, self._etid)
        return fr[key][_as_long(eid, fr[key].device)]

    def __setitem__(self, key, val):
        if isinstance(self._etid, list):
            if not isinstance(val, dict):
                raise DGLError("Current HeteroEdgeDataView has multiple edge types, please pass the edge type "
                               "and the corresponding data through a dict.")
            for etype, data in val.items():
                self._graph._set_e_repr(self._graph.get_etype_id(etype), self._edges, {key: data})
        else:
            if isinstance(val, dict):
                raise DGLError("The HeteroEdgeDataView has only one edge type; pass a tensor directly.")
            self._graph._set_e_repr(self._etid, self._edges, {key: val})

    def __delitem__(self, key):
        if isinstance(self._etid, list):
            for i in self._etid:
                self._graph._edge_frames[i].pop(key, None)
        else:
            del self._graph._edge_frames[self._etid][key]

    def _as_dict(self):
        if isinstance(self._etid, list):
            keys = set()
            for i in self._etid:
                keys.update(self._graph._edge_frames[i].keys())
            return {k: self[k] for k in keys}
        fr = self._graph._edge_frames[self._etid]
        return {k: self[k] for k in fr.keys()}

    def __len__(self):
        return len(self._as_dict())

    def __iter__(self):
        return iter(self._as_dict())

    def keys(self):
        return self._as_dict().keys()

    def values(self):
        return self._as_dict().values()

    def items(self):
        return self._as_dict().items()

    def __contains__(self, key):
        return key in self._as_dict()

    def __repr__(self):
        return repr(self._as_dict())


def _as_long(idx, device):
    if not torch.is_tensor(idx):
        idx = torch.as_tensor(idx)
    return idx.long().to(device)
