"""Sampler base classes and edge-prediction wrapper (parity: /root/reference/python/dgl/dataloading/base.py)."""
from __future__ import annotations

from collections.abc import Mapping

import torch

from ..base import EID, NID, DGLError
from ..frame import LazyFeature
from ..transforms import compact_graphs


def _set_lazy_features(x, xdata, feature_names):
    if feature_names is None:
        return
    if not isinstance(feature_names, Mapping):
        xdata.update({k: LazyFeature(k) for k in feature_names})
    else:
        for type_, names in feature_names.items():
            x[type_].data.update({k: LazyFeature(k) for k in names})


def set_node_lazy_features(g, feature_names):
    return _set_lazy_features(g.nodes, g.ndata, feature_names)


def set_edge_lazy_features(g, feature_names):
    return _set_lazy_features(g.edges, g.edata, feature_names)


def set_src_lazy_features(g, feature_names):
    return _set_lazy_features(g.srcnodes, g.srcdata, feature_names)


def set_dst_lazy_features(g, feature_names):
    return _set_lazy_features(g.dstnodes, g.dstdata, feature_names)


def _fetch_frame(frame, parent_frame, ids, device):
    for k in list(frame.keys()):
        col = frame._columns[k]
        if isinstance(col.storage, LazyFeature):
            name = col.storage.name or k
            src = parent_frame[name]
            idx = ids
            if src.is_cuda:
                val = src[idx.to(src.device).long()]
            elif idx.is_cuda and src.is_pinned():
                from ..ops.index import index_select
                val = index_select(src, idx)
            else:
                val = src[idx.cpu().long()]
            frame[k] = val.to(device, non_blocking=True)


def assign_lazy_features(result, parent_graph=None, device=None):
    """Resolve LazyFeature placeholders on blocks/subgraphs against the parent graph (the DataLoader's
    prefetch step; reference dataloader.py:_prefetch :487)."""
    def visit(x):
        if hasattr(x, "_node_frames"):
            g = x
            dev = device or g.device
            for i, fr in enumerate(g._node_frames):
                if NID in fr and parent_graph is not None:
                    nt = g._ntypes[i]
                    pf = parent_graph._node_frames[parent_graph.get_ntype_id(nt)]
                    _fetch_frame(fr, pf, fr[NID], dev)
            for i, fr in enumerate(g._edge_frames):
                if EID in fr and parent_graph is not None:
                    pf = parent_graph._edge_frames[parent_graph.get_etype_id(g.canonical_etypes[i])]
                    _fetch_frame(fr, pf, fr[EID], dev)
        elif isinstance(x, (list, tuple)):
            for y in x:
                visit(y)
        elif isinstance(x, dict):
            for y in x.values():
                visit(y)
    visit(result)
    return result


class Sampler:
    """Base sampler: ``sample(g, indices)`` returns any structure of tensors/graphs."""

    def sample(self, g, indices):
        raise NotImplementedError


class BlockSampler(Sampler):
    """Samplers producing a list of MFGs; subclasses implement ``sample_blocks``."""

    def __init__(self, prefetch_node_feats=None, prefetch_labels=None, prefetch_edge_feats=None, output_device=None):
        super().__init__()
        self.prefetch_node_feats = prefetch_node_feats or []
        self.prefetch_labels = prefetch_labels or []
        self.prefetch_edge_feats = prefetch_edge_feats or []
        self.output_device = output_device

    def sample_blocks(self, g, seed_nodes, exclude_eids=None):
        raise NotImplementedError

    def assign_lazy_features(self, result):
        input_nodes, output_nodes, blocks = result
        set_src_lazy_features(blocks[0], self.prefetch_node_feats)
        set_dst_lazy_features(blocks[-1], self.prefetch_labels)
        for block in blocks:
            set_edge_lazy_features(block, self.prefetch_edge_feats)
        return input_nodes, output_nodes, blocks

    def sample(self, g, seed_nodes, exclude_eids=None):
        result = self.sample_blocks(g, seed_nodes, exclude_eids=exclude_eids)
        return self.assign_lazy_features(result)


def _find_exclude_eids(g, exclude, seed_edges, reverse_eids=None, reverse_etypes=None, output_device=None):
    if exclude is None:
        return None
    if callable(exclude):
        return exclude(seed_edges)
    if exclude == "self":
        return seed_edges
    if exclude == "reverse_id":
        if isinstance(seed_edges, Mapping):
            return {k: torch.cat([v, reverse_eids[k][v.long().to(reverse_eids[k].device)].to(v.device)]) for k, v in seed_edges.items()}
        return torch.cat([seed_edges, reverse_eids[seed_edges.long().to(reverse_eids.device)].to(seed_edges.device)])
    if exclude == "reverse_types":
        out = {g.to_canonical_etype(k): v for k, v in seed_edges.items()}
        for k, v in list(out.items()):
            rk = g.to_canonical_etype(reverse_etypes[k] if k in reverse_etypes else reverse_etypes[k[1]])
            out[rk] = torch.cat([out[rk], v]) if rk in out else v
        return out
    raise ValueError("unsupported mode {}".format(exclude))


class EdgePredictionSampler(Sampler):
    """Wrap a node sampler for edge classification / link prediction (reference base.py:EdgePredictionSampler)."""

    def __init__(self, sampler, exclude=None, reverse_eids=None, reverse_etypes=None, negative_sampler=None,
                 prefetch_labels=None):
        super().__init__()
        if exclude == "reverse_id" and reverse_eids is None:
            raise ValueError("reverse_eids must be provided when exclude is 'reverse_id'")
        if exclude == "reverse_types" and reverse_etypes is None:
            raise ValueError("reverse_etypes must be provided when exclude is 'reverse_types'")
        self.reverse_eids = reverse_eids
        self.reverse_etypes = reverse_etypes
        self.exclude = exclude
        self.sampler = sampler
        self.negative_sampler = negative_sampler
        self.prefetch_labels = prefetch_labels or []
        self.output_device = sampler.output_device if hasattr(sampler, "output_device") else None

    def _build_neg_graph(self, g, seed_edges):
        from ..convert import heterograph
        neg_srcdst = self.negative_sampler(g, seed_edges)
        if not isinstance(neg_srcdst, Mapping):
            neg_srcdst = {g.canonical_etypes[0]: neg_srcdst}
        dtype = next(iter(neg_srcdst.values()))[0].dtype
        dev = next(iter(neg_srcdst.values()))[0].device
        neg_edges = {etype: neg_srcdst.get(etype, (torch.empty(0, dtype=dtype, device=dev), torch.empty(0, dtype=dtype, device=dev)))
                     for etype in g.canonical_etypes}
        neg_pair_graph = heterograph(neg_edges, {ntype: g.num_nodes(ntype) for ntype in g.ntypes})
        return neg_pair_graph, neg_srcdst

    def assign_lazy_features(self, result):
        pair_graph = result[1]
        set_edge_lazy_features(pair_graph, self.prefetch_labels)
        return result

    def sample(self, g, seed_edges):
        if isinstance(seed_edges, Mapping):
            seed_edges = {g.to_canonical_etype(k): v for k, v in seed_edges.items()}
        exclude = self.exclude
        pair_graph = g.edge_subgraph(seed_edges, relabel_nodes=False, output_device=self.output_device)
        eids = pair_graph.edata[EID]
        if self.negative_sampler is not None:
            neg_graph, _ = self._build_neg_graph(g, seed_edges)
            if neg_graph.device != pair_graph.device:
                neg_graph = neg_graph.to(pair_graph.device)
            pair_graph, neg_graph = compact_graphs([pair_graph, neg_graph])
        else:
            pair_graph = compact_graphs(pair_graph)
        pair_graph.edata[EID] = eids
        seed_nodes = pair_graph.ndata[NID]
        exclude_eids = _find_exclude_eids(g, exclude, seed_edges, reverse_eids=self.reverse_eids,
                                          reverse_etypes=self.reverse_etypes, output_device=self.output_device)
        input_nodes, _, blocks = self.sampler.sample(g, seed_nodes, exclude_eids)
        if self.negative_sampler is None:
            return self.assign_lazy_features((input_nodes, pair_graph, blocks))
        return self.assign_lazy_features((input_nodes, pair_graph, neg_graph, blocks))


def as_edge_prediction_sampler(sampler, exclude=None, reverse_eids=None, reverse_etypes=None, negative_sampler=None,
                               prefetch_labels=None):
    return EdgePredictionSampler(sampler, exclude=exclude, reverse_eids=reverse_eids, reverse_etypes=reverse_etypes,
                                 negative_sampler=negative_sampler, prefetch_labels=prefetch_labels)
