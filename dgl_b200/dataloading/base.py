"""Sampler protocol, feature-prefetch markers and the node-sampler -> edge-sampler adaptor.

Public surface follows /root/reference/python/dgl/dataloading/base.py (``Sampler`` :147, ``BlockSampler`` :170,
``EdgePredictionSampler`` :305, ``as_edge_prediction_sampler`` :453, the ``set_*_lazy_features`` helpers :68-144). The
mechanics are organised around three small pieces of this repo's own:

* ``_mark(view_kind, graph, names)`` - one routine that plants ``LazyFeature`` placeholders on any of the four frame
  views of a (sub)graph; ``assign_lazy_features`` later resolves every placeholder in one pass per frame, choosing the
  gather path from where the parent column lives (device table / pinned host table read through UVA / pageable host);
* ``_EXCLUDERS`` - a table of seed-edge exclusion policies instead of an if-chain;
* ``EdgePredictionSampler`` - builds the positive pair graph, the optional negative pair graph and compacts both into one
  shared node id space before delegating neighbourhood sampling to the wrapped sampler.
"""
from __future__ import annotations

from collections.abc import Mapping

import torch

from ..base import EID, NID
from ..frame import LazyFeature
from ..transforms import compact_graphs

_VIEWS = {"node": ("nodes", "ndata"), "edge": ("edges", "edata"), "src": ("srcnodes", "srcdata"), "dst": ("dstnodes", "dstdata")}


def _mark(kind, g, names):
    """Plant LazyFeature placeholders: ``names`` is a list (single type) or ``{type: list}``."""
    if not names:
        return
    typed_view, flat_view = _VIEWS[kind]
    if isinstance(names, Mapping):
        for key, cols in names.items():
            getattr(g, typed_view)[key].data.update({c: LazyFeature(c) for c in cols})
    else:
        getattr(g, flat_view).update({c: LazyFeature(c) for c in names})


def set_node_lazy_features(g, feature_names):
    _mark("node", g, feature_names)


def set_edge_lazy_features(g, feature_names):
    _mark("edge", g, feature_names)


def set_src_lazy_features(g, feature_names):
    _mark("src", g, feature_names)


def set_dst_lazy_features(g, feature_names):
    _mark("dst", g, feature_names)


def _gather_rows(table, ids):
    """rows ``ids`` of a parent column, wherever it lives."""
    if table.is_cuda:
        return table[ids.to(table.device).long()]
    if ids.is_cuda and table.is_pinned():
        from ..ops.index import index_select            # zero-copy read of the pinned table from the GPU
        return index_select(table, ids)
    return table[ids.cpu().long()]


def _resolve_frame(frame, parent_frame, ids, device):
    pending = [(k, frame._columns[k].storage) for k in list(frame.keys()) if isinstance(frame._columns[k].storage, LazyFeature)]
    for key, marker in pending:
        frame[key] = _gather_rows(parent_frame[marker.name or key], ids).to(device, non_blocking=True)


def _graphs_in(obj):
    if hasattr(obj, "_node_frames"):
        yield obj
    elif isinstance(obj, Mapping):
        for v in obj.values():
            yield from _graphs_in(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _graphs_in(v)


def assign_lazy_features(result, parent_graph=None, device=None):
    """Replace every placeholder found on the graphs inside ``result`` by rows of ``parent_graph``'s columns (the prefetch
    step of the DataLoader; reference dataloader.py:487)."""
    if parent_graph is None:
        return result
    for g in _graphs_in(result):
        dev = device or g.device
        for nt, fr in zip(g._ntypes, g._node_frames):
            if NID in fr:
                _resolve_frame(fr, parent_graph._node_frames[parent_graph.get_ntype_id(nt)], fr[NID], dev)
        for rel, fr in zip(g.canonical_etypes, g._edge_frames):
            if EID in fr:
                _resolve_frame(fr, parent_graph._edge_frames[parent_graph.get_etype_id(rel)], fr[EID], dev)
    return result


class Sampler:
    """``sample(g, indices)`` -> any nesting of tensors and graphs."""

    def sample(self, g, indices):
        raise NotImplementedError


class BlockSampler(Sampler):
    """Samplers that emit ``(input_nodes, output_nodes, [MFG, ...])``; subclasses provide ``sample_blocks``."""

    def __init__(self, prefetch_node_feats=None, prefetch_labels=None, prefetch_edge_feats=None, output_device=None):
        super().__init__()
        self.prefetch_node_feats = prefetch_node_feats or []
        self.prefetch_labels = prefetch_labels or []
        self.prefetch_edge_feats = prefetch_edge_feats or []
        self.output_device = output_device

    def sample_blocks(self, g, seed_nodes, exclude_eids=None):
        raise NotImplementedError

    def assign_lazy_features(self, result):
        blocks = result[2]
        _mark("src", blocks[0], self.prefetch_node_feats)      # inputs of the first layer
        _mark("dst", blocks[-1], self.prefetch_labels)         # targets of the last layer
        for b in blocks:
            _mark("edge", b, self.prefetch_edge_feats)
        return result

    def sample(self, g, seed_nodes, exclude_eids=None):
        return self.assign_lazy_features(self.sample_blocks(g, seed_nodes, exclude_eids=exclude_eids))


# ------------------------------------------------------------------------------------------ seed-edge exclusion policies
def _with_reverse_ids(g, seeds, reverse_eids, reverse_etypes):
    def both(ids, table):
        return torch.cat([ids, table[ids.long().to(table.device)].to(ids.device)])
    if isinstance(seeds, Mapping):
        return {rel: both(ids, reverse_eids[rel]) for rel, ids in seeds.items()}
    return both(seeds, reverse_eids)


def _with_reverse_types(g, seeds, reverse_eids, reverse_etypes):
    out = {g.to_canonical_etype(rel): ids for rel, ids in seeds.items()}
    for rel, ids in list(out.items()):
        partner = reverse_etypes[rel] if rel in reverse_etypes else reverse_etypes[rel[1]]
        partner = g.to_canonical_etype(partner)
        out[partner] = torch.cat([out[partner], ids]) if partner in out else ids
    return out


_EXCLUDERS = {"self": lambda g, seeds, *_: seeds, "reverse_id": _with_reverse_ids, "reverse_types": _with_reverse_types}


def _find_exclude_eids(g, exclude, seed_edges, reverse_eids=None, reverse_etypes=None, output_device=None):
    if exclude is None:
        return None
    if callable(exclude):
        return exclude(seed_edges)
    try:
        policy = _EXCLUDERS[exclude]
    except (KeyError, TypeError):
        raise ValueError(f"unsupported mode {exclude}") from None
    return policy(g, seed_edges, reverse_eids, reverse_etypes)


class EdgePredictionSampler(Sampler):
    """Turns a node-wise sampler into an edge-wise one: a batch of seed edges becomes a pair graph (plus a negative pair
    graph when a negative sampler is given) over a compacted node set, whose nodes seed the wrapped sampler."""

    _NEEDS = {"reverse_id": "reverse_eids", "reverse_types": "reverse_etypes"}

    def __init__(self, sampler, exclude=None, reverse_eids=None, reverse_etypes=None, negative_sampler=None,
                 prefetch_labels=None):
        super().__init__()
        needed = self._NEEDS.get(exclude) if isinstance(exclude, str) else None
        if needed and {"reverse_eids": reverse_eids, "reverse_etypes": reverse_etypes}[needed] is None:
            raise ValueError(f"{needed} must be provided when exclude is '{exclude}'")
        self.sampler, self.exclude = sampler, exclude
        self.reverse_eids, self.reverse_etypes = reverse_eids, reverse_etypes
        self.negative_sampler = negative_sampler
        self.prefetch_labels = prefetch_labels or []
        self.output_device = getattr(sampler, "output_device", None)

    def _build_neg_graph(self, g, seed_edges):
        from ..convert import heterograph
        drawn = self.negative_sampler(g, seed_edges)
        if not isinstance(drawn, Mapping):
            drawn = {g.canonical_etypes[0]: drawn}
        proto = next(iter(drawn.values()))[0]
        nothing = proto.new_empty(0)
        per_rel = {rel: drawn.get(rel, (nothing, nothing)) for rel in g.canonical_etypes}
        return heterograph(per_rel, {nt: g.num_nodes(nt) for nt in g.ntypes}), drawn

    def assign_lazy_features(self, result):
        _mark("edge", result[1], self.prefetch_labels)
        return result

    def sample(self, g, seed_edges):
        if isinstance(seed_edges, Mapping):
            seed_edges = {g.to_canonical_etype(rel): ids for rel, ids in seed_edges.items()}
        positives = g.edge_subgraph(seed_edges, relabel_nodes=False, output_device=self.output_device)
        parent_eids = positives.edata[EID]
        graphs = [positives]
        if self.negative_sampler is not None:
            negatives, _ = self._build_neg_graph(g, seed_edges)
            graphs.append(negatives if negatives.device == positives.device else negatives.to(positives.device))
        graphs = compact_graphs(graphs)
        positives = graphs[0]
        positives.edata[EID] = parent_eids
        banned = _find_exclude_eids(g, self.exclude, seed_edges, self.reverse_eids, self.reverse_etypes, self.output_device)
        input_nodes, _, blocks = self.sampler.sample(g, positives.ndata[NID], banned)
        return self.assign_lazy_features((input_nodes, *graphs, blocks))


def as_edge_prediction_sampler(sampler, exclude=None, reverse_eids=None, reverse_etypes=None, negative_sampler=None,
                               prefetch_labels=None):
    """Edge classification / link prediction flavour of ``sampler`` (see ``EdgePredictionSampler``)."""
    return EdgePredictionSampler(sampler, exclude=exclude, reverse_eids=reverse_eids, reverse_etypes=reverse_etypes,
                                 negative_sampler=negative_sampler, prefetch_labels=prefetch_labels)


def __getattr__(name):
    # names the reference keeps in this module but that live elsewhere in this package
    if name.startswith("__"):
        raise AttributeError(name)
    import importlib
    for modname in ('dgl_b200.dataloading.tensorized',):
        mod = importlib.import_module(modname)
        if name in vars(mod):
            return vars(mod)[name]
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
