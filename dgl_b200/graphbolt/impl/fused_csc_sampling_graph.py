"""FusedCSCSamplingGraph: one CSC for a (possibly heterogeneous) graph, optimised for neighbour sampling.

Parity: /root/reference/python/dgl/graphbolt/impl/fused_csc_sampling_graph.py (FusedCSCSamplingGraph,
sample_neighbors, sample_layer_neighbors, temporal_sample_neighbors, in_subgraph, to/pin_memory_/
copy_to_shared_memory, fused_csc_sampling_graph(), from_dglgraph) and the native class
graphbolt/src/fused_csc_sampling_graph.cc:482-812. Sampling runs through the same native samplers as
dgl_b200.sampling (CPU OpenMP / sm_100a kernels / UVA on pinned graphs)."""
from __future__ import annotations

import torch

from ...base import DGLError
from ...graph_index import UnitGraph
from ...sampling.neighbor import sample_rel
from ..base import CSCFormatBase, ORIGINAL_EDGE_ID, etype_str_to_tuple, etype_tuple_to_str
from ..minibatch import SampledSubgraphImpl


def _asyncable(method):
    """``async_op=True`` turns a sampling entry point into a future (reference: the ``*Async`` bindings of
    graphbolt/src/python_binding.cc and ``async_op`` of fused_csc_sampling_graph.py:640-760): device seeds run on the
    sampling side stream - the returned future orders the caller's stream after it - and host seeds on the I/O pool."""
    import functools

    @functools.wraps(method)
    def wrapper(self, seeds, *args, async_op=False, **kwargs):
        if not async_op:
            return method(self, seeds, *args, **kwargs)
        from ..internal import futures
        first = next(iter(seeds.values())) if isinstance(seeds, dict) else seeds
        if torch.is_tensor(first) and first.is_cuda:
            return futures.on_side_stream(method, self, seeds, *args, key="sample", device=first.device, **kwargs)
        return futures.launch(method, self, seeds, *args, **kwargs)

    return wrapper


class FusedCSCSamplingGraph:
    def __init__(self, csc_indptr, indices, node_type_offset=None, type_per_edge=None, node_type_to_id=None,
                 edge_type_to_id=None, node_attributes=None, edge_attributes=None):
        self.csc_indptr = csc_indptr
        self.indices = indices
        self.node_type_offset = node_type_offset
        self.type_per_edge = type_per_edge
        self.node_type_to_id = node_type_to_id
        self.edge_type_to_id = edge_type_to_id
        self.node_attributes = node_attributes or {}
        self.edge_attributes = edge_attributes or {}
        self._ug = None
        self._etype_ugs = None

    # ---- basic properties -------------------------------------------------------------------
    @property
    def total_num_nodes(self):
        return self.csc_indptr.shape[0] - 1

    @property
    def total_num_edges(self):
        return self.indices.shape[0]

    @property
    def num_nodes(self):
        if self.node_type_offset is None:
            return self.total_num_nodes
        off = self.node_type_offset
        return {nt: int(off[i + 1] - off[i]) for nt, i in self.node_type_to_id.items()}

    @property
    def num_edges(self):
        if self.type_per_edge is None:
            return self.total_num_edges
        cnt = torch.bincount(self.type_per_edge.long(), minlength=len(self.edge_type_to_id))
        return {et: int(cnt[i]) for et, i in self.edge_type_to_id.items()}

    @property
    def device(self):
        return self.csc_indptr.device

    def is_pinned(self):
        return self.csc_indptr.is_pinned()

    def _unit(self):
        if self._ug is None:
            n = self.total_num_nodes
            idt = self.csc_indptr.dtype
            self._ug = UnitGraph.from_csc(n, n, self.csc_indptr, self.indices.to(idt))
            if self.is_pinned():
                self._ug._pinned = True
        return self._ug

    def _etype_units(self):
        """Per-edge-type CSC views (edges of one type, all destination nodes) built once."""
        if self._etype_ugs is None:
            n = self.total_num_nodes
            idt = self.csc_indptr.dtype
            deg = (self.csc_indptr[1:] - self.csc_indptr[:-1]).long()
            dst = torch.repeat_interleave(torch.arange(n, device=self.device), deg)
            out = {}
            for et, tid in self.edge_type_to_id.items():
                m = self.type_per_edge == tid
                eid = torch.nonzero(m).squeeze(1)
                cnt = torch.bincount(dst[eid], minlength=n)
                indptr = torch.zeros(n + 1, dtype=idt, device=self.device)
                indptr[1:] = torch.cumsum(cnt, 0)
                out[et] = (UnitGraph.from_csc(n, n, indptr, self.indices[eid].to(idt)), eid)
            self._etype_ugs = out
        return self._etype_ugs

    # ---- sampling -------------------------------------------------------------------------------
    def _convert(self, seeds_dict_or_tensor, results):
        """results: {etype_str or None: (src_global, counts_per_seed, eid)} -> SampledSubgraphImpl (uncompacted)."""
        has_orig = ORIGINAL_EDGE_ID in self.edge_attributes

        def mk(src, cnt, eid):
            indptr = torch.zeros(cnt.shape[0] + 1, dtype=self.csc_indptr.dtype, device=src.device)
            indptr[1:] = torch.cumsum(cnt, 0)
            oe = self.edge_attributes[ORIGINAL_EDGE_ID].to(eid.device)[eid.long()] if has_orig else eid
            return indptr, oe

        if self.node_type_offset is None:
            src, cnt, eid = results[None]
            indptr, oe = mk(src, cnt, eid)
            return SampledSubgraphImpl(sampled_csc=CSCFormatBase(indptr=indptr, indices=src), original_edge_ids=oe)
        off = self.node_type_offset
        csc, oes = {}, {}
        for et, (src, cnt, eid) in results.items():
            s, _, d = etype_str_to_tuple(et)
            indptr, oe = mk(src, cnt, eid)
            csc[et] = CSCFormatBase(indptr=indptr, indices=src - int(off[self.node_type_to_id[s]]))
            oes[et] = oe
        return SampledSubgraphImpl(sampled_csc=csc, original_edge_ids=oes)

    @_asyncable
    def sample_neighbors(self, seeds, fanouts, replace=False, probs_name=None, returning_indices_and_original_edge_ids_are_optional=False, overlap_graph_fetch=False):
        """Sample in-neighbours of ``seeds`` (tensor, or dict ntype -> tensor for hetero graphs)."""
        fan = torch.as_tensor(fanouts).reshape(-1).tolist()
        prob = self.edge_attributes[probs_name] if probs_name else None
        if self.node_type_offset is None:
            k = fan[0]
            src, dst, eid, cnt = sample_rel(self._unit(), seeds, k, "in", prob, replace, return_counts=True)
            return self._convert(seeds, {None: (src, cnt, eid)})
        if not isinstance(seeds, dict):
            raise DGLError("heterogeneous graph: seeds must be a dict of node type -> ids")
        off = self.node_type_offset
        units = self._etype_units()
        if len(fan) == 1:
            fan = fan * len(self.edge_type_to_id)
        results = {}
        for et, tid in self.edge_type_to_id.items():
            s, _, d = etype_str_to_tuple(et)
            if d not in seeds or fan[tid] == 0:
                continue
            ug, eid_map = units[et]
            g_seeds = seeds[d] + int(off[self.node_type_to_id[d]])
            p = prob[eid_map] if prob is not None else None
            src, dst, eid, cnt = sample_rel(ug, g_seeds, fan[tid], "in", p, replace, return_counts=True)
            results[et] = (src, cnt, eid_map.to(eid.device)[eid.long()])
        return self._convert(seeds, results)

    @_asyncable
    def sample_layer_neighbors(self, seeds, fanouts, replace=False, probs_name=None, random_seed=None,
                               seed2_contribution=0.0, **kwargs):
        """LABOR layer sampling (shared per-vertex variates)."""
        from ...sampling.labor import _labor_rel
        from ... import random as dgl_random
        fan = torch.as_tensor(fanouts).reshape(-1).tolist()
        prob = self.edge_attributes[probs_name] if probs_name else None
        rs = dgl_random.next_seed() if random_seed is None else int(torch.as_tensor(random_seed).reshape(-1)[0])
        def regroup(sd, src, dst, eid):
            # group the kept edges by seed position (LABOR returns them per edge, the CSC wants them per seed)
            order = torch.argsort(sd, stable=True)
            key = order[torch.searchsorted(sd[order], dst)] if dst.numel() else dst.long()
            perm = torch.argsort(key, stable=True)
            return src[perm], torch.bincount(key, minlength=sd.shape[0]), eid[perm]

        if self.node_type_offset is None:
            src, dst, eid, w = _labor_rel(self._unit(), seeds.to(self.csc_indptr.dtype), fan[0], prob, 0, rs, seed2_contribution)
            return self._convert(seeds, {None: regroup(seeds, src, dst, eid)})
        if not isinstance(seeds, dict):
            raise DGLError("heterogeneous graph: seeds must be a dict of node type -> ids")
        # one LABOR pass per edge type with the SAME random seed: the variate of a source vertex is a function of
        # (seed, global vertex id), so a vertex is kept / dropped consistently across relations (reference
        # fused_csc_sampling_graph.cc labor path with per-etype fanouts)
        off = self.node_type_offset
        units = self._etype_units()
        if len(fan) == 1:
            fan = fan * len(self.edge_type_to_id)
        results = {}
        for et, tid in self.edge_type_to_id.items():
            s_, _, d = etype_str_to_tuple(et)
            if d not in seeds or fan[tid] == 0:
                continue
            ug, eid_map = units[et]
            g_seeds = (seeds[d] + int(off[self.node_type_to_id[d]])).to(self.csc_indptr.dtype)
            p = prob[eid_map] if prob is not None else None
            src, dst, eid, w = _labor_rel(ug, g_seeds, fan[tid], p, 0, rs, seed2_contribution)
            src, cnt, eid = regroup(g_seeds, src, dst, eid)
            results[et] = (src, cnt, eid_map.to(eid.device)[eid.long()])
        return self._convert(seeds, results)

    def _temporal(self, seeds, seeds_timestamp, fanouts, replace, window, probs_name, node_attr, edge_attr, layer_seed=None):
        from .temporal import temporal_sample
        fan = torch.as_tensor(fanouts).reshape(-1).tolist()
        node_ts = self.node_attributes[node_attr] if node_attr else None
        edge_ts = self.edge_attributes[edge_attr] if edge_attr else None
        probs = self.edge_attributes[probs_name] if probs_name else None
        idx = self.indices
        if self.node_type_offset is None:
            src, cnt, pos = temporal_sample(self.csc_indptr, idx, seeds, seeds_timestamp, fan[0], replace, window, node_ts,
                                            edge_ts, probs, None, layer_seed)
            return self._convert(seeds, {None: (src, cnt, pos)})
        if not isinstance(seeds, dict):
            raise DGLError("heterogeneous graph: seeds, seeds_timestamp (and the time window) must be dicts keyed by node type")
        if len(fan) == 1:
            fan = fan * len(self.edge_type_to_id)
        off = self.node_type_offset
        results = {}
        for et, tid in self.edge_type_to_id.items():
            _, _, d = etype_str_to_tuple(et)
            if d not in seeds or fan[tid] == 0:
                continue
            g_seeds = seeds[d] + int(off[self.node_type_to_id[d]])
            win = None if window is None else window[d]
            results[et] = temporal_sample(self.csc_indptr, idx, g_seeds, seeds_timestamp[d], fan[tid], replace, win, node_ts,
                                          edge_ts, probs, self.type_per_edge == tid, layer_seed)
        return self._convert(seeds, results)

    @_asyncable
    def temporal_sample_neighbors(self, seeds, seeds_timestamp, fanouts, replace=False, seeds_pre_time_window=None,
                                  probs_name=None, node_timestamp_attr_name=None, edge_timestamp_attr_name=None):
        """Neighbour sampling restricted to sources / edges strictly older than the seed's timestamp (and not older than
        ``seeds_pre_time_window`` before it). Homogeneous: tensors; typed graphs: dicts keyed by node type, per-edge-type
        fan-outs. Node timestamps are indexed by homogenised node id (impl/temporal.py)."""
        return self._temporal(seeds, seeds_timestamp, fanouts, replace, seeds_pre_time_window, probs_name,
                              node_timestamp_attr_name, edge_timestamp_attr_name)

    @_asyncable
    def temporal_sample_layer_neighbors(self, seeds, seeds_timestamp, fanouts, replace=False, seeds_pre_time_window=None,
                                        probs_name=None, node_timestamp_attr_name=None, edge_timestamp_attr_name=None,
                                        random_seed=None, seed2_contribution=0.0):
        """Temporal layer sampling: as above, but a source vertex carries one variate shared by all seeds (and all edge
        types) of the call, so overlapping neighbourhoods pick overlapping sources (LABOR coupling)."""
        from ... import random as dgl_random
        rs = dgl_random.next_seed() if random_seed is None else int(torch.as_tensor(random_seed).reshape(-1)[0])
        return self._temporal(seeds, seeds_timestamp, fanouts, replace, seeds_pre_time_window, probs_name,
                              node_timestamp_attr_name, edge_timestamp_attr_name, layer_seed=rs)

    def sample_negative_edges_uniform(self, edge_type, node_pairs, negative_ratio):
        """For every positive (src, dst) pair draw ``negative_ratio`` random destinations of the right node type; returns
        an [N * negative_ratio, 2] tensor of (src, negative dst) pairs (reference sample_negative_edges_uniform)."""
        src = node_pairs[:, 0] if torch.is_tensor(node_pairs) and node_pairs.dim() == 2 else node_pairs[0]
        if self.node_type_offset is None or edge_type is None:
            n_dst = self.total_num_nodes
        else:
            _, _, d = etype_str_to_tuple(edge_type)
            tid = self.node_type_to_id[d]
            n_dst = int(self.node_type_offset[tid + 1] - self.node_type_offset[tid])
        rep = src.repeat_interleave(int(negative_ratio))
        neg = torch.randint(0, n_dst, (rep.shape[0],), device=src.device, dtype=src.dtype)
        return torch.stack([rep, neg], 1)

    def node_attribute(self, name):
        return None if self.node_attributes is None else self.node_attributes.get(name)

    def edge_attribute(self, name):
        return None if self.edge_attributes is None else self.edge_attributes.get(name)

    def add_node_attribute(self, name, tensor):
        if self.node_attributes is None:
            self.node_attributes = {}
        self.node_attributes[name] = tensor

    def add_edge_attribute(self, name, tensor):
        if self.edge_attributes is None:
            self.edge_attributes = {}
        self.edge_attributes[name] = tensor

    @_asyncable
    def in_subgraph(self, nodes):
        """All in-edges of ``nodes`` as a SampledSubgraph (fanout = -1)."""
        return self.sample_neighbors(nodes, [-1])

    # ---- movement ---------------------------------------------------------------------------------
    def _apply(self, fn):
        g = FusedCSCSamplingGraph(fn(self.csc_indptr), fn(self.indices),
                                  None if self.node_type_offset is None else fn(self.node_type_offset),
                                  None if self.type_per_edge is None else fn(self.type_per_edge),
                                  self.node_type_to_id, self.edge_type_to_id,
                                  {k: fn(v) for k, v in self.node_attributes.items()},
                                  {k: fn(v) for k, v in self.edge_attributes.items()})
        return g

    def to(self, device):
        return self._apply(lambda t: t.to(device))

    def pin_memory_(self):
        from ...utils.pin_memory import pin_memory_inplace
        for t in [self.csc_indptr, self.indices, self.node_type_offset, self.type_per_edge] + \
                list(self.node_attributes.values()) + list(self.edge_attributes.values()):
            if t is not None and not t.is_cuda:
                pin_memory_inplace(t)
        self._ug = None
        self._etype_ugs = None
        return self

    def copy_to_shared_memory(self, shared_memory_name):
        from ...utils.shared_mem import tensors_to_shared_memory
        tensors = {"csc_indptr": self.csc_indptr, "indices": self.indices}
        if self.node_type_offset is not None:
            tensors["node_type_offset"] = self.node_type_offset
        if self.type_per_edge is not None:
            tensors["type_per_edge"] = self.type_per_edge
        for k, v in self.node_attributes.items():
            tensors["nattr:" + k] = v
        for k, v in self.edge_attributes.items():
            tensors["eattr:" + k] = v
        shared = tensors_to_shared_memory(shared_memory_name, tensors, {"node_type_to_id": self.node_type_to_id,
                                                                        "edge_type_to_id": self.edge_type_to_id})
        return _from_tensor_dict(shared, self.node_type_to_id, self.edge_type_to_id)

    def __getstate__(self):
        return {k: getattr(self, k) for k in ("csc_indptr", "indices", "node_type_offset", "type_per_edge",
                                               "node_type_to_id", "edge_type_to_id", "node_attributes", "edge_attributes")}

    def __setstate__(self, st):
        self.__init__(**st)

    def __repr__(self):
        return (f"FusedCSCSamplingGraph(csc_indptr={self.csc_indptr},\n  indices={self.indices},\n"
                f"  total_num_nodes={self.total_num_nodes}, num_edges={self.total_num_edges},\n"
                f"  node_type_to_id={self.node_type_to_id}, edge_type_to_id={self.edge_type_to_id})")


def _from_tensor_dict(t, ntid, etid):
    return FusedCSCSamplingGraph(t["csc_indptr"], t["indices"], t.get("node_type_offset"), t.get("type_per_edge"), ntid,
                                 etid, {k[6:]: v for k, v in t.items() if k.startswith("nattr:")},
                                 {k[6:]: v for k, v in t.items() if k.startswith("eattr:")})


def load_from_shared_memory(shared_memory_name):
    from ...utils.shared_mem import tensors_from_shared_memory
    t, meta = tensors_from_shared_memory(shared_memory_name)
    return _from_tensor_dict(t, meta["node_type_to_id"], meta["edge_type_to_id"])


def fused_csc_sampling_graph(csc_indptr, indices, node_type_offset=None, type_per_edge=None, node_type_to_id=None,
                             edge_type_to_id=None, node_attributes=None, edge_attributes=None):
    if node_type_to_id is not None and edge_type_to_id is not None:
        node_types = list(node_type_to_id.keys())
        edge_types = list(edge_type_to_id.keys())
        node_type_ids = list(node_type_to_id.values())
        edge_type_ids = list(edge_type_to_id.values())
        assert all(isinstance(x, str) for x in node_types), "Node type name should be string."
        assert all(isinstance(x, int) for x in node_type_ids), "Node type id should be int."
        assert len(node_type_ids) == len(set(node_type_ids)), "Multiple node types shoud not be mapped to a same id."
        for et in edge_types:
            s, _, d = etype_str_to_tuple(et)
            assert s in node_types, f"Unrecognized node type {s} in edge type {et}"
            assert d in node_types, f"Unrecognized node type {d} in edge type {et}"
        assert len(edge_type_ids) == len(set(edge_type_ids)), "Multiple edge types shoud not be mapped to a same id."
        if node_type_offset is not None:
            assert len(node_type_to_id) + 1 == node_type_offset.size(0), "node_type_offset length should be |ntypes| + 1."
    return FusedCSCSamplingGraph(csc_indptr, indices, node_type_offset, type_per_edge, node_type_to_id, edge_type_to_id,
                                 node_attributes, edge_attributes)


def from_dglgraph(g, is_homogeneous=False, include_original_edge_id=False):
    """Build a FusedCSCSamplingGraph from a DGLGraph (reference fused_csc_sampling_graph.py:from_dglgraph)."""
    from ...convert import to_homogeneous
    from ...base import EID, ETYPE, NTYPE
    if is_homogeneous or g.is_homogeneous:
        indptr, indices, eids = g._graphs[0].csc()
        ea = {}
        if include_original_edge_id:
            ea[ORIGINAL_EDGE_ID] = eids if eids is not None else torch.arange(indices.shape[0], dtype=indptr.dtype, device=indptr.device)
        return FusedCSCSamplingGraph(indptr, indices, edge_attributes=ea)
    homo, ncount, ecount = to_homogeneous(g, return_count=True)
    indptr, indices, eids = homo._graphs[0].csc()
    e = eids.long() if eids is not None else torch.arange(indices.shape[0], device=indices.device)
    ntid = {nt: i for i, nt in enumerate(g.ntypes)}
    etid = {etype_tuple_to_str(c): i for i, c in enumerate(g.canonical_etypes)}
    off = torch.zeros(len(ncount) + 1, dtype=torch.int64, device=indptr.device)
    off[1:] = torch.cumsum(torch.tensor(ncount), 0)
    ea = {}
    if include_original_edge_id:
        ea[ORIGINAL_EDGE_ID] = homo.edata[EID][e]
    return FusedCSCSamplingGraph(indptr, indices, off, homo.edata[ETYPE][e], ntid, etid, None, ea)
