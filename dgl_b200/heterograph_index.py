"""Index-level view of a heterograph (parity: /root/reference/python/dgl/heterograph_index.py).

In the reference ``DGLGraph._graph`` is a handle on a C++ ``HeteroGraph`` and the Python class is a thin veneer over
it. Here the relation list lives directly in ``DGLGraph`` (``_graphs``: one :class:`~dgl_b200.graph_index.UnitGraph`
per canonical edge type), so this module is the reverse adapter: :class:`HeteroGraphIndex` presents that state
through the reference's id-based (ntype id / etype id) method names for code written against ``g._graph``, and the
module-level creators build graphs from relation lists the way ``create_heterograph_from_relations`` & co. do."""
from __future__ import annotations

import torch

from .base import DGLError
from .graph_index import UnitGraph

__all__ = ["HeteroGraphIndex", "HeteroSubgraphIndex", "create_metagraph_index", "create_unitgraph_from_coo",
           "create_unitgraph_from_csr", "create_heterograph_from_relations", "create_heterograph_from_shared_memory",
           "joint_union", "disjoint_union", "disjoint_partition", "slice_gidx", "FlattenedHeteroGraph", "HeteroPickleStates"]


class _MetaGraph:
    """Node types as vertices, one edge per canonical edge type (ids are positions in the type lists)."""

    def __init__(self, num_ntypes, pairs):
        self._n = num_ntypes
        self._pairs = [(int(s), int(d)) for s, d in pairs]

    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return len(self._pairs)

    number_of_edges = num_edges

    def find_edge(self, eid):
        return self._pairs[eid]

    def edges(self, order=None):
        s = torch.tensor([p[0] for p in self._pairs], dtype=torch.int64)
        d = torch.tensor([p[1] for p in self._pairs], dtype=torch.int64)
        return s, d, torch.arange(len(self._pairs))

    def edge_id(self, u, v):
        return [i for i, p in enumerate(self._pairs) if p == (u, v)]


class HeteroGraphIndex:
    """Id-based structural interface of a :class:`~dgl_b200.DGLGraph` (methods take ntype / etype *ids*)."""

    def __init__(self, graph):
        self._g = graph

    # ---- types -------------------------------------------------------------------------------
    @property
    def metagraph(self):
        g = self._g
        return _MetaGraph(len(g.ntypes), [(g.get_ntype_id_from_src(c[0]) if g.is_block else g.get_ntype_id(c[0]),
                                           g.get_ntype_id_from_dst(c[2]) + len(g.srctypes) if g.is_block
                                           else g.get_ntype_id(c[2])) for c in g.canonical_etypes])

    def is_metagraph_unibipartite(self):
        return bool(self._g.is_unibipartite)

    def number_of_ntypes(self):
        return len(self._g.ntypes)

    def number_of_etypes(self):
        return len(self._g.canonical_etypes)

    def _ntype(self, ntid):
        return self._g.ntypes[ntid]

    def _etype(self, etid):
        return self._g.canonical_etypes[etid]

    def get_relation_graph(self, etype):
        return HeteroGraphIndex(self._g[self._etype(etype)])

    def get_unitgraph(self, etype, ctx=None):
        ug = self._g._graphs[etype]
        return ug if ctx is None else ug.to(torch.device(ctx))

    # ---- dtype / placement -----------------------------------------------------------------------
    @property
    def dtype(self):
        return "int32" if self._g.idtype == torch.int32 else "int64"

    @property
    def ctx(self):
        return self._g.device

    def bits_needed(self, etype):
        ug = self._g._graphs[etype]
        big = max(ug.num_src, ug.num_dst, ug.num_edges()) >= 0x80000000 - 3
        return 64 if big else 32

    def asbits(self, bits):
        return HeteroGraphIndex(self._g.int() if bits == 32 else self._g.long())

    def copy_to(self, ctx):
        return HeteroGraphIndex(self._g.to(ctx))

    def pin_memory(self):
        g = self._g.clone()
        g.pin_memory_()
        return HeteroGraphIndex(g)

    def pin_memory_(self):
        self._g.pin_memory_()

    def unpin_memory_(self):
        self._g.unpin_memory_()

    def is_pinned(self):
        return self._g.is_pinned()

    def record_stream(self, stream):
        self._g.record_stream(stream)

    def shared_memory(self, name, ntypes=None, etypes=None, formats=("coo", "csr", "csc")):
        return HeteroGraphIndex(self._g.shared_memory(name, formats=formats))

    def formats(self, formats=None):
        if formats is None:
            return self._g.formats()
        return HeteroGraphIndex(self._g.formats(formats))

    def create_formats_(self):
        self._g.create_formats_()

    def reverse(self):
        from .transforms import reverse
        return HeteroGraphIndex(reverse(self._g, copy_ndata=False, copy_edata=False))

    # ---- mutation (the reference's index is immutable too: mutation builds a new one) ----------
    def add_nodes(self, ntype, num):
        self._g.add_nodes(num, ntype=self._ntype(ntype))

    def add_edges(self, etype, u, v):
        self._g.add_edges(u, v, etype=self._etype(etype))

    def add_edge(self, etype, u, v):
        self.add_edges(etype, [u], [v])

    def clear(self):
        raise DGLError("HeteroGraphIndex.clear is not supported (the reference raises here as well).")

    # ---- queries -----------------------------------------------------------------------------------
    def is_multigraph(self):
        return bool(self._g.is_multigraph)

    def is_readonly(self):
        return True

    def num_nodes(self, ntype):
        return self._g.num_nodes(self._ntype(ntype))

    number_of_nodes = num_nodes

    def num_edges(self, etype):
        return self._g.num_edges(self._etype(etype))

    number_of_edges = num_edges

    def has_nodes(self, ntype, vids):
        return self._g.has_nodes(vids, ntype=self._ntype(ntype))

    def has_edges_between(self, etype, u, v):
        return self._g.has_edges_between(u, v, etype=self._etype(etype))

    def predecessors(self, etype, v):
        return self._g.predecessors(v, etype=self._etype(etype))

    def successors(self, etype, v):
        return self._g.successors(v, etype=self._etype(etype))

    def edge_ids_all(self, etype, u, v):
        return self._g.edge_ids(u, v, return_uv=True, etype=self._etype(etype))

    def edge_ids_one(self, etype, u, v):
        return self._g.edge_ids(u, v, etype=self._etype(etype))

    def find_edges(self, etype, eid):
        u, v = self._g.find_edges(eid, etype=self._etype(etype))
        return u, v, torch.as_tensor(eid)

    def in_edges(self, etype, v):
        return self._g.in_edges(v, form="all", etype=self._etype(etype))

    def out_edges(self, etype, v):
        return self._g.out_edges(v, form="all", etype=self._etype(etype))

    def edges(self, etype, order=None):
        return self._g.edges(form="all", order=order or "eid", etype=self._etype(etype))

    def in_degrees(self, etype, v):
        return self._g.in_degrees(v, etype=self._etype(etype))

    def out_degrees(self, etype, v):
        return self._g.out_degrees(v, etype=self._etype(etype))

    def adjacency_matrix_tensors(self, etype, transpose, fmt):
        """Raw arrays of one relation: ``fmt='coo'`` -> (row, col); ``'csr'`` -> (indptr, indices, edge ids).
        Rows are sources unless ``transpose``."""
        ug = self._g._graphs[etype]
        if fmt == "coo":
            u, v = ug.coo()
            return (v, u) if transpose else (u, v)
        if fmt == "csr":
            ip, idx, e = ug.csc() if transpose else ug.csr()
            if e is None:
                e = torch.arange(idx.shape[0], dtype=idx.dtype, device=idx.device)
            return ip, idx, e
        raise DGLError(f"unsupported adjacency format {fmt}")

    def adjacency_matrix(self, etype, transpose, ctx):
        u, v = self._g._graphs[etype].coo()
        ug = self._g._graphs[etype]
        idx = torch.stack([v, u] if not transpose else [u, v]).long()
        shape = (ug.num_dst, ug.num_src) if not transpose else (ug.num_src, ug.num_dst)
        adj = torch.sparse_coo_tensor(idx, torch.ones(idx.shape[1]), shape).to(ctx)
        return adj, None

    def adjacency_matrix_scipy(self, etype, transpose, fmt, return_edge_ids=None):
        import scipy.sparse as sp
        ug = self._g._graphs[etype]
        u, v = [t.cpu().numpy() for t in ug.coo()]
        data = torch.arange(len(u)).numpy() if return_edge_ids else torch.ones(len(u)).numpy()
        rows, cols, shape = (v, u, (ug.num_dst, ug.num_src)) if not transpose else (u, v, (ug.num_src, ug.num_dst))
        m = sp.coo_matrix((data, (rows, cols)), shape=shape)
        return m.tocsr() if fmt == "csr" else m

    def incidence_matrix(self, etype, typestr, ctx):
        return self._g.incidence_matrix(typestr, ctx=ctx, etype=self._etype(etype)), None

    def get_csr_shuffle_order(self, etype):
        """Edge ids in CSR order of relation ``etype`` (identity when the COO is already row-sorted)."""
        ug = self._g._graphs[etype]
        e = ug.csr()[2]
        return e if e is not None else torch.arange(ug.num_edges(), dtype=ug.idtype, device=ug.device)

    # ---- subgraphs ---------------------------------------------------------------------------------
    def node_subgraph(self, induced_nodes):
        from .base import EID, NID
        sg = self._g.subgraph({self._ntype(i): n for i, n in enumerate(induced_nodes)}, relabel_nodes=True,
                              store_ids=True)
        return HeteroSubgraphIndex(sg, [sg.nodes[t].data[NID] for t in sg.ntypes],
                                   [sg.edges[c].data[EID] for c in sg.canonical_etypes])

    def edge_subgraph(self, induced_edges, preserve_nodes):
        from .base import EID, NID
        sg = self._g.edge_subgraph({self._etype(i): e for i, e in enumerate(induced_edges)},
                                   relabel_nodes=not preserve_nodes, store_ids=True)
        nodes = [sg.nodes[t].data[NID] if NID in sg.nodes[t].data else torch.arange(sg.num_nodes(t)) for t in sg.ntypes]
        return HeteroSubgraphIndex(sg, nodes, [sg.edges[c].data[EID] for c in sg.canonical_etypes])

    def flatten_relations(self, etypes):
        from .convert import to_homogeneous
        return to_homogeneous(self._g.edge_type_subgraph([self._etype(e) for e in etypes]))

    @property
    def graph(self):
        """The wrapped :class:`DGLGraph`."""
        return self._g


class HeteroSubgraphIndex:
    """Subgraph structure plus the parent ids it was induced from."""

    def __init__(self, graph, induced_nodes, induced_edges):
        self._g, self._nodes, self._edges = graph, induced_nodes, induced_edges

    @property
    def graph(self):
        return HeteroGraphIndex(self._g)

    @property
    def induced_nodes(self):
        return self._nodes

    @property
    def induced_edges(self):
        return self._edges


# ------------------------------------------------------------------------------------------------ creators
def create_metagraph_index(ntypes, canonical_etypes):
    """(metagraph, ntypes, etypes, canonical etypes) with ids following the order of ``ntypes``."""
    ntypes = list(ntypes)
    pos = {t: i for i, t in enumerate(ntypes)}
    pairs = [(pos[s], pos[d]) for s, _, d in canonical_etypes]
    return _MetaGraph(len(ntypes), pairs), ntypes, [c[1] for c in canonical_etypes], [tuple(c) for c in canonical_etypes]


def create_unitgraph_from_coo(num_ntypes, num_src, num_dst, row, col, formats=("coo", "csr", "csc"), row_sorted=False,
                              col_sorted=False):
    ug = UnitGraph.from_coo(int(num_src), int(num_dst), row, col)
    return ug.with_formats(list(formats)) if set(formats) != {"coo", "csr", "csc"} else ug


def create_unitgraph_from_csr(num_ntypes, num_src, num_dst, indptr, indices, edge_ids, formats=("coo", "csr", "csc"),
                              transpose=False):
    eids = edge_ids if edge_ids is not None and len(edge_ids) else None
    make = UnitGraph.from_csc if transpose else UnitGraph.from_csr
    ug = make(int(num_src), int(num_dst), indptr, indices, eids)
    return ug.with_formats(list(formats)) if set(formats) != {"coo", "csr", "csc"} else ug


def create_heterograph_from_relations(metagraph, rel_graphs, num_nodes_per_type, ntypes=None, canonical_etypes=None):
    """Assemble relations (UnitGraphs) under a metagraph; returns the graph's :class:`HeteroGraphIndex`."""
    from .heterograph import DGLGraph
    n = metagraph.num_nodes()
    ntypes = list(ntypes) if ntypes is not None else [f"_N{i}" for i in range(n)]
    if canonical_etypes is None:
        canonical_etypes = [(ntypes[s], f"_E{i}", ntypes[d]) for i, (s, d) in enumerate(metagraph._pairs)]
    if num_nodes_per_type is None:
        counts = [0] * n
        for (s, d), ug in zip(metagraph._pairs, rel_graphs):
            counts[s], counts[d] = max(counts[s], ug.num_src), max(counts[d], ug.num_dst)
        num_nodes_per_type = counts
    g = DGLGraph(list(rel_graphs), ntypes, canonical_etypes, [int(x) for x in num_nodes_per_type])
    return HeteroGraphIndex(g)


def create_heterograph_from_shared_memory(name):
    from .convert import hetero_from_shared_memory
    g = hetero_from_shared_memory(name)
    return HeteroGraphIndex(g), g.ntypes, g.etypes


def joint_union(metagraph, gidx_list):
    """Relations of several graphs over the same node sets, side by side (reference :func:`dgl.merge` building block)."""
    from .merge import merge
    return HeteroGraphIndex(merge([x.graph for x in gidx_list]))


def disjoint_union(metagraph, graphs):
    from .batch import batch
    return HeteroGraphIndex(batch([x.graph for x in graphs]))


def disjoint_partition(graph, bnn_all_types, bne_all_types):
    """Split a batched index back given per-type node / edge counts of every component."""
    from .batch import unbatch
    g = graph.graph
    nn = {t: torch.as_tensor(bnn_all_types[i]) for i, t in enumerate(g.ntypes)}
    ne = {c: torch.as_tensor(bne_all_types[i]) for i, c in enumerate(g.canonical_etypes)}
    return [HeteroGraphIndex(x) for x in unbatch(g, node_split=nn, edge_split=ne)]


def slice_gidx(graph, num_nodes, start_nid, num_edges, start_eid):
    """Contiguous slice (per type) of a batched index: ``num_nodes[t]`` nodes from ``start_nid[t]`` and likewise edges."""
    g = graph.graph
    nodes = {t: torch.arange(int(start_nid[i]), int(start_nid[i]) + int(num_nodes[i]), dtype=g.idtype, device=g.device)
             for i, t in enumerate(g.ntypes)}
    sg = g.subgraph(nodes, store_ids=False)
    return HeteroGraphIndex(sg)


class FlattenedHeteroGraph:
    """Result of flattening several relations into one homogeneous relation (reference FlattenedHeteroGraph): the
    merged graph plus, per node / edge, the type it came from and its id within that type."""

    def __init__(self, graph, induced_srctype, induced_srctype_set, induced_srcid, induced_etype, induced_etype_set,
                 induced_eid, induced_dsttype, induced_dsttype_set, induced_dstid):
        self.graph = graph
        self.induced_srctype, self.induced_srctype_set, self.induced_srcid = induced_srctype, induced_srctype_set, induced_srcid
        self.induced_etype, self.induced_etype_set, self.induced_eid = induced_etype, induced_etype_set, induced_eid
        self.induced_dsttype, self.induced_dsttype_set, self.induced_dstid = induced_dsttype, induced_dsttype_set, induced_dstid


class HeteroPickleStates:
    """Picklable structure snapshot of a heterograph index: version, (metagraph pairs, node counts) and the COO arrays
    of every relation (reference HeteroPickleStates; the features are pickled by the frames, not here)."""

    def __init__(self, version=2, meta=None, arrays=None):
        self._version, self._meta, self._arrays = version, meta, list(arrays or [])

    @staticmethod
    def from_index(gidx):
        g = gidx.graph
        meta = (gidx.metagraph._pairs, [g.num_nodes(t) for t in g.ntypes], list(g.ntypes), list(g.canonical_etypes))
        arrays = [t for ug in g._graphs for t in ug.coo()]
        return HeteroPickleStates(2, meta, arrays)

    def to_index(self):
        pairs, counts, ntypes, cet = self._meta
        rels = [UnitGraph.from_coo(counts[s], counts[d], self._arrays[2 * i], self._arrays[2 * i + 1])
                for i, (s, d) in enumerate(pairs)]
        return create_heterograph_from_relations(_MetaGraph(len(ntypes), pairs), rels, counts, ntypes,
                                                 [tuple(c) for c in cet])

    @property
    def version(self):
        return self._version

    @property
    def meta(self):
        return self._meta

    @property
    def arrays(self):
        return self._arrays

    def __getstate__(self):
        return self._version, self._meta, self._arrays

    def __setstate__(self, state):
        self._version, self._meta, self._arrays = state
