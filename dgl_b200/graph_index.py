"""Per-relation sparse structure with lazily materialised COO / CSR / CSC.

This is the counterpart of the reference's C++ ``UnitGraph``
(/root/reference/src/graph/unit_graph.h:389-401, unit_graph.cc:1418-1519) but
built directly on torch tensors: one relation = (num_src, num_dst) bipartite
structure whose formats are created on demand, can be restricted
(``formats([...])``), pinned in place and moved between devices. g-SpMM
consumes CSC (rows = destinations), its backward consumes CSR; g-SDDMM consumes
COO (src/array/kernel.cc:24,228 in the reference).
"""
from __future__ import annotations

import torch

from .base import DGLError

ALL_FORMATS = ("coo", "csr", "csc")


def _stable_argsort(x):
    return torch.sort(x, stable=True)[1]


def coo_to_csr(row, col, num_rows, idtype):
    """Return (indptr, indices, eids) with rows sorted stably (edge order kept
    inside each row)."""
    E = row.shape[0]
    if E == 0:
        z = torch.zeros(num_rows + 1, dtype=idtype, device=row.device)
        return z, col.clone(), torch.empty(0, dtype=idtype, device=row.device)
    # sorted already? (common for generated/ sampled graphs)
    if E > 1 and bool((row[1:] >= row[:-1]).all()):
        perm = None
        srow, scol = row, col
    elif row.is_cuda and row.dtype == idtype and E < 2 ** 31:
        # native build: stable radix sort on the row bits with the edge-id permutation + binary-search row bounds
        # (csrc/cuda/structure_build.cu) - no torch.sort, one pass over the edge list per radix digit
        from .ops.cuda_struct import cuda_coo_to_csr
        return cuda_coo_to_csr(row.contiguous(), col, num_rows)
    elif not row.is_cuda and row.dtype == idtype and col.dtype == idtype:
        # host build: native stable counting sort (csrc/cpu/spmat.cc), O(E) instead of a comparison sort
        from .ops.cpu_struct import coo_to_csr as _native_coo_to_csr
        return _native_coo_to_csr(row, col, num_rows)
    else:
        perm = _stable_argsort(row)
        srow, scol = row[perm], col[perm]
    eids = perm.to(idtype) if perm is not None else None
    if row.is_cuda and row.dtype == idtype:
        from ._graph_kernels import row_counts_to_indptr
        return row_counts_to_indptr(srow.contiguous(), num_rows), scol.contiguous(), eids
    counts = torch.bincount(srow.long(), minlength=num_rows)
    indptr = torch.zeros(num_rows + 1, dtype=torch.int64, device=row.device)
    torch.cumsum(counts, 0, out=indptr[1:])
    return indptr.to(idtype), scol.contiguous(), eids


class UnitGraph:
    """One relation's adjacency. Edge ids are positions in the COO arrays."""

    __slots__ = ("num_src", "num_dst", "idtype", "_coo", "_csr", "_csc", "_formats", "_pinned", "_row_sorted",
                 "_degs", "_rev_of", "_dst_view")

    def __init__(self, num_src, num_dst, idtype, coo=None, csr=None, csc=None, formats=ALL_FORMATS):
        self.num_src = int(num_src)
        self.num_dst = int(num_dst)
        self.idtype = idtype
        self._coo = coo  # (row, col)
        self._csr = csr  # (indptr[num_src+1], indices(dst), eids|None)
        self._csc = csc  # (indptr[num_dst+1], indices(src), eids|None)
        self._formats = tuple(formats)
        self._pinned = False
        self._degs = {}
        self._dst_view = None
        self._rev_of = None  # set on reversed views: formats they build are cached on the viewed graph too

    # ---- construction -------------------------------------------------
    @staticmethod
    def from_coo(num_src, num_dst, row, col, formats=ALL_FORMATS):
        if row.dtype != col.dtype:
            raise DGLError("src/dst id dtypes differ")
        return UnitGraph(num_src, num_dst, row.dtype, coo=(row.contiguous(), col.contiguous()), formats=formats)

    @staticmethod
    def from_csr(num_src, num_dst, indptr, indices, eids=None, formats=ALL_FORMATS):
        return UnitGraph(num_src, num_dst, indptr.dtype, csr=(indptr, indices, eids), formats=formats)

    @staticmethod
    def from_csc(num_src, num_dst, indptr, indices, eids=None, formats=ALL_FORMATS):
        return UnitGraph(num_src, num_dst, indptr.dtype, csc=(indptr, indices, eids), formats=formats)

    # ---- properties ---------------------------------------------------
    @property
    def device(self):
        for f in (self._coo, self._csr, self._csc):
            if f is not None:
                return f[0].device
        return torch.device("cpu")

    def num_edges(self):
        if self._coo is not None:
            return int(self._coo[0].shape[0])
        if self._csr is not None:
            return int(self._csr[1].shape[0])
        return int(self._csc[1].shape[0])

    def created_formats(self):
        out = []
        if self._coo is not None:
            out.append("coo")
        if self._csr is not None:
            out.append("csr")
        if self._csc is not None:
            out.append("csc")
        return out

    def allowed_formats(self):
        return list(self._formats)

    def _check_allowed(self, fmt):
        if fmt not in self._formats:
            raise DGLError(f"format '{fmt}' is not allowed on this graph (allowed: {self._formats}); "
                           "use g.formats([...]) to clone with different restrictions")

    def _post_create(self, tensors):
        if self._pinned:
            for t in tensors:
                if t is not None and not t.is_pinned():
                    _pin_inplace(t)

    # ---- format access --------------------------------------------------
    def coo(self):
        if self._coo is None:
            self._check_allowed("coo")
            if self._csr is not None:
                indptr, indices, eids = self._csr
                deg = (indptr[1:] - indptr[:-1]).long()
                row = torch.repeat_interleave(torch.arange(self.num_src, device=indptr.device, dtype=self.idtype), deg)
                col = indices
                if eids is not None:
                    r2 = torch.empty_like(row)
                    c2 = torch.empty_like(col)
                    r2[eids.long()] = row
                    c2[eids.long()] = col
                    row, col = r2, c2
            else:
                indptr, indices, eids = self._csc
                deg = (indptr[1:] - indptr[:-1]).long()
                col = torch.repeat_interleave(torch.arange(self.num_dst, device=indptr.device, dtype=self.idtype), deg)
                row = indices
                if eids is not None:
                    r2 = torch.empty_like(row)
                    c2 = torch.empty_like(col)
                    r2[eids.long()] = row
                    c2[eids.long()] = col
                    row, col = r2, c2
            self._coo = (row.contiguous(), col.contiguous())
            self._post_create(self._coo)
        return self._coo

    def csr(self):
        """(indptr, indices=dst ids, eids|None), rows = source nodes."""
        if self._csr is None:
            self._check_allowed("csr")
            row, col = self.coo() if self._coo is not None or "coo" in self._formats else self._coo_temp()
            self._csr = coo_to_csr(row, col, self.num_src, self.idtype)
            self._post_create(self._csr)
            if self._rev_of is not None and self._rev_of._csc is None and "csc" in self._rev_of._formats:
                self._rev_of._csc = self._csr   # a format built on a reversed VIEW is cached on the graph it views
        return self._csr

    def csc(self):
        """(indptr, indices=src ids, eids|None), rows = destination nodes."""
        if self._csc is None:
            self._check_allowed("csc")
            row, col = self.coo() if self._coo is not None or "coo" in self._formats else self._coo_temp()
            self._csc = coo_to_csr(col, row, self.num_dst, self.idtype)
            self._post_create(self._csc)
            if self._rev_of is not None and self._rev_of._csr is None and "csr" in self._rev_of._formats:
                self._rev_of._csr = self._csc
        return self._csc

    def _coo_temp(self):
        # build COO without caching when 'coo' is a disallowed format
        saved = self._formats
        self._formats = ALL_FORMATS
        try:
            coo = self.coo()
        finally:
            self._formats = saved
            self._coo = None
        return coo

    def edges(self, order="eid"):
        """(src, dst, eid) with order in {'eid', 'srcdst', 'dst'... }"""
        if order in (None, "eid"):
            row, col = self.coo() if "coo" in self._formats or self._coo is not None else self._coo_temp()
            return row, col, None
        if order == "srcdst":
            indptr, indices, eids = self.csr() if ("csr" in self._formats or self._csr is not None) else \
                coo_to_csr(*self.coo(), self.num_src, self.idtype)
            deg = (indptr[1:] - indptr[:-1]).long()
            row = torch.repeat_interleave(torch.arange(self.num_src, device=indptr.device, dtype=self.idtype), deg)
            # sort dst within row for canonical srcdst order
            key = row.long() * max(self.num_dst, 1) + indices.long()
            perm = _stable_argsort(key)
            e = eids if eids is not None else torch.arange(indices.shape[0], device=indptr.device, dtype=self.idtype)
            return row[perm], indices[perm], e[perm]
        raise DGLError(f"unknown edge order {order}")

    def in_degrees(self):
        d = self._degs.get("in")
        if d is None:
            if self._csc is not None:
                ip = self._csc[0]
                d = ip[1:] - ip[:-1]
            else:
                _, col = self.coo() if self._coo is not None else self._coo_temp()
                d = torch.bincount(col.long(), minlength=self.num_dst).to(self.idtype)
            self._degs["in"] = d
        return d

    def out_degrees(self):
        d = self._degs.get("out")
        if d is None:
            if self._csr is not None:
                ip = self._csr[0]
                d = ip[1:] - ip[:-1]
            else:
                row, _ = self.coo() if self._coo is not None else self._coo_temp()
                d = torch.bincount(row.long(), minlength=self.num_src).to(self.idtype)
            self._degs["out"] = d
        return d

    # ---- transforms -----------------------------------------------------
    def dst_sorted_view(self):
        """(view, eid_map): a relation with the same edges whose edge ids are the positions of this graph's CSC, i.e.
        edges grouped by destination. g-SpMM / edge-softmax on the view need no edge-id indirection and g-SDDMM writes
        its result already in that order, so a layer that creates AND consumes its edge data internally (attention)
        never permutes an [E, ...] tensor. ``eid_map[i]`` = original id of view edge i (None if already ordered).
        The view is cached; the formats it builds (its CSR for the backward pass) live as long as the graph."""
        indptr, indices, eids = self.csc()
        if eids is None:
            return self, None
        v = getattr(self, "_dst_view", None)
        if v is None:
            v = UnitGraph.from_csc(self.num_src, self.num_dst, indptr, indices, None, ALL_FORMATS)
            self._dst_view = v
        return v, eids

    def reverse(self):
        """Transposed relation sharing storage (CSR <-> CSC swap)."""
        g = UnitGraph(self.num_dst, self.num_src, self.idtype,
                      coo=None if self._coo is None else (self._coo[1], self._coo[0]),
                      csr=self._csc, csc=self._csr, formats=_swap_formats(self._formats))
        g._pinned = self._pinned
        g._rev_of = self
        return g

    def _map(self, fn):
        def m(f):
            return None if f is None else tuple(None if t is None else fn(t) for t in f)

        g = UnitGraph(self.num_src, self.num_dst, self.idtype, m(self._coo), m(self._csr), m(self._csc), self._formats)
        return g

    def to(self, device, non_blocking=False):
        device = torch.device(device)
        if device == self.device:
            return self
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def astype(self, idtype):
        if idtype == self.idtype:
            return self
        g = self._map(lambda t: t.to(idtype))
        g.idtype = idtype
        return g

    def with_formats(self, formats):
        formats = tuple(formats)
        for f in formats:
            if f not in ALL_FORMATS:
                raise DGLError(f"unknown sparse format {f}")
        keep = {}
        for name in ALL_FORMATS:
            cur = getattr(self, "_" + name)
            keep[name] = cur if name in formats else None
        if all(v is None for v in keep.values()):
            # need at least one format: derive the first allowed one
            first = formats[0]
            keep[first] = getattr(self, first)() if first in self._formats else _derive(self, first)
        return UnitGraph(self.num_src, self.num_dst, self.idtype, keep["coo"], keep["csr"], keep["csc"], formats)

    def create_formats_(self):
        for f in self._formats:
            getattr(self, f)()

    # ---- pinning / streams ---------------------------------------------------
    def pin_memory_(self):
        if self.device.type != "cpu":
            raise DGLError("only CPU graphs can be pinned")
        for f in (self._coo, self._csr, self._csc):
            if f is not None:
                for t in f:
                    if t is not None and not t.is_pinned():
                        _pin_inplace(t)
        self._pinned = True

    def unpin_memory_(self):
        for f in (self._coo, self._csr, self._csc):
            if f is not None:
                for t in f:
                    if t is not None and t.is_pinned():
                        _unpin_inplace(t)
        self._pinned = False

    def is_pinned(self):
        return self._pinned

    def record_stream(self, stream):
        for f in (self._coo, self._csr, self._csc):
            if f is not None:
                for t in f:
                    if t is not None and t.is_cuda:
                        t.record_stream(stream)


def _derive(g, fmt):
    saved = g._formats
    g._formats = ALL_FORMATS
    try:
        return getattr(g, fmt)()
    finally:
        g._formats = saved


def _swap_formats(formats):
    m = {"coo": "coo", "csr": "csc", "csc": "csr"}
    return tuple(m[f] for f in formats)


def _pin_inplace(t):
    """cudaHostRegister the tensor's storage (in-place pin, reference:
    src/runtime/cuda/cuda_device_api.cc:238-247)."""
    if t.numel() == 0:
        return
    if torch.cuda.is_available():
        rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
        if int(rc) != 0:
            raise DGLError(f"cudaHostRegister failed with code {int(rc)}")


def _unpin_inplace(t):
    if t.numel() == 0:
        return
    if torch.cuda.is_available():
        torch.cuda.cudart().cudaHostUnregister(t.data_ptr())


# ======================================================================================================
# Legacy homogeneous index (reference python/dgl/graph_index.py: GraphIndex / SubgraphIndex and creators).
# The pre-0.5 API addressed one homogeneous graph through a ``GraphIndex`` handle. Here it is the same
# adapter as ``heterograph_index.HeteroGraphIndex`` pinned to the single relation, with the legacy
# singular method names (``has_node``, ``edge_id``, ``in_degree`` ...) and in-place mutation.
# ======================================================================================================
class BoolFlag:
    """Tri-state flag of the legacy C API."""
    BOOL_UNKNOWN = -1
    BOOL_FALSE = 0
    BOOL_TRUE = 1


class GraphIndex:
    """Structure-only handle on a homogeneous graph. Mutable unless created ``readonly``."""

    def __init__(self, graph=None, readonly=False):
        if graph is None:
            from .convert import graph as make_graph
            graph = make_graph(([], []))
        self._g = graph
        self._readonly = bool(readonly)

    # pickling keeps only structure
    def __getstate__(self):
        u, v = self._g.edges()
        return self._g.num_nodes(), self._readonly, u, v

    def __setstate__(self, state):
        from .convert import graph as make_graph
        n, self._readonly, u, v = state
        self._g = make_graph((u, v), num_nodes=n)

    @property
    def graph(self):
        return self._g

    # ---- mutation ------------------------------------------------------------------------------------
    def _writable(self):
        if self._readonly:
            raise DGLError("The graph index is read-only; create it with readonly=False to mutate it.")

    def add_nodes(self, num):
        self._writable()
        self._g.add_nodes(int(num))

    def add_edge(self, u, v):
        self.add_edges([u], [v])

    def add_edges(self, u, v):
        self._writable()
        u, v = torch.as_tensor(u).reshape(-1), torch.as_tensor(v).reshape(-1)
        if u.numel() == 1 and v.numel() > 1:
            u = u.expand_as(v)
        if v.numel() == 1 and u.numel() > 1:
            v = v.expand_as(u)
        self._g.add_edges(u, v)

    def clear(self):
        self._writable()
        self.__init__(None, False)

    def clear_cache(self):
        """Cached compressed formats are dropped and rebuilt on demand."""
        ug = self._g._graphs[0]
        u, v = ug.coo()
        self._g._graphs[0] = UnitGraph.from_coo(ug.num_src, ug.num_dst, u, v)

    def readonly(self, readonly_state=True):
        self._readonly = bool(readonly_state)
        return self

    def is_readonly(self):
        return self._readonly

    def is_multigraph(self):
        return bool(self._g.is_multigraph)

    # ---- queries --------------------------------------------------------------------------------------
    def num_nodes(self):
        return self._g.num_nodes()

    number_of_nodes = num_nodes

    def num_edges(self):
        return self._g.num_edges()

    number_of_edges = num_edges

    def has_node(self, vid):
        return 0 <= int(vid) < self._g.num_nodes()

    def has_nodes(self, vids):
        return self._g.has_nodes(torch.as_tensor(vids)).long()

    def has_edge_between(self, u, v):
        return bool(self._g.has_edges_between(torch.tensor([u]), torch.tensor([v]))[0])

    def has_edges_between(self, u, v):
        return self._g.has_edges_between(torch.as_tensor(u), torch.as_tensor(v)).long()

    def predecessors(self, v, radius=1):
        if radius != 1:
            raise DGLError("only radius 1 is supported")
        return self._g.predecessors(int(v))

    def successors(self, v, radius=1):
        if radius != 1:
            raise DGLError("only radius 1 is supported")
        return self._g.successors(int(v))

    def edge_id(self, u, v):
        """Ids of every edge u -> v."""
        _, _, e = self._g.edge_ids(torch.tensor([u]), torch.tensor([v]), return_uv=True)
        return e

    def edge_ids(self, u, v):
        return self._g.edge_ids(torch.as_tensor(u), torch.as_tensor(v), return_uv=True)

    def find_edge(self, eid):
        u, v = self._g.find_edges(torch.tensor([eid]))
        return int(u[0]), int(v[0])

    def find_edges(self, eid):
        eid = torch.as_tensor(eid)
        u, v = self._g.find_edges(eid)
        return u, v, eid

    def in_edges(self, v):
        return self._g.in_edges(torch.as_tensor(v).reshape(-1), form="all")

    def out_edges(self, v):
        return self._g.out_edges(torch.as_tensor(v).reshape(-1), form="all")

    def edges(self, order=None):
        return self._g.edges(form="all", order=order or "eid")

    def sort_csr(self):
        """Column-sort the rows of the cached CSR (a no-op contract here: queries never depend on in-row order)."""
        self._g._graphs[0].csr()

    def in_degree(self, v):
        return int(self._g.in_degrees(int(v)))

    def in_degrees(self, v):
        return self._g.in_degrees(torch.as_tensor(v))

    def out_degree(self, v):
        return int(self._g.out_degrees(int(v)))

    def out_degrees(self, v):
        return self._g.out_degrees(torch.as_tensor(v))

    # ---- subgraphs ---------------------------------------------------------------------------------------
    def node_subgraph(self, v):
        from .base import EID
        v = torch.as_tensor(v)
        sg = self._g.subgraph(v, store_ids=True)
        return SubgraphIndex(GraphIndex(sg, True), self, v, sg.edata[EID])

    def node_subgraphs(self, vs_arr):
        return [self.node_subgraph(v) for v in vs_arr]

    def node_halo_subgraph(self, v, num_hops):
        """Subgraph induced by ``v`` plus everything within ``num_hops`` in-hops (the halo)."""
        from .base import EID, NID
        v = torch.as_tensor(v)
        frontier, nodes = v, v
        for _ in range(int(num_hops)):
            frontier = torch.unique(self._g.in_edges(frontier)[0])
            nodes = torch.unique(torch.cat([nodes, frontier]))
        rest = nodes[~torch.isin(nodes, v)]
        order = torch.cat([v, rest])
        sg = self._g.subgraph(order, store_ids=True)
        return SubgraphIndex(GraphIndex(sg, True), self, sg.ndata[NID], sg.edata[EID])

    def edge_subgraph(self, e, preserve_nodes=False):
        from .base import NID
        e = torch.as_tensor(e)
        sg = self._g.edge_subgraph(e, relabel_nodes=not preserve_nodes, store_ids=True)
        nodes = sg.ndata[NID] if NID in sg.ndata else torch.arange(sg.num_nodes())
        return SubgraphIndex(GraphIndex(sg, True), self, nodes, e)

    # ---- matrices / conversion ------------------------------------------------------------------------------
    def adjacency_matrix_scipy(self, transpose, fmt, return_edge_ids=None):
        return self._g._graph.adjacency_matrix_scipy(0, transpose, fmt, return_edge_ids)

    def adjacency_matrix(self, transpose, ctx):
        return self._g._graph.adjacency_matrix(0, transpose, ctx)

    def incidence_matrix(self, typestr, ctx):
        return self._g._graph.incidence_matrix(0, typestr, ctx)

    def get_csr_shuffle_order(self):
        return self._g._graph.get_csr_shuffle_order(0)

    def get_immutable_gidx(self, ctx):
        return GraphIndex(self._g.to(ctx), True)

    def to_immutable(self):
        return GraphIndex(self._g, True)

    def to_networkx(self):
        import networkx as nx
        u, v, e = [t.tolist() for t in self.edges()]
        G = nx.MultiDiGraph() if self.is_multigraph() else nx.DiGraph()
        G.add_nodes_from(range(self.num_nodes()))
        for a, b, i in zip(u, v, e):
            G.add_edge(a, b, id=i)
        return G

    def line_graph(self, backtracking=True):
        from .transforms import line_graph
        return GraphIndex(line_graph(self._g, backtracking=backtracking), True)

    @property
    def ctx(self):
        return self._g.device

    @property
    def dtype(self):
        return "int32" if self._g.idtype == torch.int32 else "int64"

    def nbits(self):
        return 32 if self._g.idtype == torch.int32 else 64

    def bits_needed(self):
        return self._g._graph.bits_needed(0)

    def asbits(self, bits):
        return GraphIndex(self._g.int() if int(bits) == 32 else self._g.long(), self._readonly)

    def copy_to(self, ctx):
        return GraphIndex(self._g.to(ctx), self._readonly)

    def copyto_shared_mem(self, shared_mem_name):
        return GraphIndex(self._g.shared_memory(shared_mem_name), True)


class SubgraphIndex:
    """A subgraph's index together with its parent and the induced parent ids."""

    def __init__(self, graph, parent, induced_nodes, induced_edges):
        self.graph, self.parent = graph, parent
        self.induced_nodes, self.induced_edges = induced_nodes, induced_edges


def from_coo(num_nodes, src, dst, readonly):
    from .convert import graph as make_graph
    return GraphIndex(make_graph((torch.as_tensor(src), torch.as_tensor(dst)), num_nodes=int(num_nodes)), readonly)


def from_csr(indptr, indices, direction):
    """``direction='out'``: rows are sources; ``'in'``: rows are destinations."""
    from .convert import graph as make_graph
    fmt = "csr" if direction == "out" else "csc"
    return GraphIndex(make_graph((fmt, (torch.as_tensor(indptr), torch.as_tensor(indices), []))), True)


def from_shared_mem_graph_index(shared_mem_name):
    from .convert import hetero_from_shared_memory
    return GraphIndex(hetero_from_shared_memory(shared_mem_name), True)


def from_networkx(nx_graph, readonly):
    from .convert import from_networkx as conv
    import networkx as nx
    if not isinstance(nx_graph, nx.Graph):
        nx_graph = nx.DiGraph(nx_graph)
    has_id = nx_graph.number_of_edges() > 0 and "id" in next(iter(nx_graph.edges(data=True)))[-1]
    return GraphIndex(conv(nx_graph, edge_id_attr_name="id" if has_id else None), readonly)


def from_scipy_sparse_matrix(adj, readonly):
    from .convert import from_scipy
    return GraphIndex(from_scipy(adj), readonly)


def from_edge_list(elist, readonly):
    from .convert import graph as make_graph
    if isinstance(elist, tuple) and len(elist) == 2 and not isinstance(elist[0], (tuple, list)) or (
            isinstance(elist, tuple) and len(elist) == 2 and len(elist[0]) != 2):
        src, dst = elist
    else:
        src, dst = ([e[0] for e in elist], [e[1] for e in elist]) if len(elist) else ([], [])
    return GraphIndex(make_graph((torch.as_tensor(src, dtype=torch.int64), torch.as_tensor(dst, dtype=torch.int64))),
                      readonly)


def map_to_subgraph_nid(induced_nodes, parent_nids):
    """Positions of ``parent_nids`` inside ``induced_nodes`` (parent id -> subgraph id)."""
    induced = torch.as_tensor(induced_nodes).long()
    parent_nids = torch.as_tensor(parent_nids).long()
    order = torch.argsort(induced)
    pos = torch.searchsorted(induced[order], parent_nids)
    if pos.numel() and (bool((pos >= induced.numel()).any()) or not torch.equal(induced[order][pos], parent_nids)):
        raise DGLError("some parent node ids are not part of the subgraph")
    return order[pos]


def transform_ids(mapping, ids):
    """``mapping[ids]``."""
    return torch.as_tensor(mapping)[torch.as_tensor(ids).long()]


def disjoint_union(graphs):
    from .batch import batch
    return GraphIndex(batch([x.graph for x in graphs]), True)


def disjoint_partition(graph, num_or_size_splits):
    """Inverse of :func:`disjoint_union`: an int splits into that many equal parts, a list gives node counts."""
    from .batch import unbatch
    g = graph.graph
    if isinstance(num_or_size_splits, int):
        k = num_or_size_splits
        if g.num_nodes() % k:
            raise DGLError("graph cannot be split into equally sized parts")
        sizes = torch.full((k,), g.num_nodes() // k, dtype=torch.int64)
    else:
        sizes = torch.as_tensor(num_or_size_splits, dtype=torch.int64)
    bounds = torch.cumsum(sizes, 0)
    u, _ = g.edges()
    owner = torch.searchsorted(bounds, u.long(), right=True)
    esizes = torch.bincount(owner, minlength=sizes.shape[0])
    return [GraphIndex(x, True) for x in unbatch(g, node_split=sizes, edge_split=esizes)]


def create_graph_index(graph_data, readonly):
    """Anything networkx-like, a scipy matrix, an edge list, another index or ``None`` (empty mutable index)."""
    if isinstance(graph_data, GraphIndex):
        return graph_data
    if graph_data is None:
        if readonly:
            raise DGLError("can't create an empty immutable graph")
        return GraphIndex(None, False)
    if isinstance(graph_data, (list, tuple)):
        return from_edge_list(graph_data, readonly)
    if hasattr(graph_data, "tocoo"):
        return from_scipy_sparse_matrix(graph_data, readonly)
    try:
        return from_networkx(graph_data, readonly)
    except Exception as exc:  # noqa: BLE001
        raise DGLError(f'Error while creating graph from input of type "{type(graph_data)}".') from exc
