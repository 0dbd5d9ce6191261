"""DGLGraph / DGLBlock: the user-facing (heterogeneous) graph.

Parity: /root/reference/python/dgl/heterograph.py (DGLGraph :37, update_all
:5018, apply_edges :4597, send_and_recv :4714, pull :4842, push :4940,
multi_update_all :5161, prop_nodes :5283, to :5658, pin_memory_ :5773,
local_scope :6017, formats :6090, create_formats_ :6201, DGLBlock :6631).
Storage is a list of :class:`UnitGraph` (torch tensors, lazily materialised
COO/CSR/CSC) + per-type feature :class:`Frame`; no C++ graph object sits
behind it, so moving/pinning/recording streams are plain tensor operations.
"""
from __future__ import annotations

import copy
from collections import defaultdict
from contextlib import contextmanager

import torch

from . import core
from . import function as fn
from .base import ALL, EID, ETYPE, NID, NTYPE, DGLError, dgl_warning, is_all
from .frame import Frame
from .graph_index import ALL_FORMATS, UnitGraph
from .view import HeteroEdgeDataView, HeteroEdgeView, HeteroNodeDataView, HeteroNodeView

__all__ = ["DGLGraph", "DGLBlock", "DGLHeteroGraph", "combine_names"]


def _to_tensor(x, idtype, device):
    if torch.is_tensor(x):
        if x.dtype == torch.bool:
            x = torch.nonzero(x, as_tuple=False).squeeze(1)
        return x.to(device=device, dtype=idtype)
    if isinstance(x, int):
        return torch.tensor([x], dtype=idtype, device=device)
    return torch.as_tensor(x, dtype=idtype, device=device).reshape(-1)


class DGLGraph:
    """Heterogeneous graph with per-type node/edge features."""

    is_block = False

    # ------------------------------------------------------------------ init
    def __init__(self, graphs=None, ntypes=("_N",), canonical_etypes=(("_N", "_E", "_N"),), num_nodes=None,
                 node_frames=None, edge_frames=None, *, srctypes=None, dsttypes=None, **deprecated):
        if graphs is None:
            graphs = [UnitGraph.from_coo(0, 0, torch.empty(0, dtype=torch.int64), torch.empty(0, dtype=torch.int64))]
            num_nodes = [0]
        self._init(list(graphs), None if ntypes is None else list(ntypes), [tuple(c) for c in canonical_etypes], list(num_nodes), node_frames,
                   edge_frames, srctypes, dsttypes)

    def _init(self, graphs, ntypes, canonical_etypes, num_nodes, node_frames, edge_frames, srctypes=None,
              dsttypes=None):
        self._graphs = graphs
        self._canonical_etypes = canonical_etypes
        self._etypes = [c[1] for c in canonical_etypes]
        if srctypes is not None:  # block: node-type space = srctypes ++ dsttypes
            self._srctypes = list(srctypes)
            self._dsttypes = list(dsttypes)
            self._ntypes = self._srctypes + self._dsttypes
            self._is_unibipartite = True
        else:
            self._ntypes = ntypes
            self._srctypes = self._dsttypes = None
            st = {c[0] for c in canonical_etypes}
            dt = {c[2] for c in canonical_etypes}
            self._is_unibipartite = len(st & dt) == 0 and len(canonical_etypes) > 0
            # uni-bipartite (non-block) graphs expose the src/dst partition of their node types
            self._uni_src = [t for t in ntypes if t in st] if self._is_unibipartite else None
            self._uni_dst = [t for t in ntypes if t in dt] if self._is_unibipartite else None
        self._num_nodes = num_nodes
        assert len(self._num_nodes) == len(self._ntypes), (self._num_nodes, self._ntypes)
        self._ntype_to_id = {}
        if self._srctypes is not None:
            self._srctype_to_id = {t: i for i, t in enumerate(self._srctypes)}
            self._dsttype_to_id = {t: i + len(self._srctypes) for i, t in enumerate(self._dsttypes)}
        else:
            self._ntype_to_id = {t: i for i, t in enumerate(self._ntypes)}
        self._etype_to_canonical = {}
        for c in canonical_etypes:
            if c[1] in self._etype_to_canonical:
                self._etype_to_canonical[c[1]] = tuple()  # ambiguous
            else:
                self._etype_to_canonical[c[1]] = c
        self._canonical_to_id = {c: i for i, c in enumerate(canonical_etypes)}
        self._node_frames = node_frames if node_frames is not None else [Frame(num_rows=n) for n in self._num_nodes]
        self._edge_frames = edge_frames if edge_frames is not None else [Frame(num_rows=g.num_edges()) for g in graphs]
        self._batch_num_nodes = None
        self._batch_num_edges = None

    # ----------------------------------------------------------- type helpers
    @property
    def ntypes(self):
        return list(self._ntypes) if self._srctypes is None else list(self._ntypes)

    @property
    def etypes(self):
        return list(self._etypes)

    @property
    def canonical_etypes(self):
        return list(self._canonical_etypes)

    @property
    def srctypes(self):
        if self._srctypes is not None:
            return list(self._srctypes)
        return list(self._uni_src) if self._is_unibipartite else list(self._ntypes)

    @property
    def dsttypes(self):
        if self._dsttypes is not None:
            return list(self._dsttypes)
        return list(self._uni_dst) if self._is_unibipartite else list(self._ntypes)

    @property
    def is_unibipartite(self):
        return self._is_unibipartite

    @property
    def is_homogeneous(self):
        return len(self._ntypes) == 1 and len(self._etypes) == 1

    @property
    def is_multigraph(self):
        for g in self._graphs:
            row, col = g.coo() if "coo" in g.allowed_formats() else g._coo_temp()
            if row.numel() == 0:
                continue
            key = row.long() * max(g.num_dst, 1) + col.long()
            if torch.unique(key).numel() != key.numel():
                return True
        return False

    def metagraph(self):
        import networkx as nx
        mg = nx.MultiDiGraph()
        for nt in self.ntypes:
            mg.add_node(nt)
        for s, e, d in self.canonical_etypes:
            mg.add_edge(s, d, key=e)
        return mg

    def to_canonical_etype(self, etype):
        if etype is None:
            if len(self._etypes) != 1:
                raise DGLError("Edge type name must be specified if there are more than one edge types.")
            return self._canonical_etypes[0]
        if isinstance(etype, tuple):
            return etype
        ret = self._etype_to_canonical.get(etype)
        if ret is None:
            raise DGLError(f'Edge type "{etype}" does not exist.')
        if len(ret) == 0:
            raise DGLError(f'Edge type "{etype}" is ambiguous. Please use canonical edge type in the form of '
                           "(srctype, etype, dsttype)")
        return ret

    def get_ntype_id(self, ntype):
        if self._srctypes is not None:
            if ntype is None:
                if len(self._srctypes) != 1 and len(self._dsttypes) != 1:
                    raise DGLError("Node type name must be specified if there are more than one node types.")
                raise DGLError("Node type name must be specified for a block (use srcdata/dstdata).") \
                    if (len(self._srctypes) != 1 or len(self._dsttypes) != 1 or self._num_nodes[0] != self._num_nodes[-1]) \
                    else DGLError("ambiguous node type on a block")
            i = self._srctype_to_id.get(ntype, self._dsttype_to_id.get(ntype))
            if i is None:
                raise DGLError(f'Node type "{ntype}" does not exist.')
            return i
        if ntype is None:
            if len(self._ntypes) != 1:
                raise DGLError("Node type name must be specified if there are more than one node types.")
            return 0
        i = self._ntype_to_id.get(ntype)
        if i is None:
            raise DGLError(f'Node type "{ntype}" does not exist.')
        return i

    def get_ntype_id_from_src(self, ntype):
        if self._srctypes is None:
            if ntype is None and self._is_unibipartite and len(self._uni_src) == 1:
                return self._ntype_to_id[self._uni_src[0]]
            return self.get_ntype_id(ntype)
        if ntype is None:
            if len(self._srctypes) != 1:
                raise DGLError("SRC node type name must be specified if there are more than one SRC node types.")
            return 0
        i = self._srctype_to_id.get(ntype)
        if i is None:
            raise DGLError(f'SRC node type "{ntype}" does not exist.')
        return i

    def get_ntype_id_from_dst(self, ntype):
        if self._dsttypes is None:
            if ntype is None and self._is_unibipartite and len(self._uni_dst) == 1:
                return self._ntype_to_id[self._uni_dst[0]]
            return self.get_ntype_id(ntype)
        if ntype is None:
            if len(self._dsttypes) != 1:
                raise DGLError("DST node type name must be specified if there are more than one DST node types.")
            return len(self._srctypes)
        i = self._dsttype_to_id.get(ntype)
        if i is None:
            raise DGLError(f'DST node type "{ntype}" does not exist.')
        return i

    def get_etype_id(self, etype):
        if etype is None:
            if len(self._etypes) != 1:
                raise DGLError("Edge type name must be specified if there are more than one edge types.")
            return 0
        c = self.to_canonical_etype(etype)
        i = self._canonical_to_id.get(c)
        if i is None:
            raise DGLError(f'Edge type "{etype}" does not exist.')
        return i

    def _src_ntid(self, etid):
        s = self._canonical_etypes[etid][0]
        return self._srctype_to_id[s] if self._srctypes is not None else self._ntype_to_id[s]

    def _dst_ntid(self, etid):
        d = self._canonical_etypes[etid][2]
        return self._dsttype_to_id[d] if self._dsttypes is not None else self._ntype_to_id[d]

    # ------------------------------------------------------------ basic props
    @property
    def idtype(self):
        return self._graphs[0].idtype if self._graphs else torch.int64

    @property
    def device(self):
        return self._graphs[0].device if self._graphs else torch.device("cpu")

    def num_nodes(self, ntype=None):
        if ntype is None:
            if self._srctypes is not None:
                return sum(self._num_nodes)
            return sum(self._num_nodes)
        return self._num_nodes[self.get_ntype_id(ntype)]

    number_of_nodes = num_nodes

    def num_src_nodes(self, ntype=None):
        if self._srctypes is None:
            if ntype is None and self._is_unibipartite:
                return sum(self._num_nodes[self._ntype_to_id[t]] for t in self._uni_src)
            return self.num_nodes(ntype)
        if ntype is None:
            return sum(self._num_nodes[: len(self._srctypes)])
        return self._num_nodes[self.get_ntype_id_from_src(ntype)]

    number_of_src_nodes = num_src_nodes

    def num_dst_nodes(self, ntype=None):
        if self._dsttypes is None:
            if ntype is None and self._is_unibipartite:
                return sum(self._num_nodes[self._ntype_to_id[t]] for t in self._uni_dst)
            return self.num_nodes(ntype)
        if ntype is None:
            return sum(self._num_nodes[len(self._srctypes):])
        return self._num_nodes[self.get_ntype_id_from_dst(ntype)]

    number_of_dst_nodes = num_dst_nodes

    def num_edges(self, etype=None):
        if etype is None:
            return sum(g.num_edges() for g in self._graphs)
        return self._graphs[self.get_etype_id(etype)].num_edges()

    number_of_edges = num_edges

    def __len__(self):
        return self.num_nodes()

    def __contains__(self, vid):
        return bool(self.has_nodes(vid))

    # -------------------------------------------------------------- views
    @property
    def nodes(self):
        return HeteroNodeView(self, self.get_ntype_id)

    @property
    def srcnodes(self):
        return HeteroNodeView(self, self.get_ntype_id_from_src)

    @property
    def dstnodes(self):
        return HeteroNodeView(self, self.get_ntype_id_from_dst)

    @property
    def edges(self):
        return HeteroEdgeView(self)

    @property
    def ndata(self):
        if len(self._ntypes) == 1:
            return HeteroNodeDataView(self, self._ntypes[0], 0, ALL)
        return HeteroNodeDataView(self, None, list(range(len(self._ntypes))), ALL)

    @property
    def srcdata(self):
        if self._srctypes is None:
            if self._is_unibipartite:
                ids = [self._ntype_to_id[t] for t in self._uni_src]
                return HeteroNodeDataView(self, self._uni_src[0], ids[0], ALL) if len(ids) == 1 else \
                    HeteroNodeDataView(self, None, ids, ALL)
            return self.ndata
        if len(self._srctypes) == 1:
            return HeteroNodeDataView(self, self._srctypes[0], 0, ALL)
        return HeteroNodeDataView(self, None, list(range(len(self._srctypes))), ALL)

    @property
    def dstdata(self):
        if self._dsttypes is None:
            if self._is_unibipartite:
                ids = [self._ntype_to_id[t] for t in self._uni_dst]
                return HeteroNodeDataView(self, self._uni_dst[0], ids[0], ALL) if len(ids) == 1 else \
                    HeteroNodeDataView(self, None, ids, ALL)
            return self.ndata
        off = len(self._srctypes)
        if len(self._dsttypes) == 1:
            return HeteroNodeDataView(self, self._dsttypes[0], off, ALL)
        return HeteroNodeDataView(self, None, list(range(off, off + len(self._dsttypes))), ALL)

    @property
    def edata(self):
        if len(self._canonical_etypes) == 1:
            return HeteroEdgeDataView(self, self._canonical_etypes[0], ALL)
        return HeteroEdgeDataView(self, None, ALL)

    def _set_n_repr(self, ntid, u, data):
        fr = self._node_frames[ntid]
        n = self._num_nodes[ntid]
        for k, v in data.items():
            if torch.is_tensor(v) and v.device != self.device and not (v.device.type == "cpu" and self.is_pinned()):
                raise DGLError(f"Cannot assign node feature \"{k}\" on device {v.device} to a graph on device "
                               f"{self.device}. Call DGLGraph.to() to copy the graph to the same device.")
        if is_all(u):
            for k, v in data.items():
                if torch.is_tensor(v) and v.shape[0] != n:
                    raise DGLError(f"Expect number of features to match number of nodes (len(u)). Got "
                                   f"{v.shape[0]} and {n} instead.")
                fr.update_column(k, v)
        else:
            u = _to_tensor(u, self.idtype, self.device)
            fr.update_row(u, data)

    def _set_e_repr(self, etid, edges, data):
        fr = self._edge_frames[etid]
        n = self._graphs[etid].num_edges()
        for k, v in data.items():
            if torch.is_tensor(v) and v.device != self.device and not (v.device.type == "cpu" and self.is_pinned()):
                raise DGLError(f"Cannot assign edge feature \"{k}\" on device {v.device} to a graph on device "
                               f"{self.device}.")
        if is_all(edges):
            for k, v in data.items():
                if torch.is_tensor(v) and v.shape[0] != n:
                    raise DGLError(f"Expect number of features to match number of edges. Got {v.shape[0]} and "
                                   f"{n} instead.")
                fr.update_column(k, v)
        else:
            eid = self._parse_edges(edges, etid)
            fr.update_row(eid, data)

    def _parse_edges(self, edges, etid):
        if isinstance(edges, tuple) and len(edges) == 2:
            u, v = edges
            return self.edge_ids(u, v, etype=self._canonical_etypes[etid])
        return _to_tensor(edges, self.idtype, self.device)

    def node_attr_schemes(self, ntype=None):
        return self._node_frames[self.get_ntype_id(ntype)].schemes

    def edge_attr_schemes(self, etype=None):
        return self._edge_frames[self.get_etype_id(etype)].schemes

    def set_n_initializer(self, initializer, field=None, ntype=None):
        self._node_frames[self.get_ntype_id(ntype)].set_initializer(initializer, field)

    def set_e_initializer(self, initializer, field=None, etype=None):
        self._edge_frames[self.get_etype_id(etype)].set_initializer(initializer, field)

    # ------------------------------------------------------- structure query
    def has_nodes(self, vid, ntype=None):
        n = self.num_nodes(ntype) if (ntype is not None or len(self._ntypes) == 1) else self.num_nodes()
        if isinstance(vid, int):
            return 0 <= vid < n
        v = _to_tensor(vid, self.idtype, self.device)
        return (v >= 0) & (v < n)

    def _rel(self, etype):
        return self._graphs[self.get_etype_id(etype)]

    def has_edges_between(self, u, v, etype=None):
        g = self._rel(etype)
        scalar = isinstance(u, int) and isinstance(v, int)
        u = _to_tensor(u, self.idtype, self.device)
        v = _to_tensor(v, self.idtype, self.device)
        if u.numel() and (int(u.max()) >= g.num_src or int(u.min()) < 0):
            raise DGLError("u contains invalid node IDs")
        if v.numel() and (int(v.max()) >= g.num_dst or int(v.min()) < 0):
            raise DGLError("v contains invalid node IDs")
        row, col = g.coo()
        key = torch.sort(row.long() * max(g.num_dst, 1) + col.long())[0]
        q = u.long() * max(g.num_dst, 1) + v.long()
        pos = torch.searchsorted(key, q).clamp(max=max(key.numel() - 1, 0))
        res = (key[pos] == q) if key.numel() else torch.zeros_like(q, dtype=torch.bool)
        return bool(res[0]) if scalar else res

    def edge_ids(self, u, v, return_uv=False, etype=None):
        g = self._rel(etype)
        scalar = isinstance(u, int) and isinstance(v, int)
        u = _to_tensor(u, self.idtype, self.device)
        v = _to_tensor(v, self.idtype, self.device)
        row, col = g.coo()
        nd = max(g.num_dst, 1)
        key = row.long() * nd + col.long()
        skey, perm = torch.sort(key, stable=True)
        q = u.long() * nd + v.long()
        lo = torch.searchsorted(skey, q)
        if return_uv:
            hi = torch.searchsorted(skey, q, right=True)
            cnt = hi - lo
            idx = torch.repeat_interleave(lo, cnt) + (torch.arange(int(cnt.sum()), device=lo.device)
                                                      - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
            eid = perm[idx].to(self.idtype)
            return torch.repeat_interleave(u, cnt), torch.repeat_interleave(v, cnt), eid
        lo_c = lo.clamp(max=max(skey.numel() - 1, 0))
        ok = (skey[lo_c] == q) if skey.numel() else torch.zeros_like(q, dtype=torch.bool)
        if not bool(ok.all()):
            bad = int(torch.nonzero(~ok)[0])
            raise DGLError(f"No edge exists between nodes {int(u[bad])} and {int(v[bad])}.")
        eid = perm[lo_c].to(self.idtype)
        return int(eid[0]) if scalar else eid

    def find_edges(self, eid, etype=None):
        g = self._rel(etype)
        eid = _to_tensor(eid, self.idtype, self.device)
        if eid.numel() and (int(eid.max()) >= g.num_edges() or int(eid.min()) < 0):
            raise DGLError("eid contains invalid edge IDs")
        row, col = g.coo()
        return row[eid.long()], col[eid.long()]

    def all_edges(self, form="uv", order="eid", etype=None):
        g = self._rel(etype)
        src, dst, eid = g.edges(order)
        if eid is None:
            eid = torch.arange(src.shape[0], dtype=self.idtype, device=src.device)
        if form == "all":
            return src, dst, eid
        if form == "uv":
            return src, dst
        if form == "eid":
            return eid
        raise DGLError(f'Invalid form: {form}. Must be "all", "uv" or "eid".')

    def in_edges(self, v, form="uv", etype=None):
        g = self._rel(etype)
        v = _to_tensor(v, self.idtype, self.device)
        indptr, indices, eids = g.csc()
        s = indptr[v.long()].long()
        cnt = indptr[v.long() + 1].long() - s
        pos = torch.repeat_interleave(s, cnt) + (torch.arange(int(cnt.sum()), device=s.device)
                                                 - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
        src = indices[pos]
        dst = torch.repeat_interleave(v, cnt)
        eid = eids[pos] if eids is not None else pos.to(self.idtype)
        return _form(src, dst, eid, form)

    def out_edges(self, u, form="uv", etype=None):
        g = self._rel(etype)
        u = _to_tensor(u, self.idtype, self.device)
        indptr, indices, eids = g.csr()
        s = indptr[u.long()].long()
        cnt = indptr[u.long() + 1].long() - s
        pos = torch.repeat_interleave(s, cnt) + (torch.arange(int(cnt.sum()), device=s.device)
                                                 - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt))
        dst = indices[pos]
        src = torch.repeat_interleave(u, cnt)
        eid = eids[pos] if eids is not None else pos.to(self.idtype)
        return _form(src, dst, eid, form)

    def predecessors(self, v, etype=None):
        return self.in_edges(v, etype=etype)[0]

    def successors(self, v, etype=None):
        return self.out_edges(v, etype=etype)[1]

    def in_degrees(self, v=ALL, etype=None):
        g = self._rel(etype)
        d = g.in_degrees()
        if is_all(v):
            return d
        if isinstance(v, int):
            return int(d[v])
        return d[_to_tensor(v, self.idtype, self.device).long()]

    def out_degrees(self, u=ALL, etype=None):
        g = self._rel(etype)
        d = g.out_degrees()
        if is_all(u):
            return d
        if isinstance(u, int):
            return int(d[u])
        return d[_to_tensor(u, self.idtype, self.device).long()]

    def adj_tensors(self, fmt, etype=None):
        g = self._rel(etype)
        if fmt == "coo":
            return g.coo()
        if fmt == "csr":
            ip, idx, e = g.csr()
        elif fmt == "csc":
            ip, idx, e = g.csc()
        else:
            raise DGLError(f"unknown format {fmt}")
        if e is None:
            e = torch.arange(idx.shape[0], dtype=self.idtype, device=idx.device)
        return ip, idx, e

    def adj(self, etype=None, eweight_name=None):
        """Adjacency as a :class:`dgl_b200.sparse.SparseMatrix` (rows = src)."""
        from .sparse import spmatrix
        etid = self.get_etype_id(etype)
        g = self._graphs[etid]
        row, col = g.coo()
        val = self._edge_frames[etid][eweight_name] if eweight_name is not None else None
        return spmatrix(torch.stack([row, col]), val, shape=(g.num_src, g.num_dst))

    def adj_external(self, transpose=False, ctx=None, scipy_fmt=None, etype=None):
        g = self._rel(etype)
        row, col = g.coo()
        if transpose:
            row, col = col, row
            shape = (g.num_dst, g.num_src)
        else:
            shape = (g.num_src, g.num_dst)
        if scipy_fmt is None:
            idx = torch.stack([row.long(), col.long()])
            return torch.sparse_coo_tensor(idx, torch.ones(row.shape[0], device=row.device), shape)
        import scipy.sparse as sp
        m = sp.coo_matrix((torch.ones(row.shape[0]).numpy(), (row.cpu().numpy(), col.cpu().numpy())), shape=shape)
        return m.asformat(scipy_fmt)

    def inc(self, typestr, ctx=None, etype=None):
        g = self._rel(etype)
        row, col = g.coo()
        E = row.shape[0]
        eid = torch.arange(E, device=row.device)
        if typestr == "in":
            idx = torch.stack([col.long(), eid]); val = torch.ones(E); n = g.num_dst
        elif typestr == "out":
            idx = torch.stack([row.long(), eid]); val = torch.ones(E); n = g.num_src
        elif typestr == "both":
            mask = row != col
            idx = torch.cat([torch.stack([row.long()[mask], eid[mask]]), torch.stack([col.long()[mask], eid[mask]])], 1)
            val = torch.cat([-torch.ones(int(mask.sum())), torch.ones(int(mask.sum()))]); n = g.num_src
        else:
            raise DGLError("Invalid incidence matrix type: %s" % typestr)
        return torch.sparse_coo_tensor(idx, val.to(row.device), (n, E))

    # ------------------------------------------------------------- slicing
    def __getitem__(self, key):
        err = "Invalid slice syntax. Use G['etype'] or G['srctype', 'etype', 'dsttype'] to get view of one relation."
        if not isinstance(key, tuple):
            key = (slice(None), key, slice(None))
        if len(key) != 3:
            raise DGLError(err)
        etids = []
        for i, (s, e, d) in enumerate(self._canonical_etypes):
            if (is_all(key[0]) or key[0] == s) and (is_all(key[1]) or key[1] == e) and (is_all(key[2]) or key[2] == d):
                etids.append(i)
        if not etids:
            raise DGLError(f'Invalid key "{key}". Must be one of the edge types.')
        if len(etids) == 1:
            return self._single_relation_graph(etids[0])
        # multiple relations: flatten to one homogenised relation
        return self._flatten(etids)

    @property
    def _graph(self):
        """Id-based index view (the reference's ``g._graph`` handle; see :mod:`dgl_b200.heterograph_index`)."""
        from .heterograph_index import HeteroGraphIndex
        return HeteroGraphIndex(self)

    def _single_relation_graph(self, etid):
        s, e, d = self._canonical_etypes[etid]
        sid, did = self._src_ntid(etid), self._dst_ntid(etid)
        if s == d and self._srctypes is None:
            g = DGLGraph.__new__(DGLGraph)
            g._init([self._graphs[etid]], [s], [(s, e, d)], [self._num_nodes[sid]], [self._node_frames[sid]],
                    [self._edge_frames[etid]])
        else:
            g = DGLGraph.__new__(DGLGraph)
            g._init([self._graphs[etid]], None, [(s, e, d)], [self._num_nodes[sid], self._num_nodes[did]],
                    [self._node_frames[sid], self._node_frames[did]], [self._edge_frames[etid]],
                    srctypes=[s], dsttypes=[d])
        return g

    def _flatten(self, etids):
        srcs = sorted({self._canonical_etypes[i][0] for i in etids})
        dsts = sorted({self._canonical_etypes[i][2] for i in etids})
        soff, doff, o = {}, {}, 0
        for t in srcs:
            soff[t] = o
            o += self.num_src_nodes(t) if self._srctypes is not None else self.num_nodes(t)
        nsrc = o
        o = 0
        for t in dsts:
            doff[t] = o
            o += self.num_dst_nodes(t) if self._dsttypes is not None else self.num_nodes(t)
        ndst = o
        rows, cols = [], []
        for i in etids:
            s, _, d = self._canonical_etypes[i]
            r, c = self._graphs[i].coo()
            rows.append(r + soff[s]); cols.append(c + doff[d])
        ug = UnitGraph.from_coo(nsrc, ndst, torch.cat(rows), torch.cat(cols))
        g = DGLGraph.__new__(DGLGraph)
        stn, etn, dtn = combine_names(srcs), combine_names([self._etypes[i] for i in etids]), combine_names(dsts)
        if srcs == dsts and self._srctypes is None:
            g._init([ug], [stn], [(stn, etn, dtn)], [nsrc], None, None)
        else:
            g._init([ug], None, [(stn, etn, dtn)], [nsrc, ndst], None, None, srctypes=[stn], dsttypes=[dtn])
        return g

    def node_type_subgraph(self, ntypes, output_device=None):
        from .subgraph import node_type_subgraph
        return node_type_subgraph(self, ntypes, output_device)

    def edge_type_subgraph(self, etypes, output_device=None):
        from .subgraph import edge_type_subgraph
        return edge_type_subgraph(self, etypes, output_device)

    def subgraph(self, nodes, *, relabel_nodes=True, store_ids=True, output_device=None):
        from .subgraph import node_subgraph
        return node_subgraph(self, nodes, relabel_nodes=relabel_nodes, store_ids=store_ids, output_device=output_device)

    def edge_subgraph(self, edges, *, relabel_nodes=True, store_ids=True, output_device=None):
        from .subgraph import edge_subgraph
        return edge_subgraph(self, edges, relabel_nodes=relabel_nodes, store_ids=store_ids, output_device=output_device)

    def in_subgraph(self, nodes, **kw):
        from .subgraph import in_subgraph
        return in_subgraph(self, nodes, **kw)

    def out_subgraph(self, nodes, **kw):
        from .subgraph import out_subgraph
        return out_subgraph(self, nodes, **kw)

    def reverse(self, copy_ndata=True, copy_edata=False):
        from .transforms import reverse
        return reverse(self, copy_ndata=copy_ndata, copy_edata=copy_edata)

    def add_self_loop(self, edge_feat_names=None, fill_data=1.0, etype=None):
        from .transforms import add_self_loop
        return add_self_loop(self, edge_feat_names=edge_feat_names, fill_data=fill_data, etype=etype)

    def remove_self_loop(self, etype=None):
        from .transforms import remove_self_loop
        return remove_self_loop(self, etype=etype)

    def to_simple(self, **kw):
        from .transforms import to_simple
        return to_simple(self, **kw)

    def line_graph(self, backtracking=True, shared=False):
        from .transforms import line_graph
        return line_graph(self, backtracking=backtracking, shared=shared)

    def to_networkx(self, node_attrs=None, edge_attrs=None):
        from .convert import to_networkx
        return to_networkx(self, node_attrs, edge_attrs)

    # ------------------------------------------------------------ mutation
    def add_nodes(self, num, data=None, ntype=None):
        ntid = self.get_ntype_id(ntype)
        if num == 0:
            return
        self._num_nodes[ntid] += num
        self._node_frames[ntid].add_rows(num)
        if data is not None:
            ids = torch.arange(self._num_nodes[ntid] - num, self._num_nodes[ntid])
            self._node_frames[ntid].update_row(ids, data)
        nt = self._ntypes[ntid]
        for i, (s, _, d) in enumerate(self._canonical_etypes):
            g = self._graphs[i]
            ns = self._num_nodes[self._src_ntid(i)]
            nd = self._num_nodes[self._dst_ntid(i)]
            if g.num_src != ns or g.num_dst != nd:
                row, col = g.coo()
                self._graphs[i] = UnitGraph.from_coo(ns, nd, row, col, g.allowed_formats())
        self._reset_batch()

    def add_edges(self, u, v, data=None, etype=None):
        etid = self.get_etype_id(etype)
        u = _to_tensor(u, self.idtype, self.device)
        v = _to_tensor(v, self.idtype, self.device)
        if u.numel() == 1 and v.numel() > 1:
            u = u.expand(v.numel())
        if v.numel() == 1 and u.numel() > 1:
            v = v.expand(u.numel())
        if u.numel() == 0:
            return
        sid, did = self._src_ntid(etid), self._dst_ntid(etid)
        need_s, need_d = int(u.max()) + 1, int(v.max()) + 1
        if sid == did:
            need = max(need_s, need_d)
            if need > self._num_nodes[sid]:
                self.add_nodes(need - self._num_nodes[sid], ntype=self._ntypes[sid])
        else:
            if need_s > self._num_nodes[sid]:
                self.add_nodes(need_s - self._num_nodes[sid], ntype=self._ntypes[sid])
            if need_d > self._num_nodes[did]:
                self.add_nodes(need_d - self._num_nodes[did], ntype=self._ntypes[did])
        g = self._graphs[etid]
        row, col = g.coo()
        old = row.shape[0]
        self._graphs[etid] = UnitGraph.from_coo(self._num_nodes[sid], self._num_nodes[did], torch.cat([row, u]),
                                                torch.cat([col, v]), g.allowed_formats())
        self._edge_frames[etid].add_rows(u.numel())
        if data is not None:
            self._edge_frames[etid].update_row(torch.arange(old, old + u.numel()), data)
        self._reset_batch()

    def remove_edges(self, eids, etype=None, store_ids=False):
        etid = self.get_etype_id(etype)
        eids = _to_tensor(eids, self.idtype, self.device)
        g = self._graphs[etid]
        E = g.num_edges()
        if eids.numel() and (int(eids.max()) >= E or int(eids.min()) < 0):
            raise DGLError("eids contains invalid edge IDs")
        mask = torch.ones(E, dtype=torch.bool, device=self.device)
        mask[eids.long()] = False
        keep = torch.nonzero(mask).squeeze(1)
        row, col = g.coo()
        self._graphs[etid] = UnitGraph.from_coo(g.num_src, g.num_dst, row[keep], col[keep], g.allowed_formats())
        fr = self._edge_frames[etid].subframe(keep)
        new = Frame({k: fr[k] for k in fr.keys()}, num_rows=int(keep.numel()))
        if store_ids:
            new[EID] = keep.to(self.idtype)
        self._edge_frames[etid] = new
        self._reset_batch()

    def remove_nodes(self, nids, ntype=None, store_ids=False):
        ntid = self.get_ntype_id(ntype)
        nids = _to_tensor(nids, self.idtype, self.device)
        n = self._num_nodes[ntid]
        if nids.numel() and (int(nids.max()) >= n or int(nids.min()) < 0):
            raise DGLError("nids contains invalid node IDs")
        mask = torch.ones(n, dtype=torch.bool, device=self.device)
        mask[nids.long()] = False
        keep = torch.nonzero(mask).squeeze(1)
        relabel = torch.full((n,), -1, dtype=self.idtype, device=self.device)
        relabel[keep] = torch.arange(keep.numel(), dtype=self.idtype, device=self.device)
        for i in range(len(self._canonical_etypes)):
            sid, did = self._src_ntid(i), self._dst_ntid(i)
            if sid != ntid and did != ntid:
                continue
            g = self._graphs[i]
            row, col = g.coo()
            em = torch.ones(row.shape[0], dtype=torch.bool, device=self.device)
            if sid == ntid:
                em &= mask[row.long()]
            if did == ntid:
                em &= mask[col.long()]
            ek = torch.nonzero(em).squeeze(1)
            r, c = row[ek], col[ek]
            if sid == ntid:
                r = relabel[r.long()]
            if did == ntid:
                c = relabel[c.long()]
            ns = keep.numel() if sid == ntid else g.num_src
            nd = keep.numel() if did == ntid else g.num_dst
            self._graphs[i] = UnitGraph.from_coo(ns, nd, r, c, g.allowed_formats())
            fr = self._edge_frames[i].subframe(ek)
            new = Frame({k: fr[k] for k in fr.keys()}, num_rows=int(ek.numel()))
            if store_ids:
                new[EID] = ek.to(self.idtype)
            self._edge_frames[i] = new
        fr = self._node_frames[ntid].subframe(keep)
        new = Frame({k: fr[k] for k in fr.keys()}, num_rows=int(keep.numel()))
        if store_ids:
            new[NID] = keep.to(self.idtype)
        self._node_frames[ntid] = new
        self._num_nodes[ntid] = int(keep.numel())
        self._reset_batch()

    def _reset_batch(self):
        self._batch_num_nodes = None
        self._batch_num_edges = None

    # -------------------------------------------------------------- batching
    @property
    def batch_size(self):
        bn = self.batch_num_nodes(self._ntypes[0])
        return len(bn)

    def batch_num_nodes(self, ntype=None):
        if ntype is None and len(self._ntypes) != 1:
            raise DGLError("Node type name must be specified if there are more than one node types.")
        ntype = self._ntypes[0] if ntype is None else ntype
        if self._batch_num_nodes is None:
            self._batch_num_nodes = {t: torch.tensor([self.num_nodes(t)], dtype=self.idtype, device=self.device)
                                     for t in self._ntypes}
        return self._batch_num_nodes[ntype]

    def set_batch_num_nodes(self, val):
        if not isinstance(val, dict):
            if len(self._ntypes) != 1:
                raise DGLError("Must provide a dictionary when there are multiple node types.")
            val = {self._ntypes[0]: val}
        self._batch_num_nodes = val

    def batch_num_edges(self, etype=None):
        if etype is None and len(self._canonical_etypes) != 1:
            raise DGLError("Edge type name must be specified if there are more than one edge types.")
        c = self.to_canonical_etype(etype)
        if self._batch_num_edges is None:
            self._batch_num_edges = {ce: torch.tensor([self.num_edges(ce)], dtype=self.idtype, device=self.device)
                                     for ce in self._canonical_etypes}
        return self._batch_num_edges[c]

    def set_batch_num_edges(self, val):
        if not isinstance(val, dict):
            if len(self._canonical_etypes) != 1:
                raise DGLError("Must provide a dictionary when there are multiple edge types.")
            val = {self._canonical_etypes[0]: val}
        self._batch_num_edges = val

    # ------------------------------------------------------ message passing
    def apply_nodes(self, func, v=ALL, ntype=None):
        ntid = self.get_ntype_id(ntype)
        nt = self._ntypes[ntid]
        if is_all(v):
            out = core.invoke_node_udf(self, None, nt, func)
            self._set_n_repr(ntid, ALL, out)
        else:
            v = _to_tensor(v, self.idtype, self.device)
            out = core.invoke_node_udf(self, v, nt, func)
            self._set_n_repr(ntid, v, out)

    def apply_edges(self, func, edges=ALL, etype=None):
        etid = self.get_etype_id(etype)
        view = core._RelView(self, etid)
        if is_all(edges):
            if core.is_builtin(func):
                out = core.invoke_gsddmm(view, func)
            else:
                out = core.invoke_edge_udf(view, None, func)
            self._set_e_repr(etid, ALL, out)
        else:
            eid = self._parse_edges(edges, etid)
            if core.is_builtin(func):
                sub = self.edge_subgraph({self._canonical_etypes[etid]: eid}, relabel_nodes=False) \
                    if not self.is_homogeneous else self.edge_subgraph(eid, relabel_nodes=False)
                sview = core._RelView(sub, sub.get_etype_id(self._canonical_etypes[etid]))
                out = core.invoke_gsddmm(sview, func)
            else:
                out = core.invoke_edge_udf(view, eid, func)
            self._set_e_repr(etid, eid, out)

    def update_all(self, message_func, reduce_func, apply_node_func=None, etype=None):
        if etype is None and len(self._canonical_etypes) > 1:
            raise DGLError("Edge type name must be specified if there are more than one edge types "
                           "(or use multi_update_all).")
        etid = self.get_etype_id(etype)
        view = core._RelView(self, etid)
        ndata = core.message_passing(view, message_func, reduce_func, apply_node_func)
        if core.is_builtin(reduce_func) and reduce_func.name in ("min", "max") and self._graphs[etid].num_edges() == 0:
            pass
        self._set_n_repr(self._dst_ntid(etid), ALL, ndata)

    def multi_update_all(self, etype_dict, cross_reducer, apply_node_func=None):
        all_out = defaultdict(list)
        merge_order = defaultdict(list)
        for etype, args in etype_dict.items():
            if len(args) == 2:
                mfunc, rfunc = args
                afunc = None
            elif len(args) == 3:
                mfunc, rfunc, afunc = args
            else:
                raise DGLError('Invalid arguments for edge type "{}". Should be (msg_func, reduce_func, '
                               "[apply_node_func])".format(etype))
            etid = self.get_etype_id(etype)
            dtid = self._dst_ntid(etid)
            view = core._RelView(self, etid)
            all_out[dtid].append(core.message_passing(view, mfunc, rfunc, afunc))
            merge_order[dtid].append(etid)
        for dtid, frames in all_out.items():
            out = reduce_dict_data(frames, cross_reducer, merge_order[dtid])
            self._set_n_repr(dtid, ALL, out)
        if apply_node_func is not None:
            for dtid in all_out:
                self.apply_nodes(apply_node_func, ALL, self._ntypes[dtid])

    def send_and_recv(self, edges, message_func, reduce_func, apply_node_func=None, etype=None):
        etid = self.get_etype_id(etype)
        c = self._canonical_etypes[etid]
        eid = self._parse_edges(edges, etid)
        if eid.numel() == 0:
            return
        sub = self.edge_subgraph({c: eid} if not self.is_homogeneous else eid, relabel_nodes=False)
        sub_etid = sub.get_etype_id(c)
        view = core._RelView(sub, sub_etid)
        ndata = core.message_passing(view, message_func, reduce_func, apply_node_func)
        _, v = self.find_edges(eid, etype=c)
        dst = torch.unique(v)
        self._set_n_repr(self._dst_ntid(etid), dst, {k: t[dst.long()] for k, t in ndata.items()})

    def pull(self, v, message_func, reduce_func, apply_node_func=None, etype=None):
        v = _to_tensor(v, self.idtype, self.device)
        if v.numel() == 0:
            return
        etid = self.get_etype_id(etype)
        c = self._canonical_etypes[etid]
        eid = self.in_edges(v, form="eid", etype=c)
        if eid.numel() == 0:
            return
        sub = self.edge_subgraph({c: eid} if not self.is_homogeneous else eid, relabel_nodes=False)
        view = core._RelView(sub, sub.get_etype_id(c))
        ndata = core.message_passing(view, message_func, reduce_func, apply_node_func)
        dst = torch.unique(v)
        self._set_n_repr(self._dst_ntid(etid), dst, {k: t[dst.long()] for k, t in ndata.items()})

    def push(self, u, message_func, reduce_func, apply_node_func=None, etype=None):
        u = _to_tensor(u, self.idtype, self.device)
        eid = self.out_edges(u, form="eid", etype=etype)
        self.send_and_recv(eid, message_func, reduce_func, apply_node_func, etype=etype)

    def prop_nodes(self, nodes_generator, message_func, reduce_func, apply_node_func=None, etype=None):
        for nodes in nodes_generator:
            self.pull(nodes, message_func, reduce_func, apply_node_func, etype=etype)

    def prop_edges(self, edges_generator, message_func, reduce_func, apply_node_func=None, etype=None):
        for edges in edges_generator:
            self.send_and_recv(edges, message_func, reduce_func, apply_node_func, etype=etype)

    def filter_nodes(self, predicate, nodes=ALL, ntype=None):
        ntid = self.get_ntype_id(ntype)
        nt = self._ntypes[ntid]
        if is_all(nodes):
            mask = core.invoke_node_udf(self, None, nt, predicate)
            return torch.nonzero(mask).squeeze(1).to(self.idtype)
        v = _to_tensor(nodes, self.idtype, self.device)
        mask = core.invoke_node_udf(self, v, nt, predicate)
        return v[mask]

    def filter_edges(self, predicate, edges=ALL, etype=None):
        etid = self.get_etype_id(etype)
        view = core._RelView(self, etid)

        def wrap(eb):
            return {"_mask": predicate(eb)}

        if is_all(edges):
            mask = core.invoke_edge_udf(view, None, wrap)["_mask"]
            return torch.nonzero(mask).squeeze(1).to(self.idtype)
        eid = self._parse_edges(edges, etid)
        mask = core.invoke_edge_udf(view, eid, wrap)["_mask"]
        return eid[mask]

    # -------------------------------------------------------- device / dtype
    def _clone_with(self, graphs, node_frames, edge_frames):
        g = self.__class__.__new__(self.__class__)
        g._init(graphs, None if self._srctypes is not None else list(self._ntypes), list(self._canonical_etypes),
                list(self._num_nodes), node_frames, edge_frames, self._srctypes, self._dsttypes)
        g._batch_num_nodes = self._batch_num_nodes
        g._batch_num_edges = self._batch_num_edges
        return g

    def to(self, device, **kwargs):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device == self.device:
            return self
        graphs = [g.to(device, kwargs.get("non_blocking", False)) for g in self._graphs]
        nf = [f.to(device, **kwargs) for f in self._node_frames]
        ef = [f.to(device, **kwargs) for f in self._edge_frames]
        g = self._clone_with(graphs, nf, ef)
        if self._batch_num_nodes is not None:
            g._batch_num_nodes = {k: v.to(device) for k, v in self._batch_num_nodes.items()}
        if self._batch_num_edges is not None:
            g._batch_num_edges = {k: v.to(device) for k, v in self._batch_num_edges.items()}
        return g

    def cpu(self):
        return self.to(torch.device("cpu"))

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def pin_memory_(self):
        if self.device.type != "cpu":
            raise DGLError("The graph structure must be on CPU to be pinned.")
        if not self._graphs_formats_created():
            dgl_warning("pin_memory_() pins the sparse formats created so far; call create_formats_() first to pin "
                        "all of them.")
        for g in self._graphs:
            g.pin_memory_()
        for f in self._node_frames + self._edge_frames:
            f.pin_memory_()
        return self

    def unpin_memory_(self):
        for g in self._graphs:
            g.unpin_memory_()
        for f in self._node_frames + self._edge_frames:
            f.unpin_memory_()
        return self

    def is_pinned(self):
        return bool(self._graphs) and all(g.is_pinned() for g in self._graphs)

    def record_stream(self, stream):
        for g in self._graphs:
            g.record_stream(stream)
        for f in self._node_frames + self._edge_frames:
            f.record_stream(stream)

    def _graphs_formats_created(self):
        return all(set(g.created_formats()) == set(g.allowed_formats()) for g in self._graphs)

    def clone(self):
        g = self._clone_with(list(self._graphs), [f.clone() for f in self._node_frames],
                             [f.clone() for f in self._edge_frames])
        return g

    def local_var(self):
        return self.clone()

    def dst_sorted_view(self):
        """(graph, eid_map) for single-relation graphs: same nodes / node frames, edges re-numbered in CSC order (see
        ``UnitGraph.dst_sorted_view``); edge frames of the view start empty. Used by attention layers so that scores,
        softmax and the weighted aggregation all run in one edge order without permuting [E, heads] tensors."""
        if len(self._graphs) != 1:
            raise DGLError("dst_sorted_view needs a graph with a single relation")
        ug, emap = self._graphs[0].dst_sorted_view()
        if emap is None:
            return self, None
        from .frame import Frame
        g = self._clone_with([ug], self._node_frames, [Frame(num_rows=ug.num_edges())])
        return g, emap

    @contextmanager
    def local_scope(self):
        old_n = self._node_frames
        old_e = self._edge_frames
        self._node_frames = [f.clone() for f in old_n]
        self._edge_frames = [f.clone() for f in old_e]
        try:
            yield
        finally:
            self._node_frames = old_n
            self._edge_frames = old_e

    def formats(self, formats=None):
        if formats is None:
            created = set(ALL_FORMATS)
            allowed = set(ALL_FORMATS)
            for g in self._graphs:
                created &= set(g.created_formats())
                allowed &= set(g.allowed_formats())
            return {"created": [f for f in ALL_FORMATS if f in created],
                    "not created": [f for f in ALL_FORMATS if f in allowed and f not in created]}
        if isinstance(formats, str):
            formats = [formats]
        graphs = [g.with_formats(formats) for g in self._graphs]
        return self._clone_with(graphs, self._node_frames, self._edge_frames)

    def create_formats_(self):
        for g in self._graphs:
            g.create_formats_()
        return self

    def astype(self, idtype):
        if idtype is None or idtype == self.idtype:
            return self
        if idtype not in (torch.int32, torch.int64):
            raise DGLError("ID type must be int32 or int64")
        graphs = [g.astype(idtype) for g in self._graphs]
        g = self._clone_with(graphs, self._node_frames, self._edge_frames)
        return g

    def long(self):
        return self.astype(torch.int64)

    def int(self):
        return self.astype(torch.int32)

    def adjacency_matrix(self, etype=None):
        """Alias of :meth:`adj` (reference heterograph.py:3759)."""
        return self.adj(etype)

    def get_node_storage(self, key, ntype=None):
        """Storage object (column) of node feature ``key`` (reference heterograph.py:1769): supports ``fetch``."""
        return self._node_frames[self.get_ntype_id(ntype)]._columns[key]

    def get_edge_storage(self, key, etype=None):
        """Storage object (column) of edge feature ``key`` (reference heterograph.py:1773)."""
        return self._edge_frames[self.get_etype_id(etype)]._columns[key]

    def materialize_data(self):
        """Resolve every lazily indexed feature column on the current device (reference heterograph.py:5758)."""
        for frame in list(self._node_frames) + list(self._edge_frames):
            for col in frame._columns.values():
                col.data  # noqa: B018 - touching the property materialises the column
        return self

    def shared_memory(self, name, formats=("coo", "csr", "csc")):
        from .utils.shared_mem import graph_to_shared_memory
        return graph_to_shared_memory(self, name, formats)

    # ------------------------------------------------------------- pickling
    def __getstate__(self):
        return {"graphs": [(g.num_src, g.num_dst, g.coo() if "coo" in g.allowed_formats() else g._coo_temp(),
                            g.allowed_formats()) for g in self._graphs],
                "ntypes": self._ntypes if self._srctypes is None else None,
                "srctypes": self._srctypes, "dsttypes": self._dsttypes, "cet": self._canonical_etypes,
                "num_nodes": self._num_nodes,
                "nframes": [{k: f[k] for k in f.keys()} for f in self._node_frames],
                "eframes": [{k: f[k] for k in f.keys()} for f in self._edge_frames],
                "bnn": self._batch_num_nodes, "bne": self._batch_num_edges}

    def __setstate__(self, st):
        graphs = [UnitGraph.from_coo(ns, nd, coo[0], coo[1], fm) for ns, nd, coo, fm in st["graphs"]]
        nf = [Frame(d, num_rows=n) for d, n in zip(st["nframes"], st["num_nodes"])]
        ef = [Frame(d, num_rows=g.num_edges()) for d, g in zip(st["eframes"], graphs)]
        self._init(graphs, st["ntypes"], st["cet"], st["num_nodes"], nf, ef, st["srctypes"], st["dsttypes"])
        self._batch_num_nodes = st["bnn"]
        self._batch_num_edges = st["bne"]

    def __copy__(self):
        return self.clone()

    # ----------------------------------------------------------------- repr
    def __repr__(self):
        if self.is_block:
            return (f"Block(num_src_nodes={_fmt(self._srctypes, self._num_nodes[:len(self._srctypes)])}, "
                    f"num_dst_nodes={_fmt(self._dsttypes, self._num_nodes[len(self._srctypes):])}, "
                    f"num_edges={_fmt_e(self)})")
        if self.is_homogeneous:
            return (f"Graph(num_nodes={self.num_nodes()}, num_edges={self.num_edges()},\n"
                    f"      ndata_schemes={self.node_attr_schemes()}\n"
                    f"      edata_schemes={self.edge_attr_schemes()})")
        nn = {t: n for t, n in zip(self._ntypes, self._num_nodes)}
        ne = {c: g.num_edges() for c, g in zip(self._canonical_etypes, self._graphs)}
        return f"Graph(num_nodes={nn},\n      num_edges={ne},\n      metagraph={[(s, d, e) for s, e, d in self._canonical_etypes]})"


def _fmt(types, nums):
    if len(types) == 1:
        return nums[0]
    return {t: n for t, n in zip(types, nums)}


def _fmt_e(g):
    if len(g._canonical_etypes) == 1:
        return g._graphs[0].num_edges()
    return {c: gg.num_edges() for c, gg in zip(g._canonical_etypes, g._graphs)}


def _form(src, dst, eid, form):
    if form == "all":
        return src, dst, eid
    if form == "uv":
        return src, dst
    if form == "eid":
        return eid
    raise DGLError(f'Invalid form: {form}. Must be "all", "uv" or "eid".')


def combine_names(names, ids=None):
    if ids is None:
        return "+".join(sorted(names))
    selected = sorted([names[i] for i in ids])
    return "+".join(selected)


def make_canonical_etypes(etypes, ntypes, metagraph):
    """[(srctype, etype, dsttype)] from parallel type-name lists and a metagraph whose edge i connects the node-type
    ids of relation i (reference heterograph.py:6407)."""
    if len(etypes) != metagraph.num_edges():
        raise DGLError(f"Length of edge type list must match the number of edges in the metagraph. "
                       f"{len(etypes)} vs {metagraph.num_edges()}")
    if len(ntypes) != metagraph.num_nodes():
        raise DGLError(f"Length of nodes type list must match the number of nodes in the metagraph. "
                       f"{len(ntypes)} vs {metagraph.num_nodes()}")
    if len(etypes) == 1 and len(ntypes) == 1:
        return [(ntypes[0], etypes[0], ntypes[0])]
    out = []
    for i, name in enumerate(etypes):
        s, d = metagraph.find_edge(i)
        out.append((ntypes[int(s)], name, ntypes[int(d)]))
    return out


def find_src_dst_ntypes(ntypes, metagraph):
    """({src type: id}, {dst type: id}) when no node type is both a source and a destination (uni-bipartite
    metagraph), else ``None``. Isolated types count as sources (reference heterograph.py:6448)."""
    src, dst = set(), set()
    for i in range(metagraph.num_edges()):
        s, d = metagraph.find_edge(i)
        src.add(int(s))
        dst.add(int(d))
    if src & dst:
        return None
    src |= set(range(metagraph.num_nodes())) - dst
    return {ntypes[t]: t for t in sorted(src)}, {ntypes[t]: t for t in sorted(dst)}


def pad_tuple(tup, length, pad_val=None):
    """Tuple of exactly ``length`` (non-tuples are wrapped first); ``None`` when it is already longer."""
    if not isinstance(tup, tuple):
        tup = (tup,)
    if len(tup) > length:
        return None
    return tup + (pad_val,) * (length - len(tup))


def combine_frames(frames, ids, col_names=None):
    """Concatenate ``frames[i] for i in ids`` over their common columns (optionally restricted to ``col_names``);
    columns whose schemes disagree raise, no common column gives ``None`` (reference heterograph.py:6561)."""
    from .frame import Frame
    schemes = None
    for i in ids:
        fr = frames[i]
        if fr.num_rows == 0:
            continue
        if schemes is None:
            schemes = dict(fr.schemes) if col_names is None else {k: fr.schemes[k] for k in col_names}
            continue
        for key, scheme in list(schemes.items()):
            if key not in fr.schemes:
                del schemes[key]
            elif fr.schemes[key] != scheme:
                raise DGLError(f"Cannot concatenate column {key} with shape {scheme} and shape {fr.schemes[key]}")
    if not schemes:
        return None
    cols = {k: torch.cat([frames[i][k] for i in ids if frames[i].num_rows > 0], 0) for k in schemes}
    return Frame(cols)


def reduce_dict_data(frames, reducer, order=None):
    """Merge per-relation results for one dst type (reference heterograph.py:6500+)."""
    if len(frames) == 1 and reducer != "stack":
        return frames[0]
    if callable(reducer):
        merger = reducer
    elif reducer == "stack":
        def merger(flist):
            return torch.stack(flist, 1)
    else:
        red = {"sum": torch.sum, "max": lambda x, dim: torch.max(x, dim)[0], "min": lambda x, dim: torch.min(x, dim)[0],
               "mean": torch.mean}.get(reducer)
        if red is None:
            raise DGLError('Invalid cross type reducer. Must be one of "sum", "max", "min", "mean" or "stack".')

        def merger(flist):
            return red(torch.stack(flist, 0), dim=0) if len(flist) > 1 else flist[0]

    keys = set()
    for f in frames:
        keys.update(f.keys())
    ret = {}
    for k in keys:
        flist = [f[k] for f in frames if k in f]
        ret[k] = merger(flist)
    return ret


class DGLBlock(DGLGraph):
    """Message-flow graph: bipartite src -> dst per layer; dst nodes are a prefix
    of src nodes (reference heterograph.py:6631, transforms/to_block.py)."""

    is_block = True


DGLHeteroGraph = DGLGraph
