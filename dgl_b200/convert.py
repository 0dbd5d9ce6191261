"""Graph construction / conversion.

Parity: /root/reference/python/dgl/convert.py (graph :41, heterograph :207,
create_block :353, to_heterogeneous :642, to_homogeneous :846, from_scipy :1007,
bipartite_from_scipy, from_networkx :1219, to_networkx) and generators.rand_graph.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch

from .base import EID, ETYPE, NID, NTYPE, DGLError, dgl_warning
from .frame import Frame
from .graph_index import UnitGraph
from .heterograph import DGLBlock, DGLGraph, combine_names

__all__ = ["graph", "heterograph", "create_block", "block_to_graph", "to_heterogeneous", "to_homogeneous",
           "from_scipy", "bipartite_from_scipy", "hetero_from_shared_memory", "to_cugraph", "from_cugraph", "from_networkx", "bipartite_from_networkx", "to_networkx",
           "rand_graph", "rand_bipartite", "from_cugraph", "to_cugraph"]


def _parse_data(data, idtype, device):
    """Accept (u, v) tensors/lists, ('coo', (u,v)), ('csr', (indptr, indices, eids)), ('csc', ...),
    scipy matrices and networkx graphs."""
    if isinstance(data, tuple) and len(data) == 2 and isinstance(data[0], str):
        fmt, arrs = data
        return fmt, arrs
    if isinstance(data, tuple) and len(data) == 2:
        return "coo", data
    try:
        import scipy.sparse as sp
        if sp.issparse(data):
            m = data.tocoo()
            return "coo", (torch.from_numpy(m.row.astype(np.int64)), torch.from_numpy(m.col.astype(np.int64)))
    except ImportError:
        pass
    try:
        import networkx as nx
        if isinstance(data, nx.Graph):
            g = data if data.is_directed() else data.to_directed()
            nodes = {n: i for i, n in enumerate(g.nodes())} if not all(isinstance(n, int) for n in g.nodes()) else None
            u = [nodes[a] if nodes else a for a, _ in g.edges()]
            v = [nodes[b] if nodes else b for _, b in g.edges()]
            return "coo_n", ((torch.tensor(u, dtype=torch.int64), torch.tensor(v, dtype=torch.int64)),
                             g.number_of_nodes())
    except ImportError:
        pass
    if isinstance(data, list):
        if len(data) == 0:
            return "coo", (torch.empty(0, dtype=torch.int64), torch.empty(0, dtype=torch.int64))
        arr = torch.as_tensor(data)
        if arr.dim() == 2 and arr.shape[1] == 2:
            return "coo", (arr[:, 0], arr[:, 1])
    raise DGLError(f"Unsupported graph data type: {type(data)}")


def _tensorize(x, idtype, device):
    if torch.is_tensor(x):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x)) if not isinstance(x, (list, tuple)) else torch.as_tensor(x)
    if t.numel() == 0:
        t = t.to(torch.int64)
    if t.dtype not in (torch.int32, torch.int64):
        if t.dtype in (torch.int8, torch.int16, torch.uint8):
            t = t.to(torch.int64)
        else:
            raise DGLError(f"Expect the node ID tensor to be int32 or int64, got {t.dtype}")
    if idtype is not None:
        t = t.to(idtype)
    if device is not None:
        t = t.to(device)
    return t.reshape(-1)


def _make_unit(data, nsrc, ndst, idtype, device, same_type):
    fmt, arrs = _parse_data(data, idtype, device)
    if fmt == "coo_n":
        arrs, n = arrs
        fmt = "coo"
        if nsrc is None:
            nsrc = ndst = n
    if fmt == "coo":
        u = _tensorize(arrs[0], idtype, device)
        v = _tensorize(arrs[1], idtype, device)
        if u.dtype != v.dtype:
            v = v.to(u.dtype)
        if u.device != v.device:
            raise DGLError("src and dst tensors must be on the same device")
        if u.shape[0] != v.shape[0]:
            raise DGLError(f"src and dst have different lengths: {u.shape[0]} vs {v.shape[0]}")
        if u.numel() and (int(u.min()) < 0 or int(v.min()) < 0):
            raise DGLError("node IDs must be non-negative")
        ms = int(u.max()) + 1 if u.numel() else 0
        md = int(v.max()) + 1 if v.numel() else 0
        if same_type:
            m = max(ms, md)
            if nsrc is None:
                nsrc = ndst = m
            elif nsrc < m:
                raise DGLError(f"The num_nodes argument must be larger than the max ID in the data, "
                               f"but got {nsrc} and {m - 1}.")
        else:
            if nsrc is None:
                nsrc = ms
            if ndst is None:
                ndst = md
            if nsrc < ms or ndst < md:
                raise DGLError("num_nodes is smaller than the max node ID in the data")
        return UnitGraph.from_coo(nsrc, ndst, u, v)
    if fmt in ("csr", "csc"):
        indptr = _tensorize(arrs[0], idtype, device)
        indices = _tensorize(arrs[1], idtype, device).to(indptr.dtype)
        eids = arrs[2] if len(arrs) > 2 else None
        if eids is not None and len(eids) > 0:
            eids = _tensorize(eids, indptr.dtype, device)
        else:
            eids = None
        nrow = indptr.shape[0] - 1
        ncol_min = int(indices.max()) + 1 if indices.numel() else 0
        if fmt == "csr":
            ns = nrow if nsrc is None else nsrc
            nd = (max(ncol_min, nrow) if same_type else ncol_min) if ndst is None else ndst
            if same_type:
                ns = nd = max(ns, nd)
                if ns > nrow:
                    indptr = torch.cat([indptr, indptr[-1:].expand(ns - nrow)])
            return UnitGraph.from_csr(ns, nd, indptr, indices, eids)
        nd = nrow if ndst is None else ndst
        ns = (max(ncol_min, nrow) if same_type else ncol_min) if nsrc is None else nsrc
        if same_type:
            ns = nd = max(ns, nd)
            if nd > nrow:
                indptr = torch.cat([indptr, indptr[-1:].expand(nd - nrow)])
        return UnitGraph.from_csc(ns, nd, indptr, indices, eids)
    raise DGLError(f"unknown sparse format {fmt}")


def graph(data, *, num_nodes=None, idtype=None, device=None, row_sorted=False, col_sorted=False):
    """Create a homogeneous graph (reference convert.py:41)."""
    ug = _make_unit(data, num_nodes, num_nodes, idtype, device, True)
    return DGLGraph([ug], ["_N"], [("_N", "_E", "_N")], [ug.num_src])


def create_from_edges(sparse_fmt, arrays, utype, etype, vtype, urange, vrange, row_sorted=False, col_sorted=False):
    """One-relation graph from typed incident arrays (reference convert.py:1990). ``utype == vtype`` gives a
    homogeneous-style relation over one node set, otherwise a bipartite one. ``row_sorted`` / ``col_sorted`` are
    accepted for parity: sortedness is detected when a compressed format is first built (graph_index.coo_to_csr)."""
    same = utype == vtype
    ug = _make_unit((sparse_fmt, tuple(arrays)), urange, vrange, None, None, same)
    if same:
        return DGLGraph([ug], [utype], [(utype, etype, vtype)], [ug.num_src])
    return DGLGraph([ug], [utype, vtype], [(utype, etype, vtype)], [ug.num_src, ug.num_dst])


def heterograph(data_dict, num_nodes_dict=None, idtype=None, device=None):
    """Create a heterogeneous graph from {(stype, etype, dtype): data} (reference convert.py:207)."""
    if not data_dict:
        raise DGLError("data_dict must not be empty")
    need_infer = num_nodes_dict is None
    units = {}
    inferred = defaultdict(int)
    for (s, e, d), data in data_dict.items():
        ug = _make_unit(data, None if need_infer else num_nodes_dict.get(s), None if need_infer else num_nodes_dict.get(d),
                        idtype, device, False)
        if idtype is None:
            idtype = ug.idtype
        units[(s, e, d)] = ug
        inferred[s] = max(inferred[s], ug.num_src)
        inferred[d] = max(inferred[d], ug.num_dst)
    if need_infer:
        num_nodes_dict = dict(inferred)
    else:
        for t, n in inferred.items():
            if t not in num_nodes_dict:
                raise DGLError(f"Missing node type {t} for num_nodes_dict.")
    ntypes = sorted(num_nodes_dict.keys())
    cets = sorted(units.keys())
    graphs = []
    for c in cets:
        ug = units[c]
        ns, nd = num_nodes_dict[c[0]], num_nodes_dict[c[2]]
        if ug.num_src != ns or ug.num_dst != nd:
            row, col = ug.coo()
            ug = UnitGraph.from_coo(ns, nd, row, col)
        if ug.idtype != idtype:
            ug = ug.astype(idtype)
        graphs.append(ug)
    return DGLGraph(graphs, ntypes, cets, [num_nodes_dict[t] for t in ntypes])


def create_block(data_dict, num_src_nodes=None, num_dst_nodes=None, idtype=None, device=None):
    """Create a message-flow graph (reference convert.py:353)."""
    if not isinstance(data_dict, dict):
        data_dict = {("_N", "_E", "_N"): data_dict}
        if num_src_nodes is not None and not isinstance(num_src_nodes, dict):
            num_src_nodes = {"_N": num_src_nodes}
        if num_dst_nodes is not None and not isinstance(num_dst_nodes, dict):
            num_dst_nodes = {"_N": num_dst_nodes}
    units = {}
    inf_s, inf_d = defaultdict(int), defaultdict(int)
    for (s, e, d), data in data_dict.items():
        ug = _make_unit(data, None if num_src_nodes is None else num_src_nodes.get(s),
                        None if num_dst_nodes is None else num_dst_nodes.get(d), idtype, device, False)
        if idtype is None:
            idtype = ug.idtype
        units[(s, e, d)] = ug
        inf_s[s] = max(inf_s[s], ug.num_src)
        inf_d[d] = max(inf_d[d], ug.num_dst)
    if num_src_nodes is None:
        num_src_nodes = dict(inf_s)
    if num_dst_nodes is None:
        num_dst_nodes = dict(inf_d)
    srctypes = sorted(num_src_nodes.keys())
    dsttypes = sorted(num_dst_nodes.keys())
    cets = sorted(units.keys())
    graphs = []
    for c in cets:
        ug = units[c]
        ns, nd = num_src_nodes[c[0]], num_dst_nodes[c[2]]
        if ug.num_src != ns or ug.num_dst != nd:
            row, col = ug.coo()
            ug = UnitGraph.from_coo(ns, nd, row, col)
        graphs.append(ug.astype(idtype))
    return DGLBlock(graphs, None, cets, [num_src_nodes[t] for t in srctypes] + [num_dst_nodes[t] for t in dsttypes],
                    srctypes=srctypes, dsttypes=dsttypes)


def block_to_graph(block):
    """Convert a block to a graph with '_src'/'_dst'-suffixed node types."""
    data = {}
    for i, (s, e, d) in enumerate(block.canonical_etypes):
        data[(s + "_src", e, d + "_dst")] = block._graphs[i].coo()
    nn = {t + "_src": block.num_src_nodes(t) for t in block.srctypes}
    nn.update({t + "_dst": block.num_dst_nodes(t) for t in block.dsttypes})
    g = heterograph(data, nn, idtype=block.idtype, device=block.device)
    for t in block.srctypes:
        for k, v in block.srcnodes[t].data.items():
            g.nodes[t + "_src"].data[k] = v
    for t in block.dsttypes:
        for k, v in block.dstnodes[t].data.items():
            g.nodes[t + "_dst"].data[k] = v
    for (s, e, d) in block.canonical_etypes:
        for k, v in block.edges[(s, e, d)].data.items():
            g.edges[(s + "_src", e, d + "_dst")].data[k] = v
    return g


def to_homogeneous(G, ndata=None, edata=None, store_type=True, return_count=False):
    """Merge all node/edge types into one (reference convert.py:846). Nodes of the same type
    get consecutive ids; NTYPE/NID and ETYPE/EID record the origin."""
    dev, idt = G.device, G.idtype
    offsets, off = {}, 0
    for t in G.ntypes:
        offsets[t] = off
        off += G.num_nodes(t)
    total = off
    rows, cols, etype_ids, eids = [], [], [], []
    ecount = []
    for i, (s, e, d) in enumerate(G.canonical_etypes):
        r, c = G._graphs[i].coo()
        rows.append(r + offsets[s]); cols.append(c + offsets[d])
        ecount.append(r.shape[0])
        if store_type:
            etype_ids.append(torch.full((r.shape[0],), i, dtype=idt, device=dev))
        eids.append(torch.arange(r.shape[0], dtype=idt, device=dev))
    ug = UnitGraph.from_coo(total, total, torch.cat(rows) if rows else torch.empty(0, dtype=idt),
                            torch.cat(cols) if cols else torch.empty(0, dtype=idt))
    g = DGLGraph([ug], ["_N"], [("_N", "_E", "_N")], [total])
    ncount = [G.num_nodes(t) for t in G.ntypes]
    if store_type:
        g.ndata[NTYPE] = torch.cat([torch.full((n,), i, dtype=idt, device=dev) for i, n in enumerate(ncount)])
        g.edata[ETYPE] = torch.cat(etype_ids) if etype_ids else torch.empty(0, dtype=idt, device=dev)
    g.ndata[NID] = torch.cat([torch.arange(n, dtype=idt, device=dev) for n in ncount])
    g.edata[EID] = torch.cat(eids) if eids else torch.empty(0, dtype=idt, device=dev)
    for key in (ndata or []):
        g.ndata[key] = torch.cat([G.nodes[t].data[key] for t in G.ntypes], 0)
    for key in (edata or []):
        g.edata[key] = torch.cat([G.edges[c].data[key] for c in G.canonical_etypes], 0)
    if return_count:
        return g, ncount, ecount
    return g


def to_heterogeneous(G, ntypes, etypes, ntype_field=NTYPE, etype_field=ETYPE, metagraph=None):
    """Split a homogeneous graph by node/edge type ids (reference convert.py:642)."""
    if not G.is_homogeneous:
        raise DGLError("The input graph should be homogeneous.")
    idt, dev = G.idtype, G.device
    ntype_ids = G.ndata[ntype_field].long()
    etype_ids = G.edata[etype_field].long()
    src, dst = G.all_edges(order="eid")
    ntype_count = torch.bincount(ntype_ids, minlength=len(ntypes))
    # local ids: stable rank within type
    order = torch.sort(ntype_ids, stable=True)[1]
    starts = torch.cumsum(ntype_count, 0) - ntype_count
    local = torch.empty_like(ntype_ids)
    local[order] = torch.arange(ntype_ids.shape[0], device=dev) - torch.repeat_interleave(starts, ntype_count)
    node_groups = [order[starts[i]: starts[i] + ntype_count[i]] for i in range(len(ntypes))]
    srctype = ntype_ids[src.long()]
    dsttype = ntype_ids[dst.long()]
    data_dict, edge_groups, cets = {}, {}, []
    for eid_t, ename in enumerate(etypes):
        sel = torch.nonzero(etype_ids == eid_t).squeeze(1)
        if sel.numel() == 0:
            continue
        pairs = torch.unique(torch.stack([srctype[sel], dsttype[sel]], 1), dim=0)
        for st, dt in pairs.tolist():
            m = sel[(srctype[sel] == st) & (dsttype[sel] == dt)]
            c = (ntypes[st], ename, ntypes[dt])
            data_dict[c] = (local[src[m].long()].to(idt), local[dst[m].long()].to(idt))
            edge_groups[c] = m
    hg = heterograph(data_dict, {t: int(ntype_count[i]) for i, t in enumerate(ntypes)}, idtype=idt, device=dev)
    for i, t in enumerate(ntypes):
        hg.nodes[t].data[NID] = node_groups[i].to(idt)
        for k, v in G.ndata.items():
            if k in (ntype_field, NID):
                continue
            hg.nodes[t].data[k] = v[node_groups[i]]
    for c in hg.canonical_etypes:
        m = edge_groups[c]
        hg.edges[c].data[EID] = m.to(idt)
        for k, v in G.edata.items():
            if k in (etype_field, EID):
                continue
            hg.edges[c].data[k] = v[m]
    return hg


def from_scipy(sp_mat, eweight_name=None, idtype=None, device=None):
    m = sp_mat.tocoo()
    if m.shape[0] != m.shape[1]:
        raise DGLError(f"Expect the number of rows to be the same as the number of columns for sp_mat, got {m.shape}.")
    g = graph((torch.from_numpy(m.row.astype(np.int64)), torch.from_numpy(m.col.astype(np.int64))),
              num_nodes=m.shape[0], idtype=idtype, device=device)
    if eweight_name is not None:
        g.edata[eweight_name] = torch.from_numpy(m.data).to(g.device)
    return g


def bipartite_from_scipy(sp_mat, utype, etype, vtype, eweight_name=None, idtype=None, device=None):
    m = sp_mat.tocoo()
    g = heterograph({(utype, etype, vtype): (torch.from_numpy(m.row.astype(np.int64)),
                                             torch.from_numpy(m.col.astype(np.int64)))},
                    {utype: m.shape[0], vtype: m.shape[1]}, idtype=idtype, device=device)
    if eweight_name is not None:
        g.edata[eweight_name] = torch.from_numpy(m.data).to(g.device)
    return g


def from_networkx(nx_graph, node_attrs=None, edge_attrs=None, edge_id_attr_name=None, idtype=None, device=None):
    import networkx as nx
    if not nx_graph.is_directed():
        if edge_id_attr_name is not None:
            raise DGLError("edge_id_attr_name is only supported for directed graphs")
        nx_graph = nx_graph.to_directed()
    nodes = list(nx_graph.nodes())
    relabel = {n: i for i, n in enumerate(nodes)}
    edges = list(nx_graph.edges(data=True))
    if edge_id_attr_name is not None:
        edges.sort(key=lambda e: e[2][edge_id_attr_name])
    u = torch.tensor([relabel[a] for a, _, _ in edges], dtype=torch.int64)
    v = torch.tensor([relabel[b] for _, b, _ in edges], dtype=torch.int64)
    g = graph((u, v), num_nodes=len(nodes), idtype=idtype, device=device)
    for k in (node_attrs or []):
        g.ndata[k] = torch.as_tensor(np.stack([np.asarray(nx_graph.nodes[n][k]) for n in nodes])).to(g.device)
    for k in (edge_attrs or []):
        g.edata[k] = torch.as_tensor(np.stack([np.asarray(d[k]) for _, _, d in edges])).to(g.device)
    return g


def bipartite_from_networkx(nx_graph, utype, etype, vtype, u_attrs=None, e_attrs=None, v_attrs=None,
                            edge_id_attr_name=None, idtype=None, device=None):
    top = [n for n, d in nx_graph.nodes(data=True) if d.get("bipartite") == 0]
    bottom = [n for n, d in nx_graph.nodes(data=True) if d.get("bipartite") == 1]
    tm = {n: i for i, n in enumerate(top)}
    bm = {n: i for i, n in enumerate(bottom)}
    edges = [(a, b, d) for a, b, d in nx_graph.edges(data=True) if a in tm and b in bm]
    if edge_id_attr_name is not None:
        edges.sort(key=lambda e: e[2][edge_id_attr_name])
    u = torch.tensor([tm[a] for a, _, _ in edges], dtype=torch.int64)
    v = torch.tensor([bm[b] for _, b, _ in edges], dtype=torch.int64)
    g = heterograph({(utype, etype, vtype): (u, v)}, {utype: len(top), vtype: len(bottom)}, idtype=idtype, device=device)
    for k in (u_attrs or []):
        g.nodes[utype].data[k] = torch.as_tensor(np.stack([np.asarray(nx_graph.nodes[n][k]) for n in top]))
    for k in (v_attrs or []):
        g.nodes[vtype].data[k] = torch.as_tensor(np.stack([np.asarray(nx_graph.nodes[n][k]) for n in bottom]))
    for k in (e_attrs or []):
        g.edata[k] = torch.as_tensor(np.stack([np.asarray(d[k]) for _, _, d in edges]))
    return g


def to_networkx(g, node_attrs=None, edge_attrs=None):
    import networkx as nx
    if g.device.type != "cpu":
        raise DGLError("Cannot convert a CUDA graph to networkx. Call g.cpu() first.")
    if not g.is_homogeneous:
        raise DGLError("dgl.to_networkx only supports homogeneous graphs.")
    src, dst = g.all_edges()
    nxg = nx.MultiDiGraph()
    nxg.add_nodes_from(range(g.num_nodes()))
    for eid, (u, v) in enumerate(zip(src.tolist(), dst.tolist())):
        nxg.add_edge(u, v, id=eid)
    for k in (node_attrs or []):
        for n in range(g.num_nodes()):
            nxg.nodes[n][k] = g.ndata[k][n]
    if edge_attrs:
        for u, v, key, d in nxg.edges(keys=True, data=True):
            for k in edge_attrs:
                d[k] = g.edata[k][d["id"]]
    return nxg


def rand_graph(num_nodes, num_edges, idtype=torch.int64, device=torch.device("cpu")):
    """Random graph with ``num_edges`` uniformly drawn (src, dst) pairs (reference generators.py)."""
    gen_dev = device if torch.device(device).type == "cuda" else "cpu"
    u = torch.randint(0, num_nodes, (num_edges,), device=gen_dev)
    v = torch.randint(0, num_nodes, (num_edges,), device=gen_dev)
    return graph((u, v), num_nodes=num_nodes, idtype=idtype, device=device)


def rand_bipartite(utype, etype, vtype, num_src_nodes, num_dst_nodes, num_edges, idtype=torch.int64,
                   device=torch.device("cpu")):
    u = torch.randint(0, num_src_nodes, (num_edges,))
    v = torch.randint(0, num_dst_nodes, (num_edges,))
    return heterograph({(utype, etype, vtype): (u, v)}, {utype: num_src_nodes, vtype: num_dst_nodes}, idtype=idtype,
                       device=device)


def from_cugraph(cugraph_graph):
    raise DGLError("cugraph is not available in this build")


def to_cugraph(g):
    raise DGLError("cugraph is not available in this build")


def hetero_from_shared_memory(name):
    """Graph structure (no features) previously published with ``g.shared_memory(name)`` (reference convert.py:187)."""
    from .utils.shared_mem import graph_from_shared_memory
    return graph_from_shared_memory(name)


def to_cugraph(g):
    """cuGraph interop (reference convert.py to_cugraph) - needs the optional ``cugraph`` / ``cudf`` packages."""
    try:
        import cudf
        import cugraph
    except ImportError as e:  # pragma: no cover - optional dependency
        raise DGLError("to_cugraph requires the cugraph and cudf packages") from e
    if not g.is_homogeneous:
        raise DGLError("dgl_graph is not Homogeneous")
    u, v = g.edges()
    df = cudf.DataFrame({"source": cudf.from_dlpack(torch.utils.dlpack.to_dlpack(u.contiguous())),
                         "destination": cudf.from_dlpack(torch.utils.dlpack.to_dlpack(v.contiguous()))})
    cg = cugraph.Graph(directed=True)
    cg.from_cudf_edgelist(df)
    return cg


def from_cugraph(cugraph_graph):
    """Create a graph from a cuGraph graph on the same device (reference convert.py from_cugraph)."""
    if not cugraph_graph.is_directed():
        cugraph_graph = cugraph_graph.to_directed()
    edges = cugraph_graph.edges()
    src = torch.utils.dlpack.from_dlpack(edges["src"].to_dlpack())
    dst = torch.utils.dlpack.from_dlpack(edges["dst"].to_dlpack())
    return graph((src, dst))
