"""Subgraph extraction.

Parity: /root/reference/python/dgl/subgraph.py (node_subgraph :38, edge_subgraph :205,
in_subgraph :365, out_subgraph :510, khop_in_subgraph :655, khop_out_subgraph :826,
node_type_subgraph :1000, edge_type_subgraph :1086).
"""
from __future__ import annotations

import torch

from .base import EID, NID, DGLError
from .frame import Frame
from .graph_index import UnitGraph
from .heterograph import DGLGraph, _to_tensor

__all__ = ["node_subgraph", "edge_subgraph", "in_subgraph", "out_subgraph", "khop_in_subgraph", "khop_out_subgraph",
           "node_type_subgraph", "edge_type_subgraph"]


def _frames_from(graph, node_inds, edge_inds, store_ids):
    nf = []
    for i, fr in enumerate(graph._node_frames):
        ind = node_inds[i]
        if ind is None:
            nf.append(fr.clone())
        else:
            sub = fr.subframe(ind)
            f = Frame(num_rows=int(ind.shape[0]))
            f._columns = sub._columns
            if store_ids:
                f[NID] = ind
            nf.append(f)
    ef = []
    for i, fr in enumerate(graph._edge_frames):
        ind = edge_inds[i]
        sub = fr.subframe(ind)
        f = Frame(num_rows=int(ind.shape[0]))
        f._columns = sub._columns
        if store_ids:
            f[EID] = ind
        ef.append(f)
    return nf, ef


def _new_like(graph, graphs, num_nodes, nf, ef):
    g = DGLGraph.__new__(DGLGraph)
    g._init(graphs, None if graph._srctypes is not None else list(graph._ntypes), list(graph._canonical_etypes),
            num_nodes, nf, ef, graph._srctypes, graph._dsttypes)
    return g


def _node_dict(graph, nodes, what="nodes"):
    if not isinstance(nodes, dict):
        if len(graph.ntypes) > 1:
            raise DGLError(f"need a dict of node type and IDs for graph with multiple node types")
        nodes = {graph.ntypes[0]: nodes}
    return nodes


def node_subgraph(graph, nodes, *, relabel_nodes=True, store_ids=True, output_device=None):
    """Induced subgraph on ``nodes`` (tensor / bool mask / dict per node type)."""
    if graph.is_block:
        raise DGLError("Extracting subgraph from a block graph is not allowed.")
    nodes = _node_dict(graph, nodes)
    idt, dev = graph.idtype, graph.device
    inds, masks, relabels = [], [], []
    for ntid, nt in enumerate(graph.ntypes):
        n = graph._num_nodes[ntid]
        if nt in nodes:
            v = nodes[nt]
            v = _to_tensor(v, idt, dev)
        else:
            v = torch.empty(0, dtype=idt, device=dev)
        m = torch.zeros(n, dtype=torch.bool, device=dev)
        m[v.long()] = True
        masks.append(m)
        inds.append(v)
        r = torch.full((n,), -1, dtype=idt, device=dev)
        r[v.long()] = torch.arange(v.shape[0], dtype=idt, device=dev)
        relabels.append(r)
    graphs, einds = [], []
    for i in range(len(graph._canonical_etypes)):
        sid, did = graph._src_ntid(i), graph._dst_ntid(i)
        g = graph._graphs[i]
        row, col = g.coo()
        if dev.type == "cpu" and g._csr is not None and inds[sid].numel() * 4 < g.num_src:
            # small node set on a graph whose CSR exists: native sub-matrix slice (csrc/cpu/spmat.cc), cost proportional
            # to the degrees of the kept rows instead of a pass over every edge; edge-id order as in the masked path
            from .ops import cpu_struct
            indptr, indices, eids = g._csr
            _, _, sd = cpu_struct.slice_matrix(indptr, indices, eids, inds[sid], relabels[did])
            ek = torch.sort(sd.long())[0]
        else:
            em = masks[sid][row.long()] & masks[did][col.long()]
            ek = torch.nonzero(em).squeeze(1)
        r, c = row[ek], col[ek]
        if relabel_nodes:
            r, c = relabels[sid][r.long()], relabels[did][c.long()]
            graphs.append(UnitGraph.from_coo(inds[sid].shape[0], inds[did].shape[0], r, c))
        else:
            graphs.append(UnitGraph.from_coo(g.num_src, g.num_dst, r, c))
        einds.append(ek.to(idt))
    if relabel_nodes:
        nf, ef = _frames_from(graph, inds, einds, store_ids)
        sg = _new_like(graph, graphs, [int(v.shape[0]) for v in inds], nf, ef)
    else:
        nf, ef = _frames_from(graph, [None] * len(inds), einds, store_ids)
        sg = _new_like(graph, graphs, list(graph._num_nodes), nf, ef)
    return sg if output_device is None else sg.to(output_device)


def edge_subgraph(graph, edges, *, relabel_nodes=True, store_ids=True, output_device=None):
    """Subgraph induced by ``edges`` (tensor / bool mask / dict per edge type)."""
    if graph.is_block and relabel_nodes:
        raise DGLError("Extracting subgraph from a block graph is not allowed.")
    if not isinstance(edges, dict):
        if len(graph.canonical_etypes) > 1:
            raise DGLError("need a dict of edge type and IDs for graph with multiple edge types")
        edges = {graph.canonical_etypes[0]: edges}
    edges = {graph.to_canonical_etype(k): v for k, v in edges.items()}
    idt, dev = graph.idtype, graph.device
    einds = []
    for i, c in enumerate(graph._canonical_etypes):
        if c in edges:
            einds.append(_to_tensor(edges[c], idt, dev))
        else:
            einds.append(torch.empty(0, dtype=idt, device=dev))
    nnt = len(graph._num_nodes)
    if relabel_nodes:
        used = [[] for _ in range(nnt)]
        for i in range(len(graph._canonical_etypes)):
            row, col = graph._graphs[i].coo()
            used[graph._src_ntid(i)].append(row[einds[i].long()])
            used[graph._dst_ntid(i)].append(col[einds[i].long()])
        inds, relabels = [], []
        for ntid in range(nnt):
            n = graph._num_nodes[ntid]
            u = torch.unique(torch.cat(used[ntid])) if used[ntid] else torch.empty(0, dtype=idt, device=dev)
            inds.append(u.to(idt))
            r = torch.full((n,), -1, dtype=idt, device=dev)
            r[u.long()] = torch.arange(u.shape[0], dtype=idt, device=dev)
            relabels.append(r)
    graphs = []
    for i in range(len(graph._canonical_etypes)):
        g = graph._graphs[i]
        row, col = g.coo()
        r, c = row[einds[i].long()], col[einds[i].long()]
        if relabel_nodes:
            sid, did = graph._src_ntid(i), graph._dst_ntid(i)
            graphs.append(UnitGraph.from_coo(inds[sid].shape[0], inds[did].shape[0], relabels[sid][r.long()],
                                             relabels[did][c.long()]))
        else:
            graphs.append(UnitGraph.from_coo(g.num_src, g.num_dst, r, c))
    if relabel_nodes:
        nf, ef = _frames_from(graph, inds, einds, store_ids)
        sg = _new_like(graph, graphs, [int(v.shape[0]) for v in inds], nf, ef)
    else:
        nf, ef = _frames_from(graph, [None] * nnt, einds, store_ids)
        sg = _new_like(graph, graphs, list(graph._num_nodes), nf, ef)
    return sg if output_device is None else sg.to(output_device)


def _dir_subgraph(graph, nodes, direction, relabel_nodes, store_ids, output_device):
    if graph.is_block:
        raise DGLError("Extracting subgraph of a block graph is not allowed.")
    nodes = _node_dict(graph, nodes)
    idt, dev = graph.idtype, graph.device
    edges = {}
    for i, c in enumerate(graph.canonical_etypes):
        nt = c[2] if direction == "in" else c[0]
        if nt in nodes:
            v = _to_tensor(nodes[nt], idt, dev)
            eid = graph.in_edges(v, form="eid", etype=c) if direction == "in" else graph.out_edges(v, form="eid", etype=c)
            edges[c] = eid
        else:
            edges[c] = torch.empty(0, dtype=idt, device=dev)
    sg = edge_subgraph(graph, edges, relabel_nodes=relabel_nodes, store_ids=store_ids)
    return sg if output_device is None else sg.to(output_device)


def in_subgraph(graph, nodes, *, relabel_nodes=False, store_ids=True, output_device=None):
    return _dir_subgraph(graph, nodes, "in", relabel_nodes, store_ids, output_device)


def out_subgraph(graph, nodes, *, relabel_nodes=False, store_ids=True, output_device=None):
    return _dir_subgraph(graph, nodes, "out", relabel_nodes, store_ids, output_device)


def _khop(graph, nodes, k, direction, relabel_nodes, store_ids, output_device):
    if graph.is_block:
        raise DGLError("Extracting subgraph of a block graph is not allowed.")
    is_mapping = isinstance(nodes, dict)
    nodes = _node_dict(graph, nodes)
    idt, dev = graph.idtype, graph.device
    nodes = {nt: _to_tensor(v, idt, dev) for nt, v in nodes.items()}
    last = {nt: nodes.get(nt, torch.empty(0, dtype=idt, device=dev)) for nt in graph.ntypes}
    allv = {nt: [last[nt]] for nt in graph.ntypes}
    for _ in range(k):
        cur = {nt: [] for nt in graph.ntypes}
        for c in graph.canonical_etypes:
            s, _, d = c
            if direction == "in":
                if last[d].numel():
                    cur[s].append(graph.in_edges(last[d], etype=c)[0])
            else:
                if last[s].numel():
                    cur[d].append(graph.out_edges(last[s], etype=c)[1])
        last = {nt: (torch.unique(torch.cat(v)) if v else torch.empty(0, dtype=idt, device=dev)) for nt, v in cur.items()}
        for nt in graph.ntypes:
            allv[nt].append(last[nt])
    kn = {nt: torch.unique(torch.cat(v)) for nt, v in allv.items()}
    sg = node_subgraph(graph, kn, relabel_nodes=relabel_nodes, store_ids=store_ids)
    if output_device is not None:
        sg = sg.to(output_device)
    if relabel_nodes:
        inv = {}
        for nt, seeds in nodes.items():
            inv[nt] = torch.searchsorted(kn[nt], seeds).to(idt)
            if output_device is not None:
                inv[nt] = inv[nt].to(output_device)
        return sg, (inv if is_mapping else inv[graph.ntypes[0]])
    return sg


def khop_in_subgraph(graph, nodes, k, *, relabel_nodes=True, store_ids=True, output_device=None):
    return _khop(graph, nodes, k, "in", relabel_nodes, store_ids, output_device)


def khop_out_subgraph(graph, nodes, k, *, relabel_nodes=True, store_ids=True, output_device=None):
    return _khop(graph, nodes, k, "out", relabel_nodes, store_ids, output_device)


def node_type_subgraph(graph, ntypes, output_device=None):
    ntypes = set(ntypes)
    etypes = [c for c in graph.canonical_etypes if c[0] in ntypes and c[2] in ntypes]
    if not etypes:
        raise DGLError("There are no edges among nodes of the specified types.")
    return edge_type_subgraph(graph, etypes, output_device)


def edge_type_subgraph(graph, etypes, output_device=None):
    cets = [graph.to_canonical_etype(e) for e in etypes]
    etids = [graph.get_etype_id(c) for c in cets]
    nts = []
    for c in cets:
        for t in (c[0], c[2]):
            if t not in nts:
                nts.append(t)
    nts = [t for t in graph.ntypes if t in nts]
    g = DGLGraph.__new__(DGLGraph)
    order = sorted(range(len(cets)), key=lambda i: cets[i])
    g._init([graph._graphs[etids[i]] for i in order], nts, [cets[i] for i in order],
            [graph.num_nodes(t) for t in nts], [graph._node_frames[graph.get_ntype_id(t)] for t in nts],
            [graph._edge_frames[etids[i]] for i in order])
    return g if output_device is None else g.to(output_device)
