"""Batching of graphs into one disjoint-union graph.

Parity: /root/reference/python/dgl/batch.py (batch :15, unbatch :262, slice_batch :403).
"""
from __future__ import annotations

import torch

from .base import DGLError
from .frame import Frame
from .graph_index import UnitGraph
from .heterograph import DGLGraph

__all__ = ["batch", "unbatch", "slice_batch", "batch_hetero", "unbatch_hetero"]


def batch(graphs, ndata="__ALL__", edata="__ALL__"):
    """Disjoint union; node/edge ids are shifted by the running totals, features concatenated."""
    if len(graphs) == 0:
        raise DGLError("The input list of graphs cannot be empty.")
    g0 = graphs[0]
    if ndata is not None and ndata != "__ALL__" and not isinstance(ndata, (list, tuple)):
        raise DGLError("Invalid argument ndata: must be a string list but got {}.".format(type(ndata)))
    if any(g.is_block for g in graphs):
        raise DGLError("Batching a MFG is not supported.")
    ntypes, cets = g0.ntypes, g0.canonical_etypes
    idt, dev = g0.idtype, g0.device
    for g in graphs:
        if g.ntypes != ntypes or g.canonical_etypes != cets:
            raise DGLError("All graphs should have the same node types and edge types.")
        if g.idtype != idt:
            raise DGLError(f"Expect graph data type to be {idt}, but got {g.idtype}")
        if g.device != dev:
            raise DGLError(f"Expect graph device to be {dev}, but got {g.device}")
    bnn = {t: torch.cat([g.batch_num_nodes(t) for g in graphs]) for t in ntypes}
    bne = {c: torch.cat([g.batch_num_edges(c) for g in graphs]) for c in cets}
    num_nodes = [int(sum(g.num_nodes(t) for g in graphs)) for t in ntypes]
    offs = {t: torch.cumsum(torch.tensor([0] + [g.num_nodes(t) for g in graphs[:-1]]), 0).tolist() for t in ntypes}
    ugs = []
    for i, (s, e, d) in enumerate(cets):
        rows, cols = [], []
        for gi, g in enumerate(graphs):
            r, c = g._graphs[i].coo()
            rows.append(r + offs[s][gi])
            cols.append(c + offs[d][gi])
        ugs.append(UnitGraph.from_coo(num_nodes[ntypes.index(s)], num_nodes[ntypes.index(d)], torch.cat(rows),
                                      torch.cat(cols)))
    out = DGLGraph(ugs, ntypes, cets, num_nodes)
    out.set_batch_num_nodes(bnn)
    out.set_batch_num_edges(bne)

    def cat_frames(frames, keys_spec, what):
        if keys_spec is None:
            return {}
        nonempty = [f for f in frames if f.num_rows > 0] or frames
        keys = list(nonempty[0].keys()) if keys_spec == "__ALL__" else list(keys_spec)
        ret = {}
        for k in keys:
            for f in nonempty:
                if k not in f:
                    raise DGLError(f"feature '{k}' is missing on some {what} of the batched graphs")
            ret[k] = torch.cat([f[k] for f in frames if k in f and (f.num_rows > 0 or len(nonempty) == len(frames))], 0)
        return ret

    for j, t in enumerate(ntypes):
        data = cat_frames([g._node_frames[j] for g in graphs], ndata, "nodes")
        out._node_frames[j] = Frame(data, num_rows=num_nodes[j])
    for i, c in enumerate(cets):
        data = cat_frames([g._edge_frames[i] for g in graphs], edata, "edges")
        out._edge_frames[i] = Frame(data, num_rows=ugs[i].num_edges())
    return out


def unbatch(g, node_split=None, edge_split=None):
    """Split a batched graph back into its components."""
    ntypes, cets = g.ntypes, g.canonical_etypes
    if node_split is None:
        node_split = {t: g.batch_num_nodes(t) for t in ntypes}
    elif not isinstance(node_split, dict):
        node_split = {ntypes[0]: node_split}
    if edge_split is None:
        edge_split = {c: g.batch_num_edges(c) for c in cets}
    elif not isinstance(edge_split, dict):
        edge_split = {cets[0]: edge_split}
    edge_split = {g.to_canonical_etype(k): v for k, v in edge_split.items()}
    nb = len(next(iter(node_split.values())))
    ns = {t: [int(x) for x in node_split[t].tolist()] for t in ntypes}
    es = {c: [int(x) for x in edge_split[c].tolist()] for c in cets}
    for t in ntypes:
        if sum(ns[t]) != g.num_nodes(t):
            raise DGLError("node_split does not sum to the number of nodes")
    noff = {t: [0] for t in ntypes}
    for t in ntypes:
        for x in ns[t]:
            noff[t].append(noff[t][-1] + x)
    eoff = {c: [0] for c in cets}
    for c in cets:
        for x in es[c]:
            eoff[c].append(eoff[c][-1] + x)
    coos = [ug.coo() for ug in g._graphs]
    outs = []
    for b in range(nb):
        ugs = []
        for i, (s, e, d) in enumerate(cets):
            lo, hi = eoff[cets[i]][b], eoff[cets[i]][b + 1]
            r, c = coos[i]
            ugs.append(UnitGraph.from_coo(ns[s][b], ns[d][b], r[lo:hi] - noff[s][b], c[lo:hi] - noff[d][b]))
        sub = DGLGraph(ugs, ntypes, cets, [ns[t][b] for t in ntypes])
        for j, t in enumerate(ntypes):
            lo, hi = noff[t][b], noff[t][b + 1]
            sub._node_frames[j] = Frame({k: v[lo:hi] for k, v in g._node_frames[j].items()}, num_rows=hi - lo)
        for i, c in enumerate(cets):
            lo, hi = eoff[c][b], eoff[c][b + 1]
            sub._edge_frames[i] = Frame({k: v[lo:hi] for k, v in g._edge_frames[i].items()}, num_rows=hi - lo)
        outs.append(sub)
    return outs


def slice_batch(g, gid, store_ids=False):
    """Extract the ``gid``-th graph of a batch (reference batch.py:403)."""
    from .base import EID, NID
    ntypes, cets = g.ntypes, g.canonical_etypes
    nstart, nnum = {}, {}
    for t in ntypes:
        bn = g.batch_num_nodes(t)
        nstart[t] = int(bn[:gid].sum())
        nnum[t] = int(bn[gid])
    ugs, estart, enum_ = [], {}, {}
    for i, (s, e, d) in enumerate(cets):
        be = g.batch_num_edges(cets[i])
        lo = int(be[:gid].sum())
        n = int(be[gid])
        estart[i], enum_[i] = lo, n
        r, c = g._graphs[i].coo()
        ugs.append(UnitGraph.from_coo(nnum[s], nnum[d], r[lo:lo + n] - nstart[s], c[lo:lo + n] - nstart[d]))
    sub = DGLGraph(ugs, ntypes, cets, [nnum[t] for t in ntypes])
    for j, t in enumerate(ntypes):
        lo, hi = nstart[t], nstart[t] + nnum[t]
        data = {k: v[lo:hi] for k, v in g._node_frames[j].items()}
        if store_ids:
            data[NID] = torch.arange(lo, hi, dtype=g.idtype, device=g.device)
        sub._node_frames[j] = Frame(data, num_rows=hi - lo)
    for i in range(len(cets)):
        lo, hi = estart[i], estart[i] + enum_[i]
        data = {k: v[lo:hi] for k, v in g._edge_frames[i].items()}
        if store_ids:
            data[EID] = torch.arange(lo, hi, dtype=g.idtype, device=g.device)
        sub._edge_frames[i] = Frame(data, num_rows=hi - lo)
    return sub


batch_hetero = batch
unbatch_hetero = unbatch
