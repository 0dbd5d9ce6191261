"""Vertex-cut full-graph distributed training helpers (parity: /root/reference/python/dgl/distgnn/): Libra partitioning
of edges + per-partition subgraph construction with replicated (split) vertices."""
from __future__ import annotations

import os

import torch

from ..convert import graph as make_graph
from ..partition import libra_partition_assignment


def partition_graph(num_community, G, resultdir):
    """Libra vertex-cut partition: every edge is owned by exactly one partition; vertices incident to edges of
    several partitions are replicated (``split`` nodes). Writes part<i>.pt files and returns the list of parts."""
    ep = libra_partition_assignment(G, num_community)
    u, v = G.edges()
    os.makedirs(resultdir, exist_ok=True)
    parts = []
    for p in range(num_community):
        m = ep == p
        pu, pv = u[m], v[m]
        nodes = torch.unique(torch.cat([pu, pv]))
        local = torch.full((G.num_nodes(),), -1, dtype=torch.int64)
        local[nodes] = torch.arange(nodes.shape[0])
        pg = make_graph((local[pu.long()], local[pv.long()]), num_nodes=nodes.shape[0])
        pg.ndata["orig_id"] = nodes
        pg.edata["orig_id"] = torch.nonzero(m).squeeze(1)
        torch.save({"src": pg.edges()[0], "dst": pg.edges()[1], "nodes": nodes, "eids": pg.edata["orig_id"]},
                   os.path.join(resultdir, f"part{p}.pt"))
        parts.append(pg)
    # replication factor: how many partitions hold each vertex
    cnt = torch.zeros(G.num_nodes(), dtype=torch.int64)
    for pg in parts:
        cnt[pg.ndata["orig_id"]] += 1
    return parts, cnt
