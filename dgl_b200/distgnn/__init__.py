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


def libra_partition(num_community, G, resultdir):
    """Reference spelling (distgnn/partition/libra_partition.py:38) of :func:`partition_graph`; additionally writes
    ``replicationlist.csv`` (one line per vertex copy beyond the first) that :func:`rep_per_node` reads."""
    parts, cnt = partition_graph(num_community, G, resultdir)
    with open(os.path.join(resultdir, "replicationlist.csv"), "w") as f:
        f.write("## The Indices of Nodes that are replicated :: Header\n")
        for v in torch.nonzero(cnt > 1).view(-1).tolist():
            for _ in range(int(cnt[v]) - 1):
                f.write(f"{v}\n")
    return parts, cnt


# ---- distgnn/tools/tools.py ------------------------------------------------------------------------------------------
def rep_per_node(prefix, num_community):
    """{node id (str): number of extra copies} read from ``<prefix>/replicationlist.csv``."""
    from ..base import DGLError
    counts = {}
    with open(os.path.join(prefix, "replicationlist.csv")) as f:
        f.readline()                                     # header comment
        for line in f:
            if line.startswith("#"):
                raise DGLError("[Bug] Read Hash char in rep_per_node func.")
            node = line.strip("\n")
            counts[node] = counts.get(node, 0) + 1
    if any(v >= num_community for v in counts.values()):
        raise DGLError("[Bug] Unexpected event in rep_per_node() in tools.py.")
    return counts


def download_proteins():
    """The proteins matrix lives on an external server; this environment has no network access, so the call explains
    how to proceed (place ``proteins.mtx`` in the working directory) instead of downloading."""
    from ..base import DGLError
    if not os.path.exists("proteins.mtx"):
        raise DGLError("Error: Failed to download Proteins dataset!! Aborting.. (no network: copy proteins.mtx here)")


def proteins_mtx2dgl(path="proteins.mtx", feat_size=128, nlabels=256, train_size=1000000, test_size=500000,
                     val_size=5000):
    """MatrixMarket file -> graph with (uninitialised) features, random labels and leading train/test/val masks, sized
    like the reference's proteins set (splits are clipped to the node count for smaller matrices)."""
    from scipy.io import mmread
    coo = mmread(path).tocoo()
    g = make_graph((torch.as_tensor(coo.row, dtype=torch.int64), torch.as_tensor(coo.col, dtype=torch.int64)),
                   num_nodes=int(max(coo.shape)))
    n = g.num_nodes()
    g.ndata["feat"] = torch.zeros(n, feat_size)
    bounds = [min(train_size, n), min(train_size + test_size, n), min(train_size + test_size + val_size, n)]
    for name, lo, hi in (("train_mask", 0, bounds[0]), ("test_mask", bounds[0], bounds[1]),
                         ("val_mask", bounds[1], bounds[2])):
        m = torch.zeros(n, dtype=torch.bool)
        m[lo:hi] = True
        g.ndata[name] = m
    g.ndata["label"] = torch.randint(0, nlabels, (n,))
    return g


def save(g, dataset):
    """Write ``<dataset>/graph.dgl`` and ``<dataset>/node_feat.dgl``."""
    from ..data import save_graphs, save_tensors
    os.makedirs(dataset, mode=0o775, exist_ok=True)
    save_tensors(os.path.join(dataset, "node_feat.dgl"), dict(g.ndata))
    save_graphs(os.path.join(dataset, "graph.dgl"), [g])


def load_proteins(dataset):
    """Graph saved by :func:`save` under ``dataset``; converts ``proteins.mtx`` first when the folder is missing."""
    from ..data import load_graphs
    graph_file = os.path.join(dataset, "graph.dgl")
    if not os.path.exists(graph_file):
        download_proteins()
        save(proteins_mtx2dgl(), dataset)
    return load_graphs(graph_file)[0][0]

from . import partition, tools  # noqa: F401,E402  (reference distgnn/__init__.py)
