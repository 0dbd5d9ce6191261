#!/usr/bin/env python
"""Write a graph in the *chunked* on-disk format consumed by the distributed partitioning pipeline
(parity: /root/reference/tools/chunk_graph.py + docs "chunked graph format"):

  <out>/metadata.json                     graph_name, node/edge types, per-chunk counts, file lists
  <out>/edge_index/<src>:<rel>:<dst><i>.npy     int64 [n_i, 2] edge chunks
  <out>/node_data/<ntype>/<name>-<i>.npy        feature chunks        (edge_data/... likewise)

``chunk_graph(g, name, ndata_paths, edata_paths, num_chunks, output_path)`` is importable; the CLI chunks a graph saved
with ``dgl.save_graphs``."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def _chunks(n, k):
    base, rem = divmod(n, k)
    sizes = [base + (1 if i < rem else 0) for i in range(k)]
    offs = np.cumsum([0] + sizes)
    return sizes, offs


def chunk_graph(g, name, ndata_paths=None, edata_paths=None, num_chunks=2, output_path="chunked", data_fmt="numpy"):
    """``ndata_paths`` / ``edata_paths``: optional {type: {feat_name: path-to-.npy}} for features kept outside ``g``;
    features stored in ``g.ndata`` / ``g.edata`` are chunked automatically."""
    os.makedirs(os.path.join(output_path, "edge_index"), exist_ok=True)
    meta = {"graph_name": name, "node_type": list(g.ntypes), "num_nodes_per_chunk": [],
            "edge_type": [":".join(c) for c in g.canonical_etypes], "num_edges_per_chunk": [], "edges": {},
            "node_data": {}, "edge_data": {}}
    for nt in g.ntypes:
        sizes, _ = _chunks(g.num_nodes(nt), num_chunks)
        meta["num_nodes_per_chunk"].append(sizes)
    for c in g.canonical_etypes:
        key = ":".join(c)
        u, v = g.edges(etype=c)
        e = np.stack([u.cpu().numpy(), v.cpu().numpy()], 1).astype(np.int64)
        sizes, offs = _chunks(e.shape[0], num_chunks)
        files = []
        for i in range(num_chunks):
            p = os.path.join(output_path, "edge_index", f"{key}{i}.npy")
            np.save(p, e[offs[i]:offs[i + 1]])
            files.append(os.path.relpath(p, output_path))
        meta["num_edges_per_chunk"].append(sizes)
        meta["edges"][key] = {"format": {"name": "numpy"}, "data": files}

    def dump(domain, typ, feat, arr, n):
        d = os.path.join(output_path, domain, typ)
        os.makedirs(d, exist_ok=True)
        _, offs = _chunks(n, num_chunks)
        files = []
        for i in range(num_chunks):
            p = os.path.join(d, f"{feat}-{i}.npy")
            np.save(p, arr[offs[i]:offs[i + 1]])
            files.append(os.path.relpath(p, output_path))
        meta[domain].setdefault(typ, {})[feat] = {"format": {"name": "numpy"}, "data": files}

    for nt in g.ntypes:
        frame = g.nodes[nt].data if len(g.ntypes) > 1 else g.ndata
        for k, v in frame.items():
            if torch.is_tensor(v):
                dump("node_data", nt, k, v.cpu().numpy(), g.num_nodes(nt))
        for k, p in (ndata_paths or {}).get(nt, {}).items():
            dump("node_data", nt, k, np.load(p), g.num_nodes(nt))
    for c in g.canonical_etypes:
        key = ":".join(c)
        frame = g.edges[c].data if len(g.canonical_etypes) > 1 else g.edata
        for k, v in frame.items():
            if torch.is_tensor(v):
                dump("edge_data", key, k, v.cpu().numpy(), g.num_edges(c))
        for k, p in (edata_paths or {}).get(key, {}).items():
            dump("edge_data", key, k, np.load(p), g.num_edges(c))
    with open(os.path.join(output_path, "metadata.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return meta


def load_chunked(path):
    """Re-assemble a chunked graph (homogeneous or heterogeneous) into a DGLGraph with its features."""
    import dgl_b200 as dgl
    with open(os.path.join(path, "metadata.json")) as f:
        meta = json.load(f)
    nn_ = {nt: int(sum(s)) for nt, s in zip(meta["node_type"], meta["num_nodes_per_chunk"])}
    data = {}
    for key, spec in meta["edges"].items():
        e = np.concatenate([np.load(os.path.join(path, p)) for p in spec["data"]])
        data[tuple(key.split(":"))] = (torch.from_numpy(e[:, 0]), torch.from_numpy(e[:, 1]))
    if len(nn_) == 1 and len(data) == 1:
        (u, v), = data.values()
        g = dgl.graph((u, v), num_nodes=next(iter(nn_.values())))
    else:
        g = dgl.heterograph(data, num_nodes_dict=nn_)
    for dom, attr in (("node_data", "nodes"), ("edge_data", "edges")):
        for typ, feats in meta[dom].items():
            for k, spec in feats.items():
                arr = torch.from_numpy(np.concatenate([np.load(os.path.join(path, p)) for p in spec["data"]]))
                if len(nn_) == 1 and len(data) == 1:
                    (g.ndata if dom == "node_data" else g.edata)[k] = arr
                else:
                    getattr(g, attr)[typ if dom == "node_data" else tuple(typ.split(":"))].data[k] = arr
    return g, meta


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", required=True, help="file written by dgl.save_graphs")
    ap.add_argument("--name", default="graph")
    ap.add_argument("--num-chunks", type=int, default=2)
    ap.add_argument("--output", required=True)
    a = ap.parse_args()
    import dgl_b200 as dgl
    gs, _ = dgl.load_graphs(a.graph)
    chunk_graph(gs[0], a.name, num_chunks=a.num_chunks, output_path=a.output)
    print("chunked graph written to", a.output)
