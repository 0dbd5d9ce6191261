"""save_graphs / load_graphs / save_info / load_info (parity: /root/reference/python/dgl/data/graph_serialize.py:83,149 and
the native binary format src/graph/serialize/*). Container: one ``torch.save`` archive with a magic/version header,
COO structure per relation, node/edge frames and optional label dict; heterographs round-trip."""
from __future__ import annotations

import os
import pickle

import torch

from ..base import DGLError
from ..frame import Frame
from ..graph_index import UnitGraph
from ..heterograph import DGLGraph

MAGIC = "DGLB200-GRAPHS"
VERSION = 1


def _pack(g):
    return {"ntypes": g._ntypes if g._srctypes is None else None, "srctypes": g._srctypes, "dsttypes": g._dsttypes,
            "cet": g._canonical_etypes, "num_nodes": g._num_nodes, "idtype": str(g.idtype),
            "coo": [tuple(t.cpu() for t in ug.coo()) for ug in g._graphs],
            "nframes": [{k: f[k].cpu() for k in f.keys()} for f in g._node_frames],
            "eframes": [{k: f[k].cpu() for k in f.keys()} for f in g._edge_frames],
            "bnn": None if g._batch_num_nodes is None else {k: v.cpu() for k, v in g._batch_num_nodes.items()},
            "bne": None if g._batch_num_edges is None else {k: v.cpu() for k, v in g._batch_num_edges.items()},
            "is_block": g.is_block}


def _unpack(d):
    from ..heterograph import DGLBlock
    cet = [tuple(c) for c in d["cet"]]
    if d["srctypes"] is not None:
        nn_src = {t: n for t, n in zip(d["srctypes"], d["num_nodes"][: len(d["srctypes"])])}
        nn_dst = {t: n for t, n in zip(d["dsttypes"], d["num_nodes"][len(d["srctypes"]):])}
        graphs = [UnitGraph.from_coo(nn_src[c[0]], nn_dst[c[2]], r, cl) for c, (r, cl) in zip(cet, d["coo"])]
        g = (DGLBlock if d.get("is_block") else DGLGraph)(graphs, None, cet, d["num_nodes"], srctypes=d["srctypes"], dsttypes=d["dsttypes"])
    else:
        nn = {t: n for t, n in zip(d["ntypes"], d["num_nodes"])}
        graphs = [UnitGraph.from_coo(nn[c[0]], nn[c[2]], r, cl) for c, (r, cl) in zip(cet, d["coo"])]
        g = DGLGraph(graphs, d["ntypes"], cet, d["num_nodes"])
    g._node_frames = [Frame(f, num_rows=n) for f, n in zip(d["nframes"], d["num_nodes"])]
    g._edge_frames = [Frame(f, num_rows=ug.num_edges()) for f, ug in zip(d["eframes"], graphs)]
    g._batch_num_nodes, g._batch_num_edges = d["bnn"], d["bne"]
    return g


def save_graphs(filename, g_list, labels=None, formats=None):
    if isinstance(g_list, DGLGraph):
        g_list = [g_list]
    if labels is not None and not isinstance(labels, dict):
        raise DGLError("labels must be a dict of tensors")
    d = os.path.dirname(os.path.abspath(filename))
    os.makedirs(d, exist_ok=True)
    torch.save({"magic": MAGIC, "version": VERSION, "graphs": [_pack(g) for g in g_list], "labels": labels or {}}, filename)


def load_graphs(filename, idx_list=None):
    if not os.path.exists(filename):
        raise DGLError(f"File {filename} does not exist.")
    d = torch.load(filename, weights_only=False)
    if d.get("magic") != MAGIC:
        raise DGLError("Invalid DGL graph file (bad magic)")
    packs = d["graphs"]
    if idx_list is not None:
        packs = [packs[i] for i in idx_list]
    return [_unpack(p) for p in packs], d["labels"]


def load_labels(filename):
    return torch.load(filename, weights_only=False)["labels"]


def save_info(path, info):
    with open(path, "wb") as f:
        pickle.dump(info, f)


def load_info(path):
    with open(path, "rb") as f:
        return pickle.load(f)
