"""save_graphs / load_graphs / save_info / load_info (API parity: /root/reference/python/dgl/data/graph_serialize.py:83,149).

Container: one ``torch.save`` archive holding ONLY tensors and plain Python containers (magic/version header, COO
structure per relation, node/edge frames, optional label dict), so it is read back with ``weights_only=True``: loading a
graph file never unpickles arbitrary objects. Heterographs, blocks and batched graphs round-trip.

The reference's binary layout (src/graph/serialize/*, dmlc streams of NDArrays) is implemented by ``dgl_binary.py``:
``load_graphs`` recognises such files by their magic and reads them in place (memory mapped), and
``save_graphs(..., container="dgl")`` -- the default for ``*.bin`` / ``*.dgl`` names of plain ``DGLGraph`` lists, the
names the reference's users and its partition tools produce -- writes files the reference's ``dgl.load_graphs`` reads.
Blocks and batched graphs carry state that layout has no field for, so they always use the tensor archive."""
from __future__ import annotations

import os
import pickle

import torch

from ..base import DGLError
from ..frame import Frame
from ..graph_index import UnitGraph
from ..heterograph import DGLGraph
from . import dgl_binary

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


def _binary_ok(g_list):
    return all(type(g) is DGLGraph and g._srctypes is None and g._batch_num_nodes is None and g._batch_num_edges is None
               for g in g_list)


def save_graphs(filename, g_list, labels=None, formats=None, container=None):
    """``container``: "dgl" (reference binary layout), "archive" (tensor-only torch archive) or None = by file name."""
    if isinstance(g_list, DGLGraph):
        g_list = [g_list]
    if labels is not None and not isinstance(labels, dict):
        raise DGLError("labels must be a dict of tensors")
    if os.path.isdir(filename):
        raise DGLError(f"Filename {filename} is an existing directory.")
    d = os.path.dirname(os.path.abspath(filename))
    os.makedirs(d, exist_ok=True)
    if container is None:
        container = "dgl" if str(filename).endswith((".bin", ".dgl")) and _binary_ok(g_list) else "archive"
    if container == "dgl":
        if not _binary_ok(g_list):
            raise DGLError("the DGL binary layout stores plain DGLGraph objects only (no blocks / batched graphs)")
        dgl_binary.write_graphs(filename, g_list, labels, formats)
        return
    if container != "archive":
        raise DGLError(f"unknown container {container!r}")
    torch.save({"magic": MAGIC, "version": VERSION, "graphs": [_pack(g) for g in g_list], "labels": labels or {}}, filename)


def _read_archive(filename):
    if not os.path.exists(filename):
        raise DGLError(f"File {filename} does not exist.")
    try:
        d = torch.load(filename, weights_only=True, map_location="cpu")
    except Exception as e:  # noqa: BLE001 - not a torch archive, or it contains objects: refuse, never unpickle them
        with open(filename, "rb") as fh:
            head = fh.read(8)
        raise DGLError(f"{filename} is not a graph archive of this framework: {type(e).__name__}: {e}") from None
    if not isinstance(d, dict) or d.get("magic") != MAGIC or not isinstance(d.get("graphs"), list):
        raise DGLError("Invalid DGL graph file (bad magic)")
    if int(d.get("version", 0)) > VERSION:
        raise DGLError(f"graph archive version {d.get('version')} is newer than this build understands ({VERSION})")
    return d


def load_graphs(filename, idx_list=None):
    if dgl_binary.is_dgl_binary(filename):
        return dgl_binary.read_graphs(filename, idx_list)
    d = _read_archive(filename)
    packs = d["graphs"]
    if idx_list is not None:
        try:
            packs = [packs[i] for i in idx_list]
        except IndexError:
            raise DGLError(f"idx_list {list(idx_list)} out of range: the file holds {len(packs)} graphs") from None
    try:
        return [_unpack(p) for p in packs], d.get("labels", {})
    except (KeyError, TypeError, ValueError) as e:
        raise DGLError(f"corrupt graph archive {filename}: {type(e).__name__}: {e}") from None


def load_labels(filename):
    if dgl_binary.is_dgl_binary(filename):
        return dgl_binary.read_labels(filename)
    return _read_archive(filename).get("labels", {})


def save_info(path, info):
    with open(path, "wb") as f:
        pickle.dump(info, f)


def load_info(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def __getattr__(name):
    # container records and the versioned loaders of the reference's graph_serialize.py live in compat.py
    if name in ("StorageMetaData", "GraphData", "is_local_path", "check_local_file_exists", "load_graph_v1",
                "load_graph_v2", "load_labels_v1", "load_labels_v2"):
        from . import compat
        return getattr(compat, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
