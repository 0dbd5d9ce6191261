"""OnDiskDataset / BuiltinDataset (parity: /root/reference/python/dgl/graphbolt/impl/ondisk_dataset.py (1077 lines),
ondisk_metadata.py, dataset.py). The on-disk layout is a directory with ``metadata.yaml`` describing graph
structure files (csv / numpy edge lists), feature files and task splits; ``load()`` preprocesses once into a
FusedCSCSamplingGraph saved under ``preprocessed/`` keyed by a hash of the metadata."""
from __future__ import annotations

import hashlib
import json
import os
import shutil

import numpy as np
import torch
import yaml

from ...base import DGLError
from ...graph_index import coo_to_csr
from ..base import etype_str_to_tuple
from ..itemset import HeteroItemSet, ItemSet
from .features import TorchBasedFeatureStore
from .fused_csc_sampling_graph import FusedCSCSamplingGraph, fused_csc_sampling_graph


class OnDiskTask:
    def __init__(self, metadata, train_set, validation_set, test_set):
        self._metadata, self._train_set, self._validation_set, self._test_set = metadata, train_set, validation_set, test_set

    @property
    def metadata(self):
        return self._metadata

    @property
    def train_set(self):
        return self._train_set

    @property
    def validation_set(self):
        return self._validation_set

    @property
    def test_set(self):
        return self._test_set

    def __repr__(self):
        return f"OnDiskTask(validation_set={self._validation_set},\n           train_set={self._train_set},\n           test_set={self._test_set},\n           metadata={self._metadata},)"


def _read_array(path, fmt, in_memory=True):
    if fmt == "numpy":
        arr = np.load(path, mmap_mode=None if in_memory else "r")
        return torch.from_numpy(np.ascontiguousarray(arr))
    if fmt == "torch":
        return torch.load(path)
    if fmt == "csv":
        import pandas as pd
        return torch.from_numpy(pd.read_csv(path, header=None).to_numpy())
    raise DGLError(f"unknown data format {fmt}")


def _build_graph(dataset_dir, graph_spec, include_original_edge_id=False):
    nodes = graph_spec.get("nodes", [])
    edges = graph_spec["edges"]
    is_homo = len(nodes) <= 1 and len(edges) == 1 and nodes and nodes[0].get("type") is None or (len(edges) == 1 and edges[0].get("type") is None)
    if is_homo:
        e = _read_array(os.path.join(dataset_dir, edges[0]["path"]), edges[0].get("format", "csv"))
        if e.shape[0] == 2 and e.dim() == 2 and e.shape[1] != 2:
            src, dst = e[0], e[1]
        else:
            src, dst = e[:, 0], e[:, 1]
        n = nodes[0]["num"] if nodes else int(max(src.max(), dst.max())) + 1
        indptr, indices, eids = coo_to_csr(dst.long(), src.long(), n, torch.int64)
        ea = {}
        if include_original_edge_id:
            from ..base import ORIGINAL_EDGE_ID
            ea[ORIGINAL_EDGE_ID] = eids if eids is not None else torch.arange(indices.shape[0])
        return fused_csc_sampling_graph(indptr, indices, edge_attributes=ea)
    ntypes = [n["type"] for n in nodes]
    ntid = {t: i for i, t in enumerate(ntypes)}
    off = np.cumsum([0] + [n["num"] for n in nodes])
    etid = {e["type"]: i for i, e in enumerate(edges)}
    srcs, dsts, types = [], [], []
    for e in edges:
        arr = _read_array(os.path.join(dataset_dir, e["path"]), e.get("format", "csv"))
        if arr.shape[0] == 2 and arr.dim() == 2 and arr.shape[1] != 2:
            s, d = arr[0], arr[1]
        else:
            s, d = arr[:, 0], arr[:, 1]
        st, _, dt = etype_str_to_tuple(e["type"])
        srcs.append(s.long() + int(off[ntid[st]])); dsts.append(d.long() + int(off[ntid[dt]]))
        types.append(torch.full((s.shape[0],), etid[e["type"]], dtype=torch.int64))
    src, dst, tpe = torch.cat(srcs), torch.cat(dsts), torch.cat(types)
    total = int(off[-1])
    indptr, indices, eids = coo_to_csr(dst, src, total, torch.int64)
    if eids is not None:
        tpe = tpe[eids.long()]
    return fused_csc_sampling_graph(indptr, indices, torch.from_numpy(off.astype(np.int64)), tpe, ntid, etid)


class OnDiskDataset:
    def __init__(self, path, include_original_edge_id=False, force_preprocess=None, auto_cast_to_optimal_dtype=True):
        self._dataset_dir = path
        with open(os.path.join(path, "metadata.yaml")) as f:
            self._meta = yaml.safe_load(f)
        self._include_original_edge_id = include_original_edge_id
        self._force_preprocess = force_preprocess
        self._loaded = False
        self._graph = None
        self._feature = None
        self._tasks = []
        self._all_nodes_set = None

    def _preprocess_graph(self):
        digest = hashlib.sha1(json.dumps(self._meta, sort_keys=True, default=str).encode()).hexdigest()[:16]
        pdir = os.path.join(self._dataset_dir, "preprocessed")
        gpath = os.path.join(pdir, f"fused_csc_sampling_graph_{digest}.pt")
        if os.path.exists(gpath) and not self._force_preprocess:
            return torch.load(gpath, weights_only=False)
        g = _build_graph(self._dataset_dir, self._meta["graph"], self._include_original_edge_id)
        os.makedirs(pdir, exist_ok=True)
        torch.save(g, gpath)
        return g

    def load(self, tasks=None):
        if "graph" in self._meta:
            self._graph = self._preprocess_graph()
        elif "graph_topology" in self._meta:
            self._graph = torch.load(os.path.join(self._dataset_dir, self._meta["graph_topology"]["path"]), weights_only=False)
        feats = []
        for spec in self._meta.get("feature_data", []) or []:
            s = dict(spec)
            s["path"] = os.path.join(self._dataset_dir, spec["path"])
            feats.append(s)
        self._feature = TorchBasedFeatureStore(feats)
        self._tasks = []
        for t in self._meta.get("tasks", []) or []:
            if tasks is not None and t.get("name") not in ([tasks] if isinstance(tasks, str) else tasks):
                continue
            sets = {}
            for split in ("train_set", "validation_set", "test_set"):
                sets[split] = self._load_tvt(t.get(split))
            extra = {k: v for k, v in t.items() if k not in ("train_set", "validation_set", "test_set")}
            self._tasks.append(OnDiskTask(extra, sets["train_set"], sets["validation_set"], sets["test_set"]))
        self._loaded = True
        return self

    def _load_tvt(self, spec):
        if not spec:
            return None
        per_type = {}
        for entry in spec:
            names, tensors = [], []
            for d in entry["data"]:
                names.append(d["name"])
                tensors.append(_read_array(os.path.join(self._dataset_dir, d["path"]), d.get("format", "numpy"),
                                           d.get("in_memory", True)))
            items = tensors[0] if len(tensors) == 1 else tuple(tensors)
            per_type[entry.get("type")] = ItemSet(items, names=tuple(names))
        if list(per_type.keys()) == [None]:
            return per_type[None]
        return HeteroItemSet(per_type)

    @property
    def dataset_name(self):
        return self._meta.get("dataset_name")

    @property
    def graph(self):
        return self._graph

    @property
    def feature(self):
        return self._feature

    @property
    def tasks(self):
        return self._tasks

    @property
    def all_nodes_set(self):
        n = self._graph.num_nodes
        if isinstance(n, dict):
            return HeteroItemSet({t: ItemSet(c, names="seeds") for t, c in n.items()})
        return ItemSet(n, names="seeds")

    @property
    def yaml_data(self):
        return self._meta


class BuiltinDataset(OnDiskDataset):
    """Builtin datasets of the reference are downloads (ogbn-*, cora ...). There is no network here: the
    dataset is synthesised with the published shape, written in OnDiskDataset layout under ``root`` and loaded."""

    _all_datasets = ["cora", "ogbn-arxiv", "ogbn-products", "reddit"]

    def __init__(self, name, root="datasets", scale=1.0):
        from ...data.synthetic import SHAPES, rmat_edges
        if name not in SHAPES:
            raise DGLError(f"Dataset {name} is not available. Available datasets are {list(SHAPES)}.")
        path = os.path.join(root, f"{name}-synthetic-{scale}")
        if not os.path.exists(os.path.join(path, "metadata.yaml")):
            n, e, f, c, sym = SHAPES[name]
            n, e = max(64, int(n * scale)), max(256, int(e * scale))
            os.makedirs(path, exist_ok=True)
            u, v = rmat_edges(n, e, seed=0)
            if sym:
                u, v = torch.cat([u, v]), torch.cat([v, u])
            np.save(os.path.join(path, "edges.npy"), torch.stack([u, v], 1).numpy())
            np.save(os.path.join(path, "feat.npy"), torch.randn(n, f).numpy())
            labels = torch.randint(0, c, (n,))
            perm = torch.randperm(n)
            splits = {"train": perm[: n // 10], "validation": perm[n // 10: n // 5], "test": perm[n // 5: n // 3]}
            for k, idx in splits.items():
                np.save(os.path.join(path, f"{k}_seeds.npy"), idx.numpy())
                np.save(os.path.join(path, f"{k}_labels.npy"), labels[idx].numpy())
            meta = {"dataset_name": name, "graph": {"nodes": [{"num": n}], "edges": [{"format": "numpy", "path": "edges.npy"}]},
                    "feature_data": [{"domain": "node", "type": None, "name": "feat", "format": "numpy", "in_memory": True, "path": "feat.npy"}],
                    "tasks": [{"name": "node_classification", "num_classes": c,
                               **{f"{k}_set": [{"type": None, "data": [{"name": "seeds", "format": "numpy", "path": f"{k}_seeds.npy"},
                                                                         {"name": "labels", "format": "numpy", "path": f"{k}_labels.npy"}]}]
                                  for k in ("train", "validation", "test")}}]}
            with open(os.path.join(path, "metadata.yaml"), "w") as fh:
                yaml.safe_dump(meta, fh)
        super().__init__(path)
