"""Builtin dataset classes. The reference downloads real data (python/dgl/data/*.py, ~45 datasets); this environment
has no network, so each class synthesises a graph with the published node/edge/feature/class counts (synthetic.SHAPES),
keeps the reference's attribute surface (``dataset[0]`` graph with feat/label/train_mask/val_mask/test_mask,
``num_classes``) and caches the processed graph with save_graphs."""
from __future__ import annotations

import os

import torch

from ..convert import graph as make_graph
from .dgl_dataset import DGLBuiltinDataset, get_download_dir
from .graph_serialize import load_graphs, save_graphs
from .raw_readers import find_node_source, load_node_source
from .synthetic import SHAPES, rmat_edges


class _SyntheticNodeDataset(DGLBuiltinDataset):
    _key = None
    _source_key = None      # key into raw_readers.NODE_SOURCES when it differs from ``_key``
    _scale = 1.0

    def __init__(self, raw_dir=None, force_reload=False, verbose=False, transform=None, reverse_edge=True, scale=None,
                 reorder=False):
        self._scale = self._scale if scale is None else scale
        self._raw_options = {"reverse_edge": reverse_edge, "reorder": reorder}
        self._raw_root = find_node_source(self._source_key or self._key,
                                          raw_dir if raw_dir is not None else get_download_dir())
        if self._raw_root is not None:
            name, key = self._key, (self._key, "raw", reverse_edge, reorder)
        else:
            name, key = self._key + "-synthetic", (self._key, self._scale)
        super().__init__(name, raw_dir=raw_dir, hash_key=key, force_reload=force_reload, verbose=verbose,
                         transform=transform)

    @property
    def is_synthetic(self):
        """False when the graph was parsed from the reference's raw files found under ``raw_dir``."""
        return self._raw_root is None

    def process(self):
        if self._raw_root is not None:          # the reference's downloaded files are on disk: parse them
            parsed = load_node_source(self._source_key or self._key, self._raw_root, **self._raw_options)
            self._g, self._num_classes = parsed.graph, parsed.num_classes
            return
        n, e, f, c, sym = SHAPES[self._key]
        n, e = max(32, int(n * self._scale)), max(64, int(e * self._scale))
        u, v = rmat_edges(n, e // 2 if not sym else e, seed=0)
        u, v = torch.cat([u, v]), torch.cat([v, u])
        g = make_graph((u, v), num_nodes=n)
        gen = torch.Generator().manual_seed(1)
        g.ndata["feat"] = torch.randn(n, f, generator=gen)
        g.ndata["label"] = torch.randint(0, c, (n,), generator=gen)
        perm = torch.randperm(n, generator=gen)
        for name, lo, hi in (("train_mask", 0, 0.1), ("val_mask", 0.1, 0.3), ("test_mask", 0.3, 0.6)):
            m = torch.zeros(n, dtype=torch.bool)
            m[perm[int(lo * n): int(hi * n)]] = True
            g.ndata[name] = m
        self._g = g
        self._num_classes = c

    def has_cache(self):
        return os.path.exists(os.path.join(self.save_path, "graph.bin"))

    def save(self):
        save_graphs(os.path.join(self.save_path, "graph.bin"), [self._g], {"num_classes": torch.tensor([self._num_classes])})

    def load(self):
        gs, labels = load_graphs(os.path.join(self.save_path, "graph.bin"))
        self._g = gs[0]
        self._num_classes = int(labels["num_classes"][0])

    @property
    def num_classes(self):
        return self._num_classes

    num_labels = num_classes

    def __getitem__(self, idx):
        assert idx == 0, "This dataset has only one graph"
        return self._g if self._transform is None else self._transform(self._g)

    def __len__(self):
        return 1


class CoraGraphDataset(_SyntheticNodeDataset):
    _key = "cora"


class RedditDataset(_SyntheticNodeDataset):
    _key = "reddit"
    _scale = 0.01

    def __init__(self, self_loop=False, raw_dir=None, force_reload=False, verbose=False, transform=None, scale=None):
        self._source_key = "reddit_self_loop" if self_loop else "reddit"
        super().__init__(raw_dir, force_reload, verbose, transform, scale=scale)


class OgbnProductsSynthetic(_SyntheticNodeDataset):
    _key = "ogbn-products"
    _scale = 0.01


class OgbnArxivSynthetic(_SyntheticNodeDataset):
    _key = "ogbn-arxiv"
    _scale = 0.1




class AsNodePredDataset:
    """Adapter exposing train/val/test index tensors (reference data/adapter.py:AsNodePredDataset)."""

    def __init__(self, dataset, split_ratio=None, target_ntype=None, **kwargs):
        self.dataset = dataset
        self.g = dataset[0]
        self.num_classes = getattr(dataset, "num_classes", int(self.g.ndata["label"].max()) + 1)
        n = self.g.num_nodes()
        if split_ratio is not None or "train_mask" not in self.g.ndata:
            r = split_ratio or (0.8, 0.1, 0.1)
            perm = torch.randperm(n)
            a, b = int(r[0] * n), int((r[0] + r[1]) * n)
            for name, idx in (("train_mask", perm[:a]), ("val_mask", perm[a:b]), ("test_mask", perm[b:])):
                m = torch.zeros(n, dtype=torch.bool)
                m[idx] = True
                self.g.ndata[name] = m
        self.train_idx = torch.nonzero(self.g.ndata["train_mask"]).squeeze(1)
        self.val_idx = torch.nonzero(self.g.ndata["val_mask"]).squeeze(1)
        self.test_idx = torch.nonzero(self.g.ndata["test_mask"]).squeeze(1)

    def __getitem__(self, idx):
        return self.g

    def __len__(self):
        return 1


class MiniGCDataset(DGLBuiltinDataset):
    """Synthetic graph-classification dataset of 8 graph families (reference data/minigc.py)."""

    def __init__(self, num_graphs, min_num_v, max_num_v, seed=0, transform=None):
        self.num_graphs, self.min_num_v, self.max_num_v, self.seed = num_graphs, min_num_v, max_num_v, seed
        super().__init__("minigc", hash_key=(num_graphs, min_num_v, max_num_v, seed), transform=transform)

    def process(self):
        import networkx as nx
        from ..convert import from_networkx
        import numpy as np
        rng = np.random.RandomState(self.seed)
        gens = [lambda n: nx.cycle_graph(n), lambda n: nx.star_graph(n - 1), lambda n: nx.wheel_graph(n),
                lambda n: nx.lollipop_graph(n // 2, n - n // 2), lambda n: nx.hypercube_graph(max(1, int(np.log2(n)))),
                lambda n: nx.grid_graph([max(2, int(np.sqrt(n)))] * 2), lambda n: nx.complete_graph(n),
                lambda n: nx.circular_ladder_graph(max(2, n // 2))]
        self.graphs, self.labels = [], []
        for i in range(self.num_graphs):
            c = i % 8
            n = rng.randint(self.min_num_v, self.max_num_v)
            nxg = nx.convert_node_labels_to_integers(gens[c](n))
            self.graphs.append(from_networkx(nxg))
            self.labels.append(c)
        self.labels = torch.tensor(self.labels)

    @property
    def num_classes(self):
        return 8

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, idx):
        g = self.graphs[idx]
        return (g if self._transform is None else self._transform(g)), self.labels[idx]
