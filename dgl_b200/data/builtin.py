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


def _undirected(n, pairs):
    """Graph on ``n`` nodes with both directions of every pair plus one self loop per node."""
    import numpy as np
    a = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    u = np.concatenate([a[:, 0], a[:, 1], np.arange(n)])
    v = np.concatenate([a[:, 1], a[:, 0], np.arange(n)])
    return make_graph((torch.as_tensor(u), torch.as_tensor(v)), num_nodes=n)


def _cycle(n):
    return [(i, (i + 1) % n) for i in range(n)] if n > 2 else [(0, 1)][: n - 1]


# the eight families of MiniGC as (label order) edge-list builders; every builder returns (num_nodes, undirected pairs)
def _family_cycle(n, draw):
    return n, _cycle(n)


def _family_star(n, draw):
    return n, [(0, i) for i in range(1, n)]


def _family_wheel(n, draw):
    rim = [(1 + i, 1 + (i + 1) % (n - 1)) for i in range(n - 1)] if n > 3 else [(1, 2)][: max(n - 2, 0)]
    return n, [(0, i) for i in range(1, n)] + rim


def _family_lollipop(n, draw):
    tail = int(draw(2, n // 2))
    head = n - tail
    clique = [(i, j) for i in range(head) for j in range(i + 1, head)]
    return n, clique + [(head - 1 + i, head + i) for i in range(tail)]


def _family_hypercube(n, draw):
    import math
    d = int(math.log(n, 2))
    return 1 << d, [(i, i ^ (1 << k)) for i in range(1 << d) for k in range(d) if i < i ^ (1 << k)]


def _family_grid(n, draw):
    rows = int(draw(2, n // 2))
    cols = n // rows
    node = lambda r, c: c * rows + r  # noqa: E731 - the first dimension varies fastest in networkx's labelling
    pairs = [(node(r, c), node(r + 1, c)) for r in range(rows - 1) for c in range(cols)]
    pairs += [(node(r, c), node(r, c + 1)) for r in range(rows) for c in range(cols - 1)]
    return rows * cols, pairs


def _family_clique(n, draw):
    return n, [(i, j) for i in range(n) for j in range(i + 1, n)]


def _family_ladder(n, draw):
    k = n // 2
    return 2 * k, _cycle(k) + [(k + a, k + b) for a, b in _cycle(k)] + [(i, k + i) for i in range(k)]


MINIGC_FAMILIES = (_family_cycle, _family_star, _family_wheel, _family_lollipop, _family_hypercube, _family_grid,
                   _family_clique, _family_ladder)


class MiniGCDataset(DGLBuiltinDataset):
    """Synthetic graph classification over eight graph families (reference data/minigc.py): cycle, star, wheel,
    lollipop, hypercube, grid, clique, circular ladder - ``num_graphs // 8`` of each in this order (the last family
    takes the remainder), sizes drawn from ``[min_num_v, max_num_v)`` by one seeded stream, self loops added.

    The draws replay the reference's sequence (same seed -> same sizes, tail lengths and grid shapes) and the edge
    lists are written out directly with networkx's node numbering, so graphs are identical to the reference's without
    building networkx objects (``tests/test_raw_readers.py`` compares them)."""

    def __init__(self, num_graphs, min_num_v, max_num_v, seed=0, save_graph=True, force_reload=False, verbose=False,
                 transform=None):
        self.num_graphs, self.min_num_v, self.max_num_v, self.seed = num_graphs, min_num_v, max_num_v, seed
        self.save_graph = save_graph
        super().__init__("minigc", hash_key=(num_graphs, min_num_v, max_num_v, seed), force_reload=force_reload,
                         verbose=verbose, transform=transform)

    def process(self):
        import numpy as np
        rng = np.random.RandomState(self.seed) if self.seed is not None else np.random.RandomState()
        share = self.num_graphs // 8
        counts = [share] * 7 + [self.num_graphs - 7 * share]
        self.graphs, labels = [], []
        for label, (family, count) in enumerate(zip(MINIGC_FAMILIES, counts)):
            for _ in range(count):
                n = int(rng.randint(self.min_num_v, self.max_num_v))
                if family is _family_grid and n < 4:
                    raise ValueError(f"a grid needs at least 4 nodes, got {n}")
                self.graphs.append(_undirected(*family(n, rng.randint)))
                labels.append(label)
        self.labels = torch.tensor(labels, dtype=torch.int64)

    def has_cache(self):
        return os.path.exists(os.path.join(self.save_path, f"dgl_graph_{self.hash}.bin"))

    def save(self):
        if self.save_graph:
            save_graphs(os.path.join(self.save_path, f"dgl_graph_{self.hash}.bin"), self.graphs, {"labels": self.labels})

    def load(self):
        self.graphs, lab = load_graphs(os.path.join(self.save_path, f"dgl_graph_{self.hash}.bin"))
        self.labels = lab["labels"]

    @property
    def num_classes(self):
        return 8

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, idx):
        g = self.graphs[idx]
        return (g if self._transform is None else self._transform(g)), self.labels[idx]
