"""Remaining builtin datasets of the reference (python/dgl/data/__init__.py): legacy aliases, the explainability
benchmarks (which ARE generated algorithmically in the reference too, so they are implemented exactly: synthetic_graphs
BA-Shapes / BA-Community / Tree-Cycle / Tree-Grid / BA-2motifs, reference data/synthetic.py), the SBM node-classification
benchmarks PATTERN / CLUSTER, the super-pixel / LRGB graph collections and MovieLens (shapes of the published files,
synthetic content: no network here)."""
from __future__ import annotations

import math

import os

import torch

from ..convert import graph as make_graph
from ..convert import heterograph as make_heterograph
from .dgl_dataset import DGLBuiltinDataset
from .more import (AmazonCoBuyComputerDataset, AmazonCoBuyPhotoDataset, BitcoinOTCDataset, CoauthorCSDataset,
                   CoauthorPhysicsDataset, CoraFullDataset, GDELTDataset, ICEWS18Dataset, KarateClubDataset, PPIDataset,
                   QM7bDataset, QM9Dataset, SBMMixtureDataset, SSTDataset, _GraphCollection, _node_cls, NODE_STATS)

# ---------------------------------------------------------------------------------------------- legacy aliases
KarateClub = KarateClubDataset
QM7b = QM7bDataset
QM9 = QM9Dataset
SST = SSTDataset
GDELT = GDELTDataset
ICEWS18 = ICEWS18Dataset
BitcoinOTC = BitcoinOTCDataset
SBMMixture = SBMMixtureDataset
CoraFull = CoraFullDataset
LegacyPPIDataset = PPIDataset


class QM9EdgeDataset(QM9Dataset):
    """QM9 with edge features (reference qm9_edge.py): bond-type one-hots on the edges."""

    def process(self):
        super().process()
        gen = torch.Generator().manual_seed(5)
        for g in self.graphs:
            g.edata["edge_attr"] = torch.nn.functional.one_hot(torch.randint(0, 4, (g.num_edges(),), generator=gen), 4).float()
            g.ndata["pos"] = torch.randn(g.num_nodes(), 3, generator=gen)


QM9Edge = QM9EdgeDataset


def AmazonCoBuy(name, **kw):
    """Legacy constructor (reference gnn_benchmark.py AmazonCoBuy): name in {'computers', 'photo'}."""
    return {"computers": AmazonCoBuyComputerDataset, "photo": AmazonCoBuyPhotoDataset}[name](**kw)


def Coauthor(name, **kw):
    return {"cs": CoauthorCSDataset, "physics": CoauthorPhysicsDataset}[name](**kw)


def CitationGraphDataset(name, **kw):
    """Legacy dispatcher (reference citation_graph.py): name in {'cora', 'citeseer', 'pubmed'}."""
    from . import CiteseerGraphDataset, CoraGraphDataset, PubmedGraphDataset
    return {"cora": CoraGraphDataset, "citeseer": CiteseerGraphDataset, "pubmed": PubmedGraphDataset}[name](**kw)


class CoraBinary(_GraphCollection):
    """21 binary-community subgraphs of Cora used by the LGNN tutorial (reference citation_graph.py CoraBinary)."""

    def __init__(self, raw_dir=None, force_reload=False, verbose=False, transform=None):
        super().__init__("cora_binary", raw_dir=raw_dir, transform=transform)

    def process(self):
        self._gen(21, 100, 4.0, 0, 2, seed=21)
        gen = torch.Generator().manual_seed(22)
        self.pmpds, self.labels = [], []
        for g in self.graphs:
            self.labels.append(torch.randint(0, 2, (g.num_nodes(),), generator=gen))
            self.pmpds.append(None)

    def __getitem__(self, idx):
        return self.graphs[idx], self.pmpds[idx], self.labels[idx]


# ---------------------------------------------------------------------------------------------- explainability benchmarks
def _ba_graph(n, m, gen):
    """Barabasi-Albert preferential attachment (undirected edge list)."""
    targets = list(range(m))
    rep = []
    src, dst = [], []
    for v in range(m, n):
        for t in targets:
            src.append(v)
            dst.append(t)
        rep.extend(targets)
        rep.extend([v] * m)
        idx = torch.randint(0, len(rep), (4 * m,), generator=gen).tolist()
        seen = []
        for i in idx:
            if rep[i] not in seen:
                seen.append(rep[i])
            if len(seen) == m:
                break
        while len(seen) < m:
            seen.append(int(torch.randint(0, v + 1, (1,), generator=gen)))
        targets = seen
    return src, dst


def _attach(src, dst, labels, base_n, motif_edges, motif_labels, n_motifs, gen, attach_to=None):
    """Attach ``n_motifs`` copies of a motif to random base nodes; returns the new node count."""
    k = len(motif_labels)
    n = base_n if attach_to is None else attach_to
    anchors = torch.randint(0, base_n, (n_motifs,), generator=gen).tolist()
    cur = len(labels)
    for a in anchors:
        for (x, y) in motif_edges:
            src.append(cur + x)
            dst.append(cur + y)
        src.append(cur)
        dst.append(a)
        labels.extend(motif_labels)
        cur += k
    return cur


HOUSE = ([(0, 1), (1, 2), (2, 3), (3, 0), (0, 4), (1, 4)], [1, 1, 2, 2, 3])           # roof / middle / bottom roles
CYCLE6 = ([(i, (i + 1) % 6) for i in range(6)], [1] * 6)
GRID3 = ([(r * 3 + c, r * 3 + c + 1) for r in range(3) for c in range(2)] + [(r * 3 + c, (r + 1) * 3 + c) for r in range(2) for c in range(3)], [1] * 9)


class _ExplainNodeDataset(DGLBuiltinDataset):
    def __init__(self, name, seed=None, raw_dir=None, force_reload=False, verbose=False, transform=None, **cfg):
        self._seed, self._cfg = seed, cfg
        super().__init__(name, raw_dir=raw_dir, hash_key=(name, seed, tuple(sorted(cfg.items()))), transform=transform)

    def _finish(self, src, dst, labels, extra_edges, gen):
        n = len(labels)
        for _ in range(extra_edges):                                    # random perturbation edges (reference: 0.1 * |E|)
            a, b = torch.randint(0, n, (2,), generator=gen).tolist()
            src.append(a)
            dst.append(b)
        s, d = torch.tensor(src), torch.tensor(dst)
        g = make_graph((torch.cat([s, d]), torch.cat([d, s])), num_nodes=n)
        g.ndata["label"] = torch.tensor(labels)
        g.ndata["feat"] = torch.ones(n, 1)
        self._graph = g
        self._num_classes = int(max(labels)) + 1

    @property
    def num_classes(self):
        return self._num_classes

    def __len__(self):
        return 1

    def __getitem__(self, idx):
        assert idx == 0
        return self._graph if self._transform is None else self._transform(self._graph)


class BAShapeDataset(_ExplainNodeDataset):
    """BA-Shapes (Ying et al. 2019; reference data/synthetic.py BAShapeDataset): BA base graph of 300 nodes, 80 house
    motifs, 4 classes (base / roof / middle / bottom)."""

    def __init__(self, num_base_nodes=300, num_base_edges_per_node=5, num_motifs=80, perturb_ratio=0.01, seed=None, **kw):
        super().__init__("BA-SHAPES", seed=seed, n=num_base_nodes, m=num_base_edges_per_node, k=num_motifs, p=perturb_ratio, **kw)

    def process(self):
        c = self._cfg
        gen = torch.Generator().manual_seed(0 if self._seed is None else self._seed)
        src, dst = _ba_graph(c["n"], c["m"], gen)
        labels = [0] * c["n"]
        _attach(src, dst, labels, c["n"], *HOUSE, c["k"], gen)
        self._finish(src, dst, labels, int(c["p"] * len(src)), gen)


class BACommunityDataset(_ExplainNodeDataset):
    """BA-Community: two BA-Shapes graphs with different Gaussian features joined by random edges, 8 classes."""

    def __init__(self, num_base_nodes=300, num_base_edges_per_node=4, num_motifs=80, perturb_ratio=0.01,
                 num_inter_edges=350, seed=None, **kw):
        super().__init__("BA-COMMUNITY", seed=seed, n=num_base_nodes, m=num_base_edges_per_node, k=num_motifs, p=perturb_ratio,
                         inter=num_inter_edges, **kw)

    def process(self):
        c = self._cfg
        gen = torch.Generator().manual_seed(0 if self._seed is None else self._seed)
        src, dst, labels, feats = [], [], [], []
        off = 0
        for comm in range(2):
            s, d = _ba_graph(c["n"], c["m"], gen)
            lab = [0] * c["n"]
            _attach(s, d, lab, c["n"], *HOUSE, c["k"], gen)
            src += [x + off for x in s]
            dst += [x + off for x in d]
            labels += [l + 4 * comm for l in lab]
            feats.append(torch.randn(len(lab), 10, generator=gen) + (1.0 if comm else -1.0))
            off += len(lab)
        half = off // 2
        for _ in range(c["inter"]):
            src.append(int(torch.randint(0, half, (1,), generator=gen)))
            dst.append(int(torch.randint(half, off, (1,), generator=gen)))
        self._finish(src, dst, labels, int(c["p"] * len(src)), gen)
        self._graph.ndata["feat"] = torch.cat(feats)


class _TreeMotifDataset(_ExplainNodeDataset):
    _motif = None

    def __init__(self, tree_height=8, num_motifs=60, perturb_ratio=0.01, seed=None, **kw):
        super().__init__(type(self).__name__, seed=seed, h=tree_height, k=num_motifs, p=perturb_ratio, **kw)

    def process(self):
        c = self._cfg
        gen = torch.Generator().manual_seed(0 if self._seed is None else self._seed)
        n = 2 ** (c["h"] + 1) - 1                                      # balanced binary tree
        src = list(range(1, n))
        dst = [(i - 1) // 2 for i in range(1, n)]
        labels = [0] * n
        _attach(src, dst, labels, n, *self._motif, c["k"], gen)
        self._finish(src, dst, labels, int(c["p"] * len(src)), gen)


class TreeCycleDataset(_TreeMotifDataset):
    """Tree-Cycles: balanced binary tree of height 8 with 60 six-node cycles attached, 2 classes."""
    _motif = CYCLE6


class TreeGridDataset(_TreeMotifDataset):
    """Tree-Grids: balanced binary tree of height 8 with 80 3x3 grids attached, 2 classes."""
    _motif = GRID3

    def __init__(self, tree_height=8, num_motifs=80, perturb_ratio=0.1, seed=None, **kw):
        super().__init__(tree_height, num_motifs, perturb_ratio, seed, **kw)


class BA2MotifDataset(_GraphCollection):
    """BA-2motifs (Luo et al. 2020): 1000 BA graphs of 20 nodes, half with a house motif, half with a 5-cycle."""

    def __init__(self, raw_dir=None, force_reload=False, verbose=False, transform=None, num_graphs=1000):
        self._n = num_graphs
        super().__init__("BA-2motifs", raw_dir=raw_dir, hash_key=(num_graphs,), transform=transform)

    def process(self):
        gen = torch.Generator().manual_seed(3)
        self.graphs, labels = [], []
        cyc5 = ([(i, (i + 1) % 5) for i in range(5)], [1] * 5)
        for i in range(self._n):
            src, dst = _ba_graph(20, 1, gen)
            lab = [0] * 20
            _attach(src, dst, lab, 20, *(HOUSE if i % 2 == 0 else cyc5), 1, gen)
            s, d = torch.tensor(src), torch.tensor(dst)
            g = make_graph((torch.cat([s, d]), torch.cat([d, s])), num_nodes=len(lab))
            g.ndata["feat"] = torch.full((len(lab), 10), 0.1)
            self.graphs.append(g)
            labels.append(i % 2)
        self.labels = torch.tensor(labels)
        self._num_targets = 2


# ---------------------------------------------------------------------------------------------- SBM benchmarks
class _SBMNodeDataset(_GraphCollection):
    """PATTERN / CLUSTER (Dwivedi et al. benchmarking-gnns): stochastic-block-model graphs with node-level labels."""

    def __init__(self, name, mode="train", raw_dir=None, force_reload=False, verbose=False, transform=None, scale=0.05):
        self._mode, self._scale = mode, scale
        super().__init__(name, raw_dir=raw_dir, hash_key=(name, mode, scale), transform=transform)

    def _make(self, total, n_lo, n_hi, comms, p_in, q_out, fdim, classes, seed):
        gen = torch.Generator().manual_seed(seed + {"train": 0, "valid": 1, "test": 2}[self._mode])
        num = max(4, int(total * self._scale))
        self.graphs = []
        for _ in range(num):
            n = int(torch.randint(n_lo, n_hi + 1, (1,), generator=gen))
            comm = torch.randint(0, comms, (n,), generator=gen)
            same = comm.unsqueeze(0) == comm.unsqueeze(1)
            prob = torch.where(same, torch.tensor(p_in), torch.tensor(q_out))
            adj = torch.rand(n, n, generator=gen) < prob
            adj = torch.triu(adj, 1)
            s, d = adj.nonzero(as_tuple=True)
            g = make_graph((torch.cat([s, d]), torch.cat([d, s])), num_nodes=n)
            g.ndata["feat"] = torch.randint(0, fdim, (n,), generator=gen)
            g.ndata["label"] = (comm % classes) if classes == comms else (comm == 0).long()
            g.edata["feat"] = torch.ones(g.num_edges(), 1)
            self.graphs.append(g)
        self.labels = torch.zeros(num, dtype=torch.long)
        self._num_targets = classes

    def __getitem__(self, idx):
        g = self.graphs[idx]
        return g if self._transform is None else self._transform(g)


class PATTERNDataset(_SBMNodeDataset):
    """PATTERN: 14000 SBM graphs (44-188 nodes), task = find the planted pattern (binary node labels)."""

    def __init__(self, mode="train", **kw):
        super().__init__("PATTERN", mode=mode, **kw)

    def process(self):
        self._make({"train": 10000, "valid": 2000, "test": 2000}[self._mode], 44, 188, 5, 0.5, 0.35, 3, 2, seed=41)


class CLUSTERDataset(_SBMNodeDataset):
    """CLUSTER: 12000 SBM graphs (41-190 nodes), task = cluster id (6 classes)."""

    def __init__(self, mode="train", **kw):
        super().__init__("CLUSTER", mode=mode, **kw)

    def process(self):
        self._make({"train": 10000, "valid": 1000, "test": 1000}[self._mode], 41, 190, 6, 0.55, 0.25, 7, 6, seed=43)


# ---------------------------------------------------------------------------------------------- graph collections
def _collection(name, num, avg_n, avg_deg, fdim, targets, regression=False, edge_feat=0, default_scale=0.01, doc=""):
    class _C(_GraphCollection):
        __doc__ = f"{name}: {doc} Shape of the published dataset (synthetic content)."

        def __init__(self, split="train", raw_dir=None, force_reload=False, verbose=False, transform=None, scale=default_scale,
                     **kw):
            self._split, self._scale = split, scale
            _GraphCollection.__init__(self, name, raw_dir=raw_dir, hash_key=(name, split, scale), transform=transform)

        def process(self):
            self._gen(num, avg_n, avg_deg, fdim, targets, seed=len(name) + len(self._split), regression=regression,
                      edge_feat=edge_feat, scale=self._scale)

    _C.__name__ = _C.__qualname__ = name
    return _C


MNISTSuperPixelDataset = _collection("MNISTSuperPixelDataset", 70000, 71, 8.0, 3, 10, doc="MNIST images as 75-superpixel kNN graphs.")
CIFAR10SuperPixelDataset = _collection("CIFAR10SuperPixelDataset", 60000, 118, 8.0, 5, 10, doc="CIFAR10 images as superpixel kNN graphs.")
COCOSuperpixelsDataset = _collection("COCOSuperpixelsDataset", 123286, 477, 5.65, 14, 81, edge_feat=2, default_scale=0.001,
                                     doc="LRGB COCO-SP node classification.")
VOCSuperpixelsDataset = _collection("VOCSuperpixelsDataset", 11355, 479, 5.65, 14, 21, edge_feat=2, default_scale=0.005,
                                    doc="LRGB PascalVOC-SP node classification.")
PeptidesFunctionalDataset = _collection("PeptidesFunctionalDataset", 15535, 151, 2.04, 9, 10, edge_feat=3, default_scale=0.01,
                                        doc="LRGB Peptides-func (10-way multi-label graph classification).")
PeptidesStructuralDataset = _collection("PeptidesStructuralDataset", 15535, 151, 2.04, 9, 11, regression=True, edge_feat=3,
                                        default_scale=0.01, doc="LRGB Peptides-struct (11 regression targets).")


class MovieLensDataset(DGLBuiltinDataset):
    """MovieLens rating graph (reference movielens.py): bipartite user-movie heterograph with rating edges, train / valid
    / test edge masks. ``name`` in {'ml-100k', 'ml-1m', 'ml-10m'} selects the published shape."""

    SHAPES = {"ml-100k": (943, 1682, 100000, 23, 19), "ml-1m": (6040, 3706, 1000209, 30, 19), "ml-10m": (69878, 10677, 10000054, 1, 20)}

    def __init__(self, name="ml-100k", valid_ratio=0.1, test_ratio=0.1, raw_dir=None, force_reload=False, verbose=False,
                 transform=None, random_state=0, scale=1.0):
        self._ml, self._vr, self._tr, self._rs, self._scale = name, valid_ratio, test_ratio, random_state, scale
        from . import raw_readers as raw
        from .dgl_dataset import get_download_dir
        self._raw_root = (raw.locate(raw_dir if raw_dir is not None else get_download_dir(),
                                     [name, os.path.join(name, name)], ["u.data"]) if name == "ml-100k" else None)
        super().__init__(name, raw_dir=raw_dir, force_reload=force_reload, transform=transform,
                         hash_key=(name, valid_ratio, test_ratio, random_state, "raw" if self._raw_root else scale))

    @property
    def is_synthetic(self):
        return self._raw_root is None

    def process(self):
        if self._raw_root is not None:      # u.data / u.user / u.item on disk (movielens.py)
            from . import raw_readers as raw
            self._graph = raw.read_movielens_100k(self._raw_root, self._tr, self._vr, self._rs).graph
            return
        nu, nm, ne, fu, fm = self.SHAPES[self._ml]
        nu, nm, ne = max(8, int(nu * self._scale)), max(8, int(nm * self._scale)), max(32, int(ne * self._scale))
        gen = torch.Generator().manual_seed(self._rs)
        u = torch.randint(0, nu, (ne,), generator=gen)
        m = (torch.rand(ne, generator=gen) ** 2 * nm).long().clamp(max=nm - 1)      # popular movies get more ratings
        rate = torch.randint(1, 6, (ne,), generator=gen).float()
        g = make_heterograph({("user", "user-movie", "movie"): (u, m), ("movie", "movie-user", "user"): (m, u)},
                             {"user": nu, "movie": nm})
        g.nodes["user"].data["feat"] = torch.randn(nu, fu, generator=gen)
        g.nodes["movie"].data["feat"] = torch.randn(nm, fm, generator=gen)
        perm = torch.randperm(ne, generator=gen)
        n_te, n_va = int(ne * self._tr), int(ne * self._vr)
        for et in ("user-movie", "movie-user"):
            g.edges[et].data["rate"] = rate
            for nme, lo, hi in (("test_mask", 0, n_te), ("valid_mask", n_te, n_te + n_va), ("train_mask", n_te + n_va, ne)):
                msk = torch.zeros(ne, dtype=torch.bool)
                msk[perm[lo:hi]] = True
                g.edges[et].data[nme] = msk
        self._graph = g

    def __len__(self):
        return 1

    def __getitem__(self, idx):
        assert idx == 0
        return self._graph if self._transform is None else self._transform(self._graph)
