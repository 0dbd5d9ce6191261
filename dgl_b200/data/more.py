"""The rest of the reference's builtin dataset zoo (python/dgl/data/{citation_graph,gnn_benchmark,ppi,tu,gindt,qm9,
zinc,knowledge_graph,rdf,sbm,karate,fraud,fakenews,flickr,yelp,wikics,tree,bitcoinotc,gdelt,icews18,qm7b}.py).

There is no network here, so nothing can be downloaded: every class first looks for already-processed data under
``raw_dir`` (``<name>.bin`` written by ``save_graphs``) and otherwise synthesises a dataset with the PUBLISHED shape
(node / edge / feature / class / graph counts from the reference's docstrings), keeping the reference's attribute surface.
Karate club and the SBM mixture are exact (they are generated, not downloaded, in the reference too)."""
from __future__ import annotations

import math
import os

import torch

from ..convert import graph as make_graph
from ..convert import heterograph as make_heterograph
from .dgl_dataset import DGLBuiltinDataset, get_download_dir
from .graph_serialize import load_graphs, save_graphs
from . import raw_readers as raw
from .raw_readers import find_node_source, load_node_source
from .synthetic import rmat_edges

# ------------------------------------------------------------------ single-graph node classification
# name: (nodes, directed edges, feat dim, classes, multilabel)
NODE_STATS = {
    "citeseer": (3327, 9228, 3703, 6, False), "pubmed": (19717, 88651, 500, 3, False),
    "cora_full": (19793, 126842, 8710, 70, False), "amazon_co_buy_computer": (13752, 491722, 767, 10, False),
    "amazon_co_buy_photo": (7650, 238163, 745, 8, False), "coauthor_cs": (18333, 163788, 6805, 15, False),
    "coauthor_physics": (34493, 495924, 8415, 5, False), "wiki_cs": (11701, 431726, 300, 10, False),
    "flickr": (89250, 899756, 500, 7, False), "yelp": (716847, 13954819, 300, 100, True),
    "chameleon": (2277, 36101, 2325, 5, False), "squirrel": (5201, 217073, 2089, 5, False),
    "cornell": (183, 298, 1703, 5, False), "texas": (183, 325, 1703, 5, False), "wisconsin": (251, 515, 1703, 5, False),
    "actor": (7600, 33391, 932, 5, False), "roman_empire": (22662, 65854, 300, 18, False),
    "amazon_ratings": (24492, 186100, 300, 5, False), "minesweeper": (10000, 78804, 7, 2, False),
    "tolokers": (11758, 1038000, 10, 2, False), "questions": (48921, 307080, 301, 2, False),
}


class _StatNodeDataset(DGLBuiltinDataset):
    _key = None
    _default_scale = 1.0

    def __init__(self, raw_dir=None, force_reload=False, verbose=False, transform=None, scale=None, **kw):
        self._scale = self._default_scale if scale is None else scale
        self._raw_options = {k: kw[k] for k in ("reverse_edge", "reorder") if k in kw}
        self._raw_root = find_node_source(self._key, raw_dir if raw_dir is not None else get_download_dir())
        key = (self._key, "raw", tuple(sorted(self._raw_options.items()))) if self._raw_root else (self._key, self._scale)
        super().__init__(self._key, raw_dir=raw_dir, hash_key=key, force_reload=force_reload,
                         verbose=verbose, transform=transform)

    @property
    def is_synthetic(self):
        """False when the graph was parsed from the reference's raw files found under ``raw_dir``."""
        return self._raw_root is None

    def process(self):
        if self._raw_root is not None:          # the reference's downloaded files are on disk: parse them
            parsed = load_node_source(self._key, self._raw_root, **self._raw_options)
            self._g, self._num_classes = parsed.graph, parsed.num_classes
            return
        n, e, f, c, multi = NODE_STATS[self._key]
        n, e = max(32, int(n * self._scale)), max(64, int(e * self._scale))
        u, v = rmat_edges(n, e // 2, seed=hash(self._key) % 1000)
        g = make_graph((torch.cat([u, v]), torch.cat([v, u])), num_nodes=n)
        gen = torch.Generator().manual_seed(7)
        g.ndata["feat"] = torch.randn(n, f, generator=gen)
        g.ndata["label"] = ((torch.rand(n, c, generator=gen) < 0.1).float() if multi
                            else torch.randint(0, c, (n,), generator=gen))
        perm = torch.randperm(n, generator=gen)
        for name, lo, hi in (("train_mask", 0, 0.5), ("val_mask", 0.5, 0.75), ("test_mask", 0.75, 1.0)):
            m = torch.zeros(n, dtype=torch.bool)
            m[perm[int(lo * n): int(hi * n)]] = True
            g.ndata[name] = m
        self._g, self._num_classes = g, c

    def has_cache(self):
        return os.path.exists(os.path.join(self.save_path, "graph.bin"))

    def save(self):
        save_graphs(os.path.join(self.save_path, "graph.bin"), [self._g], {"num_classes": torch.tensor([self._num_classes])})

    def load(self):
        gs, lab = load_graphs(os.path.join(self.save_path, "graph.bin"))
        self._g, self._num_classes = gs[0], int(lab["num_classes"][0])

    @property
    def num_classes(self):
        return self._num_classes

    num_labels = num_classes

    def __getitem__(self, idx):
        assert idx == 0, "This dataset has only one graph"
        return self._g if self._transform is None else self._transform(self._g)

    def __len__(self):
        return 1


def _node_cls(name, key, scale=1.0):
    return type(name, (_StatNodeDataset,), {"_key": key, "_default_scale": scale, "__doc__": f"{name}: shape of the "
                                            f"reference dataset {NODE_STATS[key][:4]} (synthetic content)."})


CiteseerGraphDataset = _node_cls("CiteseerGraphDataset", "citeseer")
PubmedGraphDataset = _node_cls("PubmedGraphDataset", "pubmed")
CoraFullDataset = _node_cls("CoraFullDataset", "cora_full")
AmazonCoBuyComputerDataset = _node_cls("AmazonCoBuyComputerDataset", "amazon_co_buy_computer")
AmazonCoBuyPhotoDataset = _node_cls("AmazonCoBuyPhotoDataset", "amazon_co_buy_photo")
CoauthorCSDataset = _node_cls("CoauthorCSDataset", "coauthor_cs")
CoauthorPhysicsDataset = _node_cls("CoauthorPhysicsDataset", "coauthor_physics")
WikiCSDataset = _node_cls("WikiCSDataset", "wiki_cs")
FlickrDataset = _node_cls("FlickrDataset", "flickr")
YelpDataset = _node_cls("YelpDataset", "yelp", 0.02)
ChameleonDataset = _node_cls("ChameleonDataset", "chameleon")
SquirrelDataset = _node_cls("SquirrelDataset", "squirrel")
CornellDataset = _node_cls("CornellDataset", "cornell")
TexasDataset = _node_cls("TexasDataset", "texas")
WisconsinDataset = _node_cls("WisconsinDataset", "wisconsin")
ActorDataset = _node_cls("ActorDataset", "actor")
RomanEmpireDataset = _node_cls("RomanEmpireDataset", "roman_empire")
AmazonRatingsDataset = _node_cls("AmazonRatingsDataset", "amazon_ratings")
MinesweeperDataset = _node_cls("MinesweeperDataset", "minesweeper")
TolokersDataset = _node_cls("TolokersDataset", "tolokers")
QuestionsDataset = _node_cls("QuestionsDataset", "questions")


class KarateClubDataset(DGLBuiltinDataset):
    """Zachary's karate club - the real graph (34 nodes, 78 undirected edges, 2 communities), via networkx."""

    def __init__(self, transform=None):
        super().__init__("karate_club", transform=transform)

    def process(self):
        import networkx as nx
        from ..convert import from_networkx
        nxg = nx.karate_club_graph()
        g = from_networkx(nxg)
        g.ndata["label"] = torch.tensor([0 if nxg.nodes[i]["club"] == "Mr. Hi" else 1 for i in range(34)])
        self._g = g

    num_classes = 2

    def __getitem__(self, idx):
        assert idx == 0
        return self._g if self._transform is None else self._transform(self._g)

    def __len__(self):
        return 1


# ------------------------------------------------------------------ collections of graphs
# name: (num graphs, avg nodes, avg degree, node feat dim, classes)
GRAPH_STATS = {
    "ENZYMES": (600, 32.6, 3.8, 18, 6), "PROTEINS": (1113, 39.1, 3.7, 3, 2), "MUTAG": (188, 17.9, 2.2, 7, 2),
    "DD": (1178, 284.3, 5.0, 89, 2), "COLLAB": (5000, 74.5, 65.9, 0, 3), "IMDB-BINARY": (1000, 19.8, 9.8, 0, 2),
    "IMDB-MULTI": (1500, 13.0, 10.1, 0, 3), "NCI1": (4110, 29.9, 2.2, 37, 2), "PTC_MR": (344, 14.3, 2.1, 18, 2),
    "REDDIT-BINARY": (2000, 429.6, 2.3, 0, 2), "REDDIT-MULTI-5K": (4999, 508.5, 2.3, 0, 5),
}


class _GraphCollection(DGLBuiltinDataset):
    """Base of the multi-graph datasets: ``dataset[i] -> (graph, label)``."""
    _raw_root = None     # set by subclasses that found the reference's raw files under ``raw_dir``

    @staticmethod
    def _find_raw(raw_dir, names, required=()):
        return raw.locate(raw_dir if raw_dir is not None else get_download_dir(), names, required)

    @property
    def is_synthetic(self):
        """False when the graphs were parsed from the reference's raw files found under ``raw_dir``."""
        return self._raw_root is None

    def _adopt(self, parsed):
        """Take over a parsed ``RawCollection``: graphs, labels and its dataset-level attributes."""
        self.graphs = parsed.graphs
        self.labels = parsed.labels if parsed.labels is not None else torch.zeros(len(parsed.graphs), dtype=torch.long)
        for k, v in parsed.extra.items():
            if isinstance(getattr(type(self), k, None), property):     # num_labels / num_classes read _num_targets
                self._num_targets = v
            else:
                setattr(self, k, v)

    def _gen(self, num, avg_n, avg_deg, fdim, num_targets, seed, regression=False, edge_feat=0, scale=1.0):
        gen = torch.Generator().manual_seed(seed)
        num = max(4, int(num * scale))
        self.graphs, labels = [], []
        for i in range(num):
            n = max(2, int(torch.normal(avg_n, avg_n * 0.25, (1,), generator=gen).item()))
            m = max(1, int(n * avg_deg / 2))
            u = torch.randint(0, n, (m,), generator=gen)
            v = torch.randint(0, n, (m,), generator=gen)
            g = make_graph((torch.cat([u, v]), torch.cat([v, u])), num_nodes=n)
            if fdim:
                g.ndata["feat"] = torch.randn(n, fdim, generator=gen)
                g.ndata["attr"] = g.ndata["feat"]
            if edge_feat:
                g.edata["feat"] = torch.randint(0, edge_feat, (g.num_edges(),), generator=gen)
            self.graphs.append(g)
        self.labels = (torch.randn(num, num_targets, generator=gen) if regression
                       else torch.randint(0, num_targets, (num,), generator=gen))
        self._num_targets = num_targets

    @property
    def num_classes(self):
        return self._num_targets

    num_labels = num_classes

    def __len__(self):
        return len(self.graphs)

    def __getitem__(self, idx):
        g = self.graphs[idx]
        return (g if self._transform is None else self._transform(g)), self.labels[idx]


class TUDataset(_GraphCollection):
    """TU Dortmund graph-kernel benchmarks. With ``<raw_dir>/<name>/<name>_A.txt`` & co on disk (the reference's
    extracted download, tu.py) the real files are parsed; otherwise a stand-in with the published shape is generated."""
    _legacy = False

    def __init__(self, name, raw_dir=None, force_reload=False, verbose=False, transform=None, scale=1.0, **kw):
        self._tu, self._scale = name, scale
        self._raw_root = self._find_raw(raw_dir, [name, os.path.join(name, name)], [name + "_A.txt"])
        if self._raw_root is None and name not in GRAPH_STATS:
            raise KeyError(f"TUDataset {name!r}: no raw files under raw_dir and no known shape ({sorted(GRAPH_STATS)})")
        key = (name, "raw", self._legacy) if self._raw_root else (name, scale)
        super().__init__("tu_" + name, raw_dir=raw_dir, hash_key=key, force_reload=force_reload, transform=transform)

    def process(self):
        if self._raw_root is not None:
            self._adopt(raw.read_tu(self._raw_root, self._tu, legacy=self._legacy))
            self.graph_lists, self.graph_labels = self.graphs, self.labels
            return
        num, an, ad, f, c = GRAPH_STATS[self._tu]
        self._gen(num, an, ad, f, c, seed=len(self._tu), scale=self._scale)
        self.graph_labels = self.labels.unsqueeze(1)
        self.max_num_node = max(g.num_nodes() for g in self.graphs)


class LegacyTUDataset(TUDataset):
    """The older TU loader (tu.py:LegacyTUDataset): every graph carries a ``feat`` matrix (one-hot node label, node
    attributes or a constant column) and ``node_label``."""
    _legacy = True

    def __init__(self, name, use_pandas=False, hidden_size=10, max_allow_node=None, raw_dir=None, force_reload=False,
                 verbose=False, transform=None, scale=1.0):
        self.hidden_size, self.max_allow_node = hidden_size, max_allow_node
        super().__init__(name, raw_dir=raw_dir, force_reload=force_reload, verbose=verbose, transform=transform, scale=scale)

    def process(self):
        super().process()
        if self._raw_root is not None and self.max_allow_node:
            keep = [i for i, g in enumerate(self.graphs) if g.num_nodes() <= self.max_allow_node]
            self.graphs = self.graph_lists = [self.graphs[i] for i in keep]
            self.labels = self.graph_labels = self.labels[torch.tensor(keep, dtype=torch.long)]
            self.max_num_node = max(g.num_nodes() for g in self.graphs)

    def statistics(self):
        return self.graphs[0].ndata["feat"].shape[1], self.num_classes, self.max_num_node


class GINDataset(_GraphCollection):
    """GIN benchmark variants of the TU sets; node features are one-hot degrees when the source has none."""

    def __init__(self, name, self_loop=False, degree_as_nlabel=False, raw_dir=None, force_reload=False, verbose=False,
                 transform=None, scale=1.0):
        key = {"PTC": "PTC_MR", "IMDBBINARY": "IMDB-BINARY", "IMDBMULTI": "IMDB-MULTI", "REDDITBINARY": "REDDIT-BINARY",
               "REDDITMULTI5K": "REDDIT-MULTI-5K"}.get(name, name)
        self._gin, self._self_loop, self._deg, self._scale = key, self_loop, degree_as_nlabel, scale
        self._raw_root = self._find_raw(raw_dir, [os.path.join("GINDataset", "dataset", name, name + ".txt"),
                                                  os.path.join("dataset", name, name + ".txt"),
                                                  os.path.join(name, name + ".txt")])
        hk = (name, self_loop, degree_as_nlabel, "raw" if self._raw_root else scale)
        super().__init__("gin_" + name, raw_dir=raw_dir, hash_key=hk, force_reload=force_reload, transform=transform)

    def process(self):
        if self._raw_root is not None:     # the "How powerful are GNNs" text file is on disk (gindt.py)
            self._adopt(raw.read_gin_text(self._raw_root, self._self_loop, self._deg))
            self._num_targets = self.gclasses
            return
        num, an, ad, f, c = GRAPH_STATS[self._gin]
        self._gen(num, an, ad, f, c, seed=11, scale=self._scale)
        from ..transforms import add_self_loop
        for i, g in enumerate(self.graphs):
            if self._deg or "attr" not in g.ndata:
                d = g.in_degrees().clamp(max=63)
                g.ndata["attr"] = torch.nn.functional.one_hot(d.long(), 64).float()
            if self._self_loop:
                self.graphs[i] = add_self_loop(g)
        self.gclasses = c
        self.dim_nfeats = self.graphs[0].ndata["attr"].shape[1]


class QM9Dataset(_GraphCollection):
    def __init__(self, label_keys=("mu",), cutoff=5.0, raw_dir=None, force_reload=False, verbose=False, transform=None,
                 scale=0.01):
        self.label_keys, self.cutoff, self._scale = list(label_keys), cutoff, scale
        self._raw_root = self._find_raw(raw_dir, ["qm9_eV.npz", os.path.join("qm9", "qm9_eV.npz")])
        hk = (tuple(label_keys), cutoff, "raw") if self._raw_root else (tuple(label_keys), scale)
        super().__init__("qm9", raw_dir=raw_dir, hash_key=hk, force_reload=force_reload, transform=transform)

    def process(self):
        if self._raw_root is not None:     # qm9_eV.npz on disk: molecules are built lazily on access (qm9.py)
            self._mols = raw.QM9Molecules(self._raw_root, self.label_keys, self.cutoff)
            self.graphs, self.labels = self._mols, torch.as_tensor(self._mols.label.astype("float32"))
            self._num_targets = len(self.label_keys)
            return
        self._gen(130831, 18.0, 8.0, 0, len(self.label_keys), seed=9, regression=True, scale=self._scale)
        gen = torch.Generator().manual_seed(10)
        for g in self.graphs:
            g.ndata["R"] = torch.randn(g.num_nodes(), 3, generator=gen)
            g.ndata["Z"] = torch.randint(1, 10, (g.num_nodes(),), generator=gen)

    @property
    def num_tasks(self):
        return len(self.label_keys)

    def __getitem__(self, idx):
        if self._raw_root is None:
            return super().__getitem__(idx)
        g, y = self._mols[idx]
        return (g if self._transform is None else self._transform(g)), y


QM9EdgeDataset = QM9Dataset


class QM7bDataset(_GraphCollection):
    def __init__(self, raw_dir=None, force_reload=False, verbose=False, transform=None, scale=0.1):
        self._scale = scale
        self._raw_root = self._find_raw(raw_dir, ["qm7b.mat", os.path.join("qm7b", "qm7b.mat")])
        super().__init__("qm7b", raw_dir=raw_dir, hash_key=("raw",) if self._raw_root else (scale,),
                         force_reload=force_reload, transform=transform)

    def process(self):
        if self._raw_root is not None:
            self._adopt(raw.read_qm7b(self._raw_root))
            self._num_targets = self.num_tasks
            return
        self._gen(7211, 15.4, 14.0, 0, 14, seed=12, regression=True, scale=self._scale)
        for g in self.graphs:
            g.edata["h"] = torch.rand(g.num_edges(), 1)


class ZINCDataset(_GraphCollection):
    def __init__(self, mode="train", raw_dir=None, force_reload=False, verbose=False, transform=None, scale=0.01):
        self.mode, self._scale = mode, scale
        super().__init__("zinc_" + mode, raw_dir=raw_dir, hash_key=(mode, scale), transform=transform)

    def process(self):
        num = {"train": 220011, "valid": 24445, "test": 5000}[self.mode]
        self._gen(num, 23.2, 2.1, 0, 1, seed=13, regression=True, edge_feat=4, scale=self._scale)
        gen = torch.Generator().manual_seed(14)
        for g in self.graphs:
            g.ndata["feat"] = torch.randint(0, 28, (g.num_nodes(),), generator=gen)
        self.labels = self.labels.squeeze(1)

    num_atom_types, num_bond_types = 28, 4


class PPIDataset(_GraphCollection):
    """24 protein graphs (20 train / 2 valid / 2 test), 50 features, 121 binary labels per node."""

    def __init__(self, mode="train", raw_dir=None, force_reload=False, verbose=False, transform=None, scale=0.2):
        self.mode, self._scale = mode, scale
        self._raw_root = self._find_raw(raw_dir, ["ppi"], [f"{mode}_graph.json", f"{mode}_feats.npy",
                                                           f"{mode}_labels.npy", f"{mode}_graph_id.npy"])
        super().__init__("ppi_" + mode, raw_dir=raw_dir, hash_key=(mode, "raw" if self._raw_root else scale),
                         force_reload=force_reload, transform=transform)

    def process(self):
        if self._raw_root is not None:
            self._adopt(raw.read_ppi(self._raw_root, self.mode))
            self._num_targets = 121
            return
        num = {"train": 20, "valid": 2, "test": 2}[self.mode]
        self._gen(num, 2373 * self._scale, 28.8, 50, 121, seed=15, scale=1.0)
        gen = torch.Generator().manual_seed(16)
        for g in self.graphs:
            g.ndata["label"] = (torch.rand(g.num_nodes(), 121, generator=gen) < 0.3).float()

    def __getitem__(self, idx):
        g = self.graphs[idx]
        return g if self._transform is None else self._transform(g)


class BitcoinOTCDataset(_GraphCollection):
    def __init__(self, raw_dir=None, force_reload=False, verbose=False, transform=None):
        self._raw_root = self._find_raw(raw_dir, [os.path.join("bitcoinotc", "bitcoinotc.csv"),      # the reference's extraction
                                                  os.path.join("bitcoinotc", "soc-sign-bitcoinotc.csv"),
                                                  os.path.join("bitcoinotc", "soc-sign-bitcoinotc.csv.gz")])
        super().__init__("bitcoinotc", raw_dir=raw_dir, hash_key=("raw",) if self._raw_root else (),
                         force_reload=force_reload, transform=transform)

    def process(self):
        if self._raw_root is not None:
            self._adopt(raw.read_bitcoinotc(self._raw_root))
            self._num_targets = 1
            return
        self._gen(136, 5881 * 0.05, 2.0, 0, 1, seed=17)
        for g in self.graphs:
            g.edata["h"] = torch.randint(-10, 11, (g.num_edges(), 1)).float()

    def __getitem__(self, idx):
        g = self.graphs[idx]
        return g if self._transform is None else self._transform(g)


class FakeNewsDataset(_GraphCollection):
    def __init__(self, name, feature_name="bert", raw_dir=None, transform=None, scale=1.0):
        self._fn, self._feat, self._scale = name, feature_name, scale
        super().__init__("fakenews_" + name, raw_dir=raw_dir, hash_key=(name, feature_name, scale), transform=transform)

    def process(self):
        num, an = {"politifact": (314, 131.0), "gossipcop": (5464, 57.5)}[self._fn]
        fdim = {"bert": 768, "content": 310, "profile": 10, "spacy": 300}[self._feat]
        self._gen(num, an, 2.0, fdim, 2, seed=18, scale=self._scale)
        n = len(self.graphs)
        perm = torch.randperm(n)
        self.train_mask, self.val_mask, self.test_mask = (torch.zeros(n, dtype=torch.bool) for _ in range(3))
        self.train_mask[perm[: int(0.2 * n)]] = True
        self.val_mask[perm[int(0.2 * n): int(0.3 * n)]] = True
        self.test_mask[perm[int(0.3 * n):]] = True
        self.feature = torch.cat([g.ndata["feat"] for g in self.graphs])


class GDELTDataset(_GraphCollection):
    _spec = (2304 // 24, 300, 6.0)
    _raw_name, _time_step = "GDELT", 15
    is_temporal = True

    def __init__(self, mode="train", raw_dir=None, force_reload=False, verbose=False, transform=None):
        self.mode = mode
        stem = {"valid": "valid", "train": "train", "test": "test"}[mode]
        self._raw_root = self._find_raw(raw_dir, [os.path.join(self._raw_name, stem + ".txt")])
        super().__init__(type(self).__name__.lower() + "_" + mode, raw_dir=raw_dir,
                         hash_key=(mode, "raw") if self._raw_root else (mode,), force_reload=force_reload,
                         transform=transform)

    def process(self):
        if self._raw_root is not None:     # subject / relation / object / time rows on disk (gdelt.py, icews18.py)
            self._adopt(raw.read_temporal_quadruples(self._raw_root, self._time_step))
            self._num_targets = 1
            return
        num, an, ad = self._spec
        self._gen(num, an, ad, 0, 1, seed=19)
        for g in self.graphs:
            g.edata["rel_type"] = torch.randint(0, 240, (g.num_edges(),))

    def __getitem__(self, idx):
        g = self.graphs[idx]
        return g if self._transform is None else self._transform(g)


class ICEWS18Dataset(GDELTDataset):
    _spec = (240, 400, 5.0)
    _raw_name, _time_step = "ICEWS18", 24


class SSTDataset(_GraphCollection):
    """Stanford sentiment treebank as constituency trees (edges child -> parent); 5 sentiment classes per node."""

    def __init__(self, mode="train", glove_embed_file=None, vocab_file=None, raw_dir=None, force_reload=False,
                 verbose=False, transform=None, scale=0.05):
        self.mode, self._scale, self._vocab_file = mode, scale, vocab_file
        self._raw_root = self._find_raw(raw_dir, ["sst"], [mode + ".txt"] + ([] if vocab_file else ["vocab.txt"]))
        super().__init__("sst_" + mode, raw_dir=raw_dir, hash_key=(mode, "raw" if self._raw_root else scale),
                         force_reload=force_reload, transform=transform)

    PAD_WORD, UNK_WORD = raw.SST_PAD, raw.SST_UNK

    def process(self):
        if self._raw_root is not None:     # treebank files on disk (tree.py), parsed without nltk
            self._adopt(raw.read_sst(self._raw_root, self.mode, self._vocab_file))
            self.trees, self.vocab_size, self._num_targets = self.graphs, len(self.vocab), 5
            self.pretrained_emb = None
            return
        num = max(4, int({"train": 8544, "dev": 1101, "test": 2210, "tiny": 5}[self.mode] * self._scale))
        gen = torch.Generator().manual_seed(20)
        self.graphs = []
        self.vocab_size = 19536
        for _ in range(num):
            leaves = int(torch.randint(3, 30, (1,), generator=gen))
            n = 2 * leaves - 1
            # random binary tree: internal node i (0..leaves-2) gets children from the pool
            pool = list(range(leaves - 1, n))
            src, dst = [], []
            for parent in range(leaves - 2, -1, -1):
                for _k in range(2):
                    j = int(torch.randint(0, len(pool), (1,), generator=gen))
                    src.append(pool.pop(j)); dst.append(parent)
                pool.append(parent)
            g = make_graph((torch.tensor(src), torch.tensor(dst)), num_nodes=n)
            g.ndata["x"] = torch.randint(0, self.vocab_size, (n,), generator=gen)
            g.ndata["y"] = torch.randint(0, 5, (n,), generator=gen)
            g.ndata["mask"] = (torch.arange(n) >= leaves - 1).long()
            self.graphs.append(g)
        self.labels = torch.zeros(num, dtype=torch.long)
        self._num_targets = 5

    def __getitem__(self, idx):
        g = self.graphs[idx]
        return g if self._transform is None else self._transform(g)


class SBMMixtureDataset(_GraphCollection):
    """Symmetric stochastic-block-model mixture (exactly the reference's generator semantics, data/sbm.py):
    ``n_graphs`` graphs of ``n_nodes`` nodes in ``n_communities`` equal blocks with intra/inter edge probabilities
    p = a/n, q = b/n drawn so that the average degree is ``avg_deg``."""

    def __init__(self, n_graphs, n_nodes, n_communities, k=2, avg_deg=3, pq="Appendix_C", rng=None, transform=None):
        self._n_graphs, self._n_nodes, self._n_comm, self._avg_deg = n_graphs, n_nodes, n_communities, avg_deg
        self._pq = pq
        super().__init__("sbmmixture", hash_key=(n_graphs, n_nodes, n_communities, avg_deg), transform=transform)

    def process(self):
        gen = torch.Generator().manual_seed(21)
        n, c = self._n_nodes, self._n_comm
        self.graphs, self.line_graphs = [], []
        labels = []
        from ..transforms import line_graph
        for _ in range(self._n_graphs):
            if isinstance(self._pq, (list, tuple)):
                p, q = self._pq
            else:
                a = float(torch.rand(1, generator=gen)) * self._avg_deg * c
                b = (self._avg_deg * c - a) / max(c - 1, 1)
                p, q = min(1.0, a / n), min(1.0, max(b, 0.0) / n)
            comm = torch.arange(n) % c
            same = comm.unsqueeze(0) == comm.unsqueeze(1)
            prob = torch.where(same, torch.tensor(p), torch.tensor(q))
            adj = torch.triu(torch.rand(n, n, generator=gen) < prob, diagonal=1)
            u, v = adj.nonzero(as_tuple=True)
            g = make_graph((torch.cat([u, v]), torch.cat([v, u])), num_nodes=n)
            self.graphs.append(g)
            labels.append(comm)
        self.node_labels = labels
        self.labels = torch.zeros(self._n_graphs, dtype=torch.long)
        self._num_targets = c

    def __getitem__(self, idx):
        return self.graphs[idx], self.node_labels[idx]

    def collate_fn(self, batch):
        from ..batch import batch as batch_graphs
        gs, ls = zip(*batch)
        return batch_graphs(list(gs)), torch.cat(ls)


class AsGraphPredDataset:
    """Adapter adding train/val/test index splits to a graph-classification dataset (reference data/adapter.py)."""

    def __init__(self, dataset, split_ratio=None, **kwargs):
        self.dataset = dataset
        n = len(dataset)
        r = split_ratio or (0.8, 0.1, 0.1)
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(0))
        a, b = int(r[0] * n), int((r[0] + r[1]) * n)
        self.train_idx, self.val_idx, self.test_idx = perm[:a], perm[a:b], perm[b:]
        self.num_classes = getattr(dataset, "num_classes", None)
        self.num_tasks = getattr(dataset, "num_tasks", 1)
        g0 = dataset[0][0] if isinstance(dataset[0], tuple) else dataset[0]
        self.node_feat_size = g0.ndata["feat"].shape[-1] if "feat" in g0.ndata and g0.ndata["feat"].dim() > 1 else None

    def __getitem__(self, idx):
        return self.dataset[idx]

    def __len__(self):
        return len(self.dataset)


class AsLinkPredDataset:
    """Adapter: splits the edges of a single-graph dataset into train/val/test with sampled negatives
    (reference data/adapter.py:AsLinkPredDataset)."""

    def __init__(self, dataset, split_ratio=None, neg_ratio=3, **kwargs):
        g = dataset[0]
        r = split_ratio or (0.8, 0.1, 0.1)
        u, v = g.edges()
        e = g.num_edges()
        gen = torch.Generator().manual_seed(0)
        perm = torch.randperm(e, generator=gen)
        a, b = int(r[0] * e), int((r[0] + r[1]) * e)

        def negs(k):
            return (torch.randint(0, g.num_nodes(), (k * neg_ratio,), generator=gen),
                    torch.randint(0, g.num_nodes(), (k * neg_ratio,), generator=gen))

        self.train_graph = make_graph((u[perm[:a]], v[perm[:a]]), num_nodes=g.num_nodes())
        for k in g.ndata:
            self.train_graph.ndata[k] = g.ndata[k]
        self.val_edges = ((u[perm[a:b]], v[perm[a:b]]), negs(b - a))
        self.test_edges = ((u[perm[b:]], v[perm[b:]]), negs(e - b))
        self.feat_size = g.ndata["feat"].shape[-1] if "feat" in g.ndata else None
        self.neg_ratio = neg_ratio

    def __getitem__(self, idx):
        return self.train_graph

    def __len__(self):
        return 1


# ------------------------------------------------------------------ knowledge graphs
# name: (entities, relations, train, valid, test)
KG_STATS = {"FB15k-237": (14541, 237, 272115, 17535, 20466), "FB15k": (14951, 1345, 483142, 50000, 59071),
            "wn18": (40943, 18, 141442, 5000, 5000)}


class _KGDataset(DGLBuiltinDataset):
    _key = None

    def __init__(self, reverse=True, raw_dir=None, force_reload=False, verbose=False, transform=None, scale=0.1):
        self.reverse, self._scale = reverse, scale
        self._raw_root = raw.locate(raw_dir if raw_dir is not None else get_download_dir(), [self._key],
                                    ["entities.dict", "relations.dict", "train.txt", "valid.txt", "test.txt"])
        super().__init__(self._key, raw_dir=raw_dir, hash_key=(self._key, reverse, "raw" if self._raw_root else scale),
                         force_reload=force_reload, transform=transform)

    @property
    def is_synthetic(self):
        return self._raw_root is None

    def process(self):
        if self._raw_root is not None:     # entities.dict / relations.dict / split triplets on disk (knowledge_graph.py)
            parsed = raw.read_kg_triplets(self._raw_root, self.reverse)
            self._g, self.num_nodes, self.num_rels = parsed.graph, parsed.extra["num_nodes"], parsed.extra["num_rels"]
            self.train, self.valid, self.test = (parsed.extra[k] for k in ("train", "valid", "test"))
            return
        ne, nr, tr, va, te = KG_STATS[self._key]
        ne = max(64, int(ne * self._scale))
        tr, va, te = (max(16, int(x * self._scale)) for x in (tr, va, te))
        gen = torch.Generator().manual_seed(23)
        m = tr + va + te
        h = torch.randint(0, ne, (m,), generator=gen)
        t = torch.randint(0, ne, (m,), generator=gen)
        r = torch.randint(0, nr, (m,), generator=gen)
        masks = torch.zeros(3, m, dtype=torch.bool)
        masks[0, :tr] = True; masks[1, tr:tr + va] = True; masks[2, tr + va:] = True
        if self.reverse:
            h, t, r = torch.cat([h, t]), torch.cat([t, h]), torch.cat([r, r + nr])
            masks = torch.cat([masks, masks], 1)
        g = make_graph((h, t), num_nodes=ne)
        g.edata["etype"] = r
        for i, k in enumerate(("train_mask", "val_mask", "test_mask")):
            g.edata[k] = masks[i]
        g.edata["train_edge_mask"] = masks[0]
        g.ndata["ntype"] = torch.zeros(ne, dtype=torch.long)
        self._g, self.num_nodes, self.num_rels = g, ne, nr * (2 if self.reverse else 1)

    def __getitem__(self, idx):
        assert idx == 0
        return self._g if self._transform is None else self._transform(self._g)

    def __len__(self):
        return 1


FB15k237Dataset = type("FB15k237Dataset", (_KGDataset,), {"_key": "FB15k-237"})
FB15kDataset = type("FB15kDataset", (_KGDataset,), {"_key": "FB15k"})
WN18Dataset = type("WN18Dataset", (_KGDataset,), {"_key": "wn18"})

# ------------------------------------------------------------------ RDF entity classification (heterographs)
# name: (entities, edges, relations, classes, labelled, predict category)
RDF_STATS = {"aifb": (7262, 48810, 78, 4, 176, "Personen"), "mutag": (27163, 148100, 46, 2, 340, "d"),
             "bgs": (94806, 672884, 206, 2, 146, "Lexicon/NamedRockUnit"), "am": (881680, 5668682, 266, 11, 1000, "proxy")}


class _RDFDataset(DGLBuiltinDataset):
    _key = None
    _default_scale = 1.0

    def __init__(self, print_every=10000, insert_reverse=True, raw_dir=None, force_reload=False, verbose=False,
                 transform=None, scale=None):
        self._scale = self._default_scale if scale is None else scale
        self.insert_reverse = insert_reverse
        super().__init__(self._key + "-hetero", raw_dir=raw_dir, hash_key=(self._key, self._scale, insert_reverse),
                         transform=transform)

    def process(self):
        n, e, nr, c, nl, cat = RDF_STATS[self._key]
        n, e = max(64, int(n * self._scale)), max(128, int(e * self._scale))
        nr = max(2, min(nr, 24))  # relation count capped: the synthetic graph keeps the typed structure, not all types
        nt = 6
        gen = torch.Generator().manual_seed(29)
        sizes = torch.full((nt,), n // nt)
        ntypes = [cat] + [f"type{i}" for i in range(1, nt)]
        data = {}
        per = e // nr
        for r in range(nr):
            s, d = r % nt, (r * 7 + 1) % nt
            u = torch.randint(0, int(sizes[s]), (per,), generator=gen)
            v = torch.randint(0, int(sizes[d]), (per,), generator=gen)
            data[(ntypes[s], f"rel{r}", ntypes[d])] = (u, v)
            if self.insert_reverse:
                data[(ntypes[d], f"rev-rel{r}", ntypes[s])] = (v, u)
        g = make_heterograph(data, num_nodes_dict={t: int(sizes[i]) for i, t in enumerate(ntypes)})
        ncat = g.num_nodes(cat)
        nl = min(nl, ncat)
        labels = torch.full((ncat,), -1, dtype=torch.long)
        idx = torch.randperm(ncat, generator=gen)[:nl]
        labels[idx] = torch.randint(0, c, (nl,), generator=gen)
        tm, sm = torch.zeros(ncat, dtype=torch.bool), torch.zeros(ncat, dtype=torch.bool)
        tm[idx[: int(0.8 * nl)]] = True
        sm[idx[int(0.8 * nl):]] = True
        g.nodes[cat].data["label"] = labels
        g.nodes[cat].data["labels"] = labels
        g.nodes[cat].data["train_mask"] = tm
        g.nodes[cat].data["test_mask"] = sm
        self._g, self.num_classes, self.predict_category = g, c, cat

    def __getitem__(self, idx):
        assert idx == 0
        return self._g if self._transform is None else self._transform(self._g)

    def __len__(self):
        return 1


AIFBDataset = type("AIFBDataset", (_RDFDataset,), {"_key": "aifb"})
MUTAGDataset = type("MUTAGDataset", (_RDFDataset,), {"_key": "mutag"})
BGSDataset = type("BGSDataset", (_RDFDataset,), {"_key": "bgs", "_default_scale": 0.2})
AMDataset = type("AMDataset", (_RDFDataset,), {"_key": "am", "_default_scale": 0.02})


class FraudDataset(DGLBuiltinDataset):
    """Multi-relation fraud graphs (yelp: 45,954 reviews / 3 relations / 32 features; amazon: 11,944 users / 3 / 25)."""

    _spec = {"yelp": (45954, 32, ["net_rsr", "net_rtr", "net_rur"], (98630, 1147232, 6805486)),
             "amazon": (11944, 25, ["net_upu", "net_usu", "net_uvu"], (351216, 7132958, 2073474))}

    def __init__(self, name, raw_dir=None, random_seed=717, train_size=0.7, val_size=0.1, force_reload=False,
                 verbose=False, transform=None, scale=0.1):
        self._fname, self.seed, self.train_size, self.val_size, self._scale = name, random_seed, train_size, val_size, scale
        self._raw_root = raw.locate(raw_dir if raw_dir is not None else get_download_dir(), [name, "fraud_" + name],
                                    [raw.FRAUD_LAYOUT[name][0]])
        hk = (name, random_seed, train_size, val_size, "raw") if self._raw_root else (name, random_seed, scale)
        super().__init__("fraud_" + name, raw_dir=raw_dir, hash_key=hk, force_reload=force_reload, transform=transform)

    @property
    def is_synthetic(self):
        return self._raw_root is None

    def process(self):
        if self._raw_root is not None:     # YelpChi.mat / Amazon.mat on disk (fraud.py)
            self._g = self.graph = raw.read_fraud_mat(self._raw_root, self._fname, self.seed, self.train_size,
                                                      self.val_size).graph
            return
        n, f, rels, es = self._spec[self._fname]
        n = max(64, int(n * self._scale))
        gen = torch.Generator().manual_seed(self.seed)
        data = {}
        for r, e in zip(rels, es):
            e = max(64, int(e * self._scale * self._scale))
            u = torch.randint(0, n, (e,), generator=gen)
            v = torch.randint(0, n, (e,), generator=gen)
            data[("review" if self._fname == "yelp" else "user", r, "review" if self._fname == "yelp" else "user")] = (u, v)
        g = make_heterograph(data)
        nt = g.ntypes[0]
        nn_ = g.num_nodes(nt)
        g.nodes[nt].data["feature"] = torch.randn(nn_, f, generator=gen)
        g.nodes[nt].data["label"] = (torch.rand(nn_, generator=gen) < 0.12).long()
        perm = torch.randperm(nn_, generator=gen)
        a, b = int(self.train_size * nn_), int((self.train_size + self.val_size) * nn_)
        for k, idx in (("train_mask", perm[:a]), ("val_mask", perm[a:b]), ("test_mask", perm[b:])):
            m = torch.zeros(nn_, dtype=torch.bool)
            m[idx] = True
            g.nodes[nt].data[k] = m
        self._g = g

    num_classes = 2

    def __getitem__(self, idx):
        assert idx == 0
        return self._g if self._transform is None else self._transform(self._g)

    def __len__(self):
        return 1


class FraudYelpDataset(FraudDataset):
    def __init__(self, **kw):
        super().__init__("yelp", **kw)


class FraudAmazonDataset(FraudDataset):
    def __init__(self, **kw):
        super().__init__("amazon", **kw)
