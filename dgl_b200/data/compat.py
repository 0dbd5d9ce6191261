"""Reference-named bases, loaders and helpers of ``dgl.data`` that sit beside the concrete dataset classes
(parity: /root/reference/python/dgl/data/{__init__,citation_graph,gnn_benchmark,geom_gcn,heterophilous_graphs,
knowledge_graph,rdf,sbm,tensor_serialize,heterograph_serialize,adapter,graph_serialize}.py).

There is no network here, so "named" bases dispatch to the shape-faithful synthetic datasets of ``more.py`` /
``builtin.py``; everything that is pure computation (SBM generator, tensor archives, negative sampling) is real."""
from __future__ import annotations

import os

import numpy as np
import torch

from ..base import DGLError
from . import dgl_binary
from . import more as _more
from .graph_serialize import MAGIC, VERSION, load_graphs, load_labels, save_graphs

__all__ = ["register_data_args", "load_data", "load_cora", "load_citeseer", "load_pubmed", "GNNBenchmarkDataset",
           "GeomGCNDataset", "HeterophilousGraphDataset", "KnowledgeGraphDataset", "RDFGraphDataset", "Entity",
           "Relation", "build_knowledge_graph", "eliminate_self_loops", "sbm", "save_tensors", "load_tensors",
           "save_heterographs", "HeteroGraphData", "GraphData", "StorageMetaData", "tensor_dict_to_ndarray_dict",
           "negative_sample", "is_local_path", "check_local_file_exists", "load_graph_v1", "load_graph_v2",
           "load_labels_v1", "load_labels_v2", "check_pytorch"]


# ---- argparse-era entry points (data/__init__.py:93-112) --------------------------------------------------
def register_data_args(parser):
    parser.add_argument("--dataset", type=str, required=False,
                        help="The input dataset. Can be cora, citeseer, pubmed, syn(synthetic dataset) or reddit")


def load_cora(raw_dir=None, force_reload=False, verbose=True, reverse_edge=True, transform=None):
    from .builtin import CoraGraphDataset
    return CoraGraphDataset(raw_dir=raw_dir, force_reload=force_reload, verbose=verbose, transform=transform)


def load_citeseer(raw_dir=None, force_reload=False, verbose=True, reverse_edge=True, transform=None):
    return _more.CiteseerGraphDataset(raw_dir=raw_dir, force_reload=force_reload, verbose=verbose, transform=transform)


def load_pubmed(raw_dir=None, force_reload=False, verbose=True, reverse_edge=True, transform=None):
    return _more.PubmedGraphDataset(raw_dir=raw_dir, force_reload=force_reload, verbose=verbose, transform=transform)


def load_data(args):
    name = getattr(args, "dataset", None)
    if name == "cora":
        return load_cora()
    if name == "citeseer":
        return load_citeseer()
    if name == "pubmed":
        return load_pubmed()
    if name is not None and name.startswith("reddit"):
        from .builtin import RedditDataset
        return RedditDataset(self_loop="self-loop" in name)
    raise ValueError(f"Unknown dataset: {name}")


# ---- name-dispatching dataset bases ---------------------------------------------------------------------------
def _named_base(doc, table):
    """A class whose constructor takes the dataset ``name`` first and becomes the matching concrete dataset."""
    class _Named:
        __doc__ = doc
        _names = table

        def __new__(cls, name, *args, **kwargs):
            key = str(name).lower().replace("-", "_")
            if key not in cls._names:
                raise DGLError(f"Unknown dataset name {name!r}; expected one of {sorted(cls._names)}")
            return cls._names[key](*args, **kwargs)
    return _Named


GNNBenchmarkDataset = _named_base(
    "Base of the Shchur et al. node-classification sets (gnn_benchmark.py): name in amazon_co_buy_computer, "
    "amazon_co_buy_photo, coauthor_cs, coauthor_physics, cora_full.",
    {"amazon_co_buy_computer": _more.AmazonCoBuyComputerDataset, "amazon_co_buy_photo": _more.AmazonCoBuyPhotoDataset,
     "coauthor_cs": _more.CoauthorCSDataset, "coauthor_physics": _more.CoauthorPhysicsDataset,
     "cora_full": _more.CoraFullDataset})
GeomGCNDataset = _named_base(
    "Base of the Geom-GCN heterophilous sets (geom_gcn.py): chameleon, squirrel, cornell, texas, wisconsin.",
    {"chameleon": _more.ChameleonDataset, "squirrel": _more.SquirrelDataset, "cornell": _more.CornellDataset,
     "texas": _more.TexasDataset, "wisconsin": _more.WisconsinDataset})
HeterophilousGraphDataset = _named_base(
    "Base of Platonov et al.'s heterophilous graphs (heterophilous_graphs.py): roman_empire, amazon_ratings, "
    "minesweeper, tolokers, questions.",
    {"roman_empire": _more.RomanEmpireDataset, "amazon_ratings": _more.AmazonRatingsDataset,
     "minesweeper": _more.MinesweeperDataset, "tolokers": _more.TolokersDataset, "questions": _more.QuestionsDataset})
KnowledgeGraphDataset = _named_base(
    "Base of the link-prediction knowledge graphs (knowledge_graph.py): FB15k-237, FB15k, wn18.",
    {"fb15k_237": _more.FB15k237Dataset, "fb15k": _more.FB15kDataset, "wn18": _more.WN18Dataset})


class RDFGraphDataset(_more._RDFDataset):
    """Base of the RDF entity-classification heterographs (rdf.py): construct with the dataset ``name`` (aifb, mutag,
    bgs, am) or subclass and set ``_key``."""

    def __init__(self, name=None, url=None, predict_category=None, print_every=10000, insert_reverse=True, **kwargs):
        if name is not None:
            key = str(name).lower().replace("-hetero", "")
            if key not in _more.RDF_STATS:
                raise DGLError(f"Unknown RDF dataset {name!r}; expected one of {sorted(_more.RDF_STATS)}")
            self._key = key
        super().__init__(**kwargs)


class Entity:
    """A typed RDF entity (rdf.py:Entity): ``e_id`` is its local name, ``cls`` its node type."""

    def __init__(self, e_id, cls):
        self.id, self.cls = e_id, cls

    def __str__(self):
        return f"{self.cls}/{self.id}"


class Relation:
    """A typed RDF predicate (rdf.py:Relation); ``cls`` becomes the edge type."""

    def __init__(self, cls):
        self.cls = cls

    def __str__(self):
        return str(self.cls)


def build_knowledge_graph(num_nodes, num_rels, train, valid, test, reverse=True):
    """One homogeneous multigraph over all triples ``(head, rel, tail)`` of the three splits with ``etype``,
    ``train_edge_mask`` / ``valid_edge_mask`` / ``test_edge_mask`` edge data and, with ``reverse``, the inverse
    triples under relation ids shifted by ``num_rels`` (knowledge_graph.py:build_knowledge_graph).
    Returns (graph, (train, valid, test) index data)."""
    from ..convert import graph as make_graph
    parts = [np.asarray(x).reshape(-1, 3) for x in (train, valid, test)]
    trip = np.concatenate(parts, 0)
    h, r, t = (torch.as_tensor(trip[:, i]).long() for i in range(3))
    sizes = [len(p) for p in parts]
    split = torch.repeat_interleave(torch.arange(3), torch.tensor(sizes))
    if reverse:
        h, t, r, split = torch.cat([h, t]), torch.cat([t, h]), torch.cat([r, r + num_rels]), torch.cat([split, split])
    g = make_graph((h, t), num_nodes=int(num_nodes))
    g.edata["etype"] = r
    for i, k in enumerate(("train_edge_mask", "valid_edge_mask", "test_edge_mask")):
        g.edata[k] = split == i
    g.ndata["ntype"] = torch.zeros(int(num_nodes), dtype=torch.long)
    return g, tuple(parts)


def eliminate_self_loops(A):
    """Scipy adjacency without its diagonal (gnn_benchmark.py:eliminate_self_loops)."""
    A = A.tolil()
    A.setdiag(0)
    A = A.tocsr()
    A.eliminate_zeros()
    return A


# ---- generators ----------------------------------------------------------------------------------------------------
def sbm(n_blocks, block_size, p, q, rng=None):
    """Symmetric stochastic block model as a scipy matrix: ``n_blocks`` communities of ``block_size`` nodes, expected
    intra-/inter-community *degrees* ``p`` and ``q`` (i.e. edge probabilities p/n and q/n) — sbm.py:16."""
    import scipy.sparse as sp
    n = n_blocks * block_size
    rng = np.random.RandomState() if rng is None else rng
    rows, cols = [], []
    for i in range(n_blocks):
        for j in range(i, n_blocks):
            density = (p if i == j else q) / n
            m = rng.binomial(block_size * block_size, min(max(density, 0.0), 1.0))
            flat = rng.choice(block_size * block_size, size=m, replace=False) if m else np.empty(0, dtype=np.int64)
            rows.append(flat // block_size + i * block_size)
            cols.append(flat % block_size + j * block_size)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    a = sp.coo_matrix((np.ones(rows.shape[0]), (rows, cols)), shape=(n, n))
    upper = sp.triu(a)
    return (upper + sp.triu(a, 1).transpose()).tocoo()


def negative_sample(g, num_samples):
    """Up to ``num_samples`` distinct non-edges without self loops as a ``[2, k]`` numpy array (adapter.py:207)."""
    n = g.num_nodes()
    dens = g.num_edges() / max(n * n, 1)
    want = int(num_samples * (1.0 / max(1e-6, 1 - dens - 1.0 / max(n, 1))) * 1.1) + 8
    cand = np.unique(np.random.randint(0, n, size=(2, want)), axis=1)
    cand = cand[:, cand[0] != cand[1]]
    has = g.has_edges_between(torch.as_tensor(cand[0]), torch.as_tensor(cand[1])).cpu().numpy().astype(bool)
    cand = cand[:, ~has]
    return cand[:, np.random.permutation(cand.shape[1])[:num_samples]]


# ---- tensor / heterograph archives --------------------------------------------------------------------------------
def save_tensors(filename, tensor_dict):
    """Dict[str, Tensor] -> one archive file (tensor_serialize.py:13)."""
    for k, v in tensor_dict.items():
        if not isinstance(k, str):
            raise DGLError("Dict key has to be str")
        if not torch.is_tensor(v):
            raise DGLError("Dict value has to be backend tensor or dgl ndarray")
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    # the reference's tensor container (read back by either framework, memory mapped here)
    dgl_binary.write_tensors(filename, tensor_dict)
    return True


def load_tensors(filename, return_dgl_ndarray=False):
    if dgl_binary.is_dgl_binary(filename):
        return dgl_binary.read_tensors(filename)
    d = torch.load(filename, weights_only=True)
    if d.get("magic") != MAGIC + "-TENSORS":
        raise DGLError("Invalid tensor archive (bad magic)")
    return d["tensors"]


def tensor_dict_to_ndarray_dict(tensor_dict):
    """The reference converts to its NDArray type for the C++ serializer; tensors are the storage type here."""
    return {k: v.detach().cpu() for k, v in tensor_dict.items()}


def save_heterographs(filename, g_list, labels, formats=None):
    """heterograph_serialize.py:save_heterographs — same container as :func:`save_graphs`."""
    save_graphs(filename, g_list, labels or None, formats)


class HeteroGraphData:
    """Serializable snapshot of one heterograph (heterograph_serialize.py:HeteroGraphData)."""

    def __init__(self, packed):
        self._packed = packed

    @staticmethod
    def create(g, formats=None):
        from .graph_serialize import _pack
        return HeteroGraphData(_pack(g))

    def get_graph(self):
        from .graph_serialize import _unpack
        return _unpack(self._packed)


class GraphData(HeteroGraphData):
    """Legacy (homogeneous, v1 file format) spelling of :class:`HeteroGraphData` (graph_serialize.py:GraphData)."""

    def get_graph(self):
        return super().get_graph()


class StorageMetaData:
    """File-level metadata of a graph archive (graph_serialize.py:StorageMetaData)."""

    def __init__(self, version, num_graphs, labels):
        self.version, self.num_graphs, self.labels = version, num_graphs, labels

    @staticmethod
    def load(filename):
        if dgl_binary.is_dgl_binary(filename):
            c = dgl_binary._Cursor(filename)
            num_graph, _ = c.header()
            return StorageMetaData(2, num_graph, dict(c.named_tensors()))
        d = torch.load(filename, weights_only=True)
        return StorageMetaData(d.get("version"), len(d.get("graphs", [])), d.get("labels", {}))


def is_local_path(filepath):
    return not any(str(filepath).startswith(s) for s in ("hdfs://", "viewfs://", "s3://"))


def check_local_file_exists(filename):
    if is_local_path(filename) and not os.path.exists(filename):
        raise DGLError(f"File {filename} does not exist.")


def load_graph_v2(filename, idx_list=None):
    """Current container (heterographs)."""
    return load_graphs(filename, idx_list)


def load_graph_v1(filename, idx_list=None):
    """The reference's v1 files held homogeneous graphs only; our container stores both kinds the same way."""
    return load_graphs(filename, idx_list)


def load_labels_v1(filename):
    return load_labels(filename)


def load_labels_v2(filename):
    return load_labels(filename)


def check_pytorch():
    """Datasets that need torch ops in the reference (MovieLens) guard on the backend here."""
    return True


# ---- superpixel graph construction (superpixel.py:15-75) ----------------------------------------------------------
def sigma(dists, kth=8):
    """Per-row kernel width: mean of the ``kth`` smallest non-self distances (1 for graphs smaller than ``kth``)."""
    n = dists.shape[0]
    if kth > n:
        s = np.ones((n, 1))
    else:
        s = np.partition(dists, kth, axis=-1)[:, :kth + 1].sum(axis=1, keepdims=True) / kth
    return s + 1e-8


def compute_adjacency_matrix_images(coord, feat, use_feat=True):
    """Dense Gaussian affinity between superpixels from their coordinates (and mean colours): symmetric, zero
    diagonal."""
    from scipy.spatial.distance import cdist
    coord = np.asarray(coord).reshape(-1, 2)
    c = cdist(coord, coord)
    expo = -((c / sigma(c)) ** 2)
    if use_feat:
        f = cdist(feat, feat)
        expo = expo - (f / sigma(f)) ** 2
    A = np.exp(expo)
    A = 0.5 * (A + A.T)
    np.fill_diagonal(A, 0)
    return A


def compute_edges_list(A, kth=9):
    """(neighbour ids [N, kth-1], affinities) of the ``kth - 1`` strongest non-self affinities per row; graphs with at
    most ``kth`` nodes become fully connected."""
    n = A.shape[0]
    if n > kth:
        cut = n - (kth - 1)
        idx = np.argpartition(A, cut - 1, axis=-1)[:, cut:]
        return idx, np.take_along_axis(A, idx, axis=-1)
    ids = np.tile(np.arange(n), n).reshape(n, n)
    if n == 1:
        return ids, A
    off = ids != np.arange(n)[:, None]
    return ids[off].reshape(n, -1), A[off].reshape(n, -1)


class SuperPixelDataset:
    """Name-dispatching base of the superpixel image-graph datasets (MNIST / CIFAR10 superpixels)."""

    def __new__(cls, raw_dir=None, name="MNIST", split="train", use_feature=False, **kwargs):
        from . import extra
        table = {"MNIST": extra.MNISTSuperPixelDataset, "CIFAR10": extra.CIFAR10SuperPixelDataset}
        if name not in table:
            raise DGLError(f"Unknown superpixel dataset {name!r}; expected MNIST or CIFAR10")
        return table[name](raw_dir=raw_dir, split=split, use_feature=use_feature, **kwargs)


__all__ += ["sigma", "compute_adjacency_matrix_images", "compute_edges_list", "SuperPixelDataset"]
