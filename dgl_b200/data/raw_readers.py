"""Readers for the raw files the reference's builtin datasets download (python/dgl/data/*.py ``process`` methods).

The dataset classes of this package synthesise shaped stand-ins when nothing is on disk (there is no network here), but
a user switching from the reference usually already has its download directory (``~/.dgl/<name>/...``). Every reader
below turns one of those raw layouts into graphs with the node / edge fields, label conventions and split masks the
reference produces, so ``CoraGraphDataset(raw_dir="~/.dgl")`` yields the real Cora when ``~/.dgl/cora_v2`` exists.

Design differences from the reference (deliberate):

* no networkx round trips - adjacency dictionaries / JSON node-link graphs / bracketed trees are turned into edge
  arrays directly, and edges are emitted in a canonical (src, dst)-sorted order where the reference's order is an
  artefact of dictionary iteration;
* tabular text is parsed with numpy in one pass (``_int_table`` / ``_float_table``), not line by line;
* every reader is a pure function ``path -> RawGraph / RawCollection`` so the dataset classes, the CSV importer and
  the command line tool share them; nothing here touches the dataset cache.

Raw layouts covered (reference file in parentheses): Planetoid ``ind.<name>.*`` pickles (citation_graph.py), Reddit
npz pair (reddit.py), attributed-graph npz (gnn_benchmark.py), GraphSAINT directory (flickr.py / yelp.py), Wiki-CS JSON
(wikics.py), heterophilous npz (heterophilous_graphs.py), Geom-GCN text + split npz (geom_gcn.py, actor.py), fraud
``.mat`` (fraud.py), knowledge-graph triplet files (knowledge_graph.py), TU ``DS_*.txt`` (tu.py), GIN text (gindt.py),
PPI JSON + npy (ppi.py), PTB-bracketed sentiment trees (tree.py), QM7b ``.mat`` / QM9 npz (qm7b.py, qm9.py), Bitcoin-OTC
csv (bitcoinotc.py), temporal quadruple text (gdelt.py / icews18.py), MovieLens tables (movielens.py), N-Triples RDF
(rdf.py).
"""
from __future__ import annotations

import gzip
import json
import os
import pickle
import re
from dataclasses import dataclass, field

import numpy as np
import torch

from ..base import DGLError
from ..convert import graph as make_graph
from ..convert import heterograph as make_heterograph


# ------------------------------------------------------------------------------------------------------ containers ----
@dataclass
class RawGraph:
    """One parsed graph plus what the dataset object reports about it."""
    graph: object
    num_classes: int | None = None
    extra: dict = field(default_factory=dict)


@dataclass
class RawCollection:
    """A parsed multi-graph dataset: graphs, a label tensor (or None) and dataset-level attributes."""
    graphs: list
    labels: object = None
    extra: dict = field(default_factory=dict)


# --------------------------------------------------------------------------------------------------------- helpers ----
_HASHED = re.compile(r".+_[0-9a-f]{8}")


def locate(raw_dir, names, required=()):
    """First ``<raw_dir>/<name>`` (directory or file) that exists and holds every file of ``required``.
    ``names`` are tried in order; returns None when nothing matches."""
    if raw_dir is None:
        return None
    base = os.path.expanduser(raw_dir)
    for name in names:
        head, _, tail = name.partition(os.sep)
        # the reference extracts a download into ``<name>_<sha1(url)[:8]>`` (dgl_dataset.py:_get_hash_url_suffix)
        heads = [head]
        if os.path.isdir(base):
            heads += sorted(d for d in os.listdir(base) if _HASHED.fullmatch(d) and d[:-9] == head)
        for h in heads:
            path = os.path.join(base, h, tail) if tail else os.path.join(base, h)
            if os.path.isdir(path) and all(os.path.exists(os.path.join(path, r)) for r in required):
                return path
            if os.path.isfile(path) and not required:
                return path
    return None


def _bool(x):
    return torch.as_tensor(np.asarray(x).astype(bool))


def _index_mask(idx, n):
    m = torch.zeros(n, dtype=torch.bool)
    m[torch.as_tensor(np.asarray(idx, dtype=np.int64))] = True
    return m


def _int_table(path, delimiter=None):
    """Whole-file integer table. Accepts ',' / tab / blank separated text; returns a 2-D int64 array."""
    with open(path, "rb") as f:
        text = f.read().decode("utf-8", "replace")
    if delimiter is None:
        delimiter = "," if "," in text[:4096] else None
    rows = [ln for ln in text.splitlines() if ln.strip()]
    if not rows:
        return np.zeros((0, 0), dtype=np.int64)
    ncol = len(rows[0].split(delimiter))
    flat = np.array(" ".join(r.replace(",", " ") if delimiter == "," else r for r in rows).split(), dtype=np.float64)
    return flat.reshape(len(rows), ncol).astype(np.int64)


def _float_table(path, delimiter=None):
    with open(path, "rb") as f:
        text = f.read().decode("utf-8", "replace")
    if delimiter is None:
        delimiter = "," if "," in text[:4096] else None
    rows = [ln for ln in text.splitlines() if ln.strip()]
    ncol = len(rows[0].split(delimiter))
    flat = np.array(" ".join(r.replace(",", " ") if delimiter == "," else r for r in rows).split(), dtype=np.float64)
    return flat.reshape(len(rows), ncol)


def _sorted_pairs(u, v, dedup=False):
    """Edges in (src, dst) lexicographic order; optionally without duplicates."""
    u = np.asarray(u, dtype=np.int64)
    v = np.asarray(v, dtype=np.int64)
    if len(u) == 0:
        return u, v
    n = int(max(u.max(), v.max())) + 1
    key = u * n + v
    key = np.unique(key) if dedup else np.sort(key, kind="stable")
    return key // n, key % n


def _symmetric(u, v, num_nodes):
    """Both directions of every pair, simple graph (what ``to_bidirected`` gives), sorted."""
    u = np.asarray(u, dtype=np.int64)
    v = np.asarray(v, dtype=np.int64)
    key = np.unique(np.concatenate([u * num_nodes + v, v * num_nodes + u]))
    return key // num_nodes, key % num_nodes


def _locality_reorder(g):
    """RCM node order + destination-sorted edges (the reference post-processes its citation / co-purchase / Reddit
    graphs with ``reorder_graph(rcmk, dst)``)."""
    from ..transforms import reorder_graph
    return reorder_graph(g, node_permute_algo="rcmk", edge_permute_algo="dst", store_ids=False)


def _scipy():
    import scipy.sparse as sp
    return sp


class _CompatUnpickler(pickle.Unpickler):
    """Planetoid pickles were written by Python 2 / old scipy; map the moved module names."""
    _moved = {"scipy.sparse.csr": "scipy.sparse", "scipy.sparse.lil": "scipy.sparse", "scipy.sparse.coo": "scipy.sparse",
              "scipy.sparse.csc": "scipy.sparse"}

    def find_class(self, module, name):
        return super().find_class(self._moved.get(module, module), name)


def _unpickle(path):
    with open(path, "rb") as f:
        return _CompatUnpickler(f, encoding="latin1").load()


# ------------------------------------------------------------------------------------------------ node-level sets ----
def read_planetoid(root, name, reverse_edge=True, reorder=False):
    """Planetoid split of Cora / Citeseer / Pubmed (citation_graph.py:CitationGraphDataset.process).

    ``allx`` + ``tx`` rows form the feature matrix with the test rows moved to the positions listed in
    ``test.index``; Citeseer's test-id range has holes (isolated nodes) that become zero rows. Features are row
    normalised, labels are the arg-max of the one-hot rows, train = first ``len(y)`` nodes, val = next 500."""
    sp = _scipy()
    part = {k: _unpickle(os.path.join(root, f"ind.{name}.{k}")) for k in ("x", "y", "tx", "ty", "allx", "ally", "graph")}
    with open(os.path.join(root, f"ind.{name}.test.index")) as f:
        test_order = np.array([int(t) for t in f.read().split()], dtype=np.int64)
    test_sorted = np.sort(test_order)
    tx, ty = sp.csr_matrix(part["tx"]), np.asarray(part["ty"])
    lo, span = int(test_sorted[0]), int(test_sorted[-1] - test_sorted[0] + 1)
    if span != tx.shape[0]:                                   # holes in the id range: scatter into a full block
        sel = sp.csr_matrix((np.ones(len(test_sorted)), (test_sorted - lo, np.arange(len(test_sorted)))),
                            shape=(span, tx.shape[0]))
        tx = sel @ tx
        full = np.zeros((span, ty.shape[1]), dtype=ty.dtype)
        full[test_sorted - lo] = ty
        ty = full
    feat = sp.vstack([sp.csr_matrix(part["allx"]), tx]).tocsr()
    onehot = np.vstack([np.asarray(part["ally"]), ty])
    n = feat.shape[0]
    perm = np.arange(n)
    perm[test_order] = test_sorted                            # row i of the result = row perm[i] of the stack
    feat, onehot = feat[perm], onehot[perm]
    rowsum = np.asarray(feat.sum(1)).ravel()
    scale = np.divide(1.0, rowsum, out=np.zeros_like(rowsum, dtype=np.float64), where=rowsum != 0)
    feat = sp.diags(scale) @ feat

    adj = part["graph"]
    u = np.fromiter((s for s, nb in adj.items() for _ in nb), dtype=np.int64)
    v = np.fromiter((d for nb in adj.values() for d in nb), dtype=np.int64)
    if reverse_edge:
        u, v = _symmetric(u, v, n)                            # simple graph, both directions (self loops kept once)
    else:
        a, b = np.minimum(u, v), np.maximum(u, v)             # one arbitrary-but-fixed direction per undirected pair
        u, v = _sorted_pairs(a, b, dedup=True)
    g = make_graph((torch.as_tensor(u), torch.as_tensor(v)), num_nodes=n)
    n_train = len(part["y"])
    g.ndata["train_mask"] = _index_mask(np.arange(n_train), n)
    g.ndata["val_mask"] = _index_mask(np.arange(n_train, n_train + 500), n)
    g.ndata["test_mask"] = _index_mask(test_sorted, n)
    g.ndata["label"] = torch.as_tensor(onehot.argmax(1).astype(np.int64))
    g.ndata["feat"] = torch.as_tensor(np.asarray(feat.todense(), dtype=np.float32))
    if reorder:
        g = _locality_reorder(g)
    return RawGraph(g, int(onehot.shape[1]))


def read_reddit(root, self_loop=False, reorder=True):
    """GraphSAGE Reddit posts (reddit.py): ``reddit[_self_loop]_graph.npz`` scipy COO + ``reddit_data.npz`` with
    ``feature`` / ``label`` / ``node_types`` (1 train, 2 val, 3 test)."""
    sp = _scipy()
    tag = "_self_loop" if self_loop else ""
    coo = sp.load_npz(os.path.join(root, f"reddit{tag}_graph.npz")).tocoo()
    data = np.load(os.path.join(root, "reddit_data.npz"))
    n = coo.shape[0]
    g = make_graph((torch.as_tensor(coo.row.astype(np.int64)), torch.as_tensor(coo.col.astype(np.int64))), num_nodes=n)
    kinds = data["node_types"]
    for key, code in (("train_mask", 1), ("val_mask", 2), ("test_mask", 3)):
        g.ndata[key] = _bool(kinds == code)
    g.ndata["feat"] = torch.as_tensor(data["feature"].astype(np.float32))
    g.ndata["label"] = torch.as_tensor(data["label"].astype(np.int64))
    return RawGraph(_locality_reorder(g) if reorder else g, 41)


def _csr_field(z, prefix):
    sp = _scipy()
    return sp.csr_matrix((z[prefix + "_data"], z[prefix + "_indices"], z[prefix + "_indptr"]), shape=z[prefix + "_shape"])


def read_attributed_npz(path, reorder=True):
    """Shchur et al. benchmark archives (gnn_benchmark.py:_load_npz): CSR adjacency ``adj_*``, attributes either CSR
    ``attr_*`` or dense ``attr_matrix``, labels either CSR ``labels_*`` or dense ``labels``. Symmetrised, simple."""
    with np.load(path, allow_pickle=True) as z:
        z = dict(z)
    adj = _csr_field(z, "adj").tocoo()
    n = int(z["adj_shape"][0])
    feat = (np.asarray(_csr_field(z, "attr").todense()) if "attr_data" in z else z.get("attr_matrix"))
    labels = (np.asarray(_csr_field(z, "labels").todense()) if "labels_data" in z else z.get("labels"))
    u, v = _symmetric(adj.row, adj.col, n)
    g = make_graph((torch.as_tensor(u), torch.as_tensor(v)), num_nodes=n)
    if feat is not None:
        g.ndata["feat"] = torch.as_tensor(np.asarray(feat, dtype=np.float32))
    ncls = None
    if labels is not None:
        g.ndata["label"] = torch.as_tensor(np.asarray(labels).astype(np.int64))
        ncls = int(np.asarray(labels).max()) + 1 if np.asarray(labels).ndim == 1 else int(np.asarray(labels).shape[1])
    return RawGraph(_locality_reorder(g) if reorder else g, ncls)


def read_graphsaint_dir(root, reorder=False):
    """GraphSAINT layout (flickr.py / yelp.py): ``adj_full.npz``, ``feats.npy``, ``class_map.json`` (id -> class, or
    id -> multi-hot list), ``role.json`` with ``tr`` / ``va`` / ``te`` id lists."""
    sp = _scipy()
    coo = sp.load_npz(os.path.join(root, "adj_full.npz")).tocoo()
    feat = np.load(os.path.join(root, "feats.npy"))
    n = feat.shape[0]
    with open(os.path.join(root, "class_map.json")) as f:
        cmap = json.load(f)
    first = next(iter(cmap.values()))
    if isinstance(first, (list, tuple)):
        lab = np.zeros((n, len(first)), dtype=np.int64)
    else:
        lab = np.full((n,), -1, dtype=np.int64)
    ids = np.fromiter((int(k) for k in cmap), dtype=np.int64, count=len(cmap))
    lab[ids] = np.asarray(list(cmap.values()), dtype=np.int64)
    with open(os.path.join(root, "role.json")) as f:
        role = json.load(f)
    g = make_graph((torch.as_tensor(coo.row.astype(np.int64)), torch.as_tensor(coo.col.astype(np.int64))), num_nodes=n)
    g.ndata["feat"] = torch.as_tensor(feat.astype(np.float32))
    g.ndata["label"] = torch.as_tensor(lab)
    for key, r in (("train_mask", "tr"), ("val_mask", "va"), ("test_mask", "te")):
        g.ndata[key] = _index_mask(role[r], n)
    ncls = int(lab.shape[1]) if lab.ndim == 2 else int(lab.max()) + 1
    return RawGraph(_locality_reorder(g) if reorder else g, ncls)


def read_wikics(root, reorder=True):
    """Wiki-CS ``data.json`` (wikics.py): ``features``, ``labels``, ``links`` (out-neighbour lists), 20 train / val /
    stopping splits stored split-major (transposed to node-major) and one test mask."""
    with open(os.path.join(root, "data.json")) as f:
        data = json.load(f)
    n = len(data["labels"])
    deg = np.fromiter((len(js) for js in data["links"]), dtype=np.int64, count=n)
    u = np.repeat(np.arange(n), deg)
    v = np.fromiter((j for js in data["links"] for j in js), dtype=np.int64, count=int(deg.sum()))
    u, v = _symmetric(u, v, n)
    g = make_graph((torch.as_tensor(u), torch.as_tensor(v)), num_nodes=n)
    g.ndata["feat"] = torch.as_tensor(np.asarray(data["features"], dtype=np.float32))
    g.ndata["label"] = torch.as_tensor(np.asarray(data["labels"], dtype=np.int64))
    for key, src in (("train_mask", "train_masks"), ("val_mask", "val_masks"), ("stopping_mask", "stopping_masks")):
        g.ndata[key] = _bool(np.asarray(data[src], dtype=bool).T)
    g.ndata["test_mask"] = _bool(data["test_mask"])
    return RawGraph(_locality_reorder(g) if reorder else g, int(max(data["labels"])) + 1)


def read_heterophilous_npz(path):
    """Platonov et al. archives (heterophilous_graphs.py): ``edges`` [E, 2], ``node_features``, ``node_labels`` and
    split-major ``train_masks`` / ``val_masks`` / ``test_masks`` (10 splits), made bidirected."""
    z = np.load(path)
    n = len(z["node_labels"])
    u, v = _symmetric(z["edges"][:, 0], z["edges"][:, 1], n)
    g = make_graph((torch.as_tensor(u), torch.as_tensor(v)), num_nodes=n)
    g.ndata["feat"] = torch.as_tensor(z["node_features"])
    g.ndata["label"] = torch.as_tensor(z["node_labels"])
    for key in ("train", "val", "test"):
        g.ndata[key + "_mask"] = torch.as_tensor(z[key + "_masks"].T.copy())
    return RawGraph(g, int(len(np.unique(z["node_labels"]))))


def read_geom_gcn(root, name, bag_of_words=False):
    """Geom-GCN text layout (geom_gcn.py / actor.py): ``out1_node_feature_label.txt`` = header + ``id \\t f1,f2,.. \\t
    label`` (for ``bag_of_words`` the middle column lists the active feature ids instead of values),
    ``out1_graph_edges.txt`` = header + ``a \\t b`` stored as (dst, src), and ten ``<name>_split_0.6_0.2_<i>.npz``."""
    with open(os.path.join(root, "out1_node_feature_label.txt")) as f:
        rows = [r.split("\t") for r in f.read().split("\n")[1:] if r.strip()]
    n = len(rows)
    ids = np.array([int(r[0]) for r in rows], dtype=np.int64)
    labels = np.zeros(n, dtype=np.int64)
    labels[ids] = [int(r[2]) for r in rows]
    if bag_of_words:
        cols = [np.array(r[1].split(","), dtype=np.int64) for r in rows]
        feat = np.zeros((n, int(max(c.max() for c in cols)) + 1), dtype=np.float32)
        feat[np.repeat(ids, [len(c) for c in cols]), np.concatenate(cols)] = 1.0
    else:
        feat = np.zeros((n, rows[0][1].count(",") + 1), dtype=np.float32)
        feat[ids] = np.array([r[1].split(",") for r in rows], dtype=np.float32)
    with open(os.path.join(root, "out1_graph_edges.txt")) as f:
        body = f.read().split("\n", 1)[1]
    edges = np.array(body.split(), dtype=np.int64).reshape(-1, 2)
    g = make_graph((torch.as_tensor(edges[:, 1].copy()), torch.as_tensor(edges[:, 0].copy())), num_nodes=n)
    g.ndata["feat"] = torch.as_tensor(feat)
    g.ndata["label"] = torch.as_tensor(labels)
    masks = {"train_mask": [], "val_mask": [], "test_mask": []}
    for i in range(10):
        z = np.load(os.path.join(root, f"{name}_split_0.6_0.2_{i}.npz"))
        for key in masks:
            masks[key].append(z[key].astype(bool))
    for key, cols in masks.items():
        g.ndata[key] = torch.as_tensor(np.stack(cols, axis=1))
    return RawGraph(g, int(labels.max()) + 1)


FRAUD_LAYOUT = {"yelp": ("YelpChi.mat", "review", ("net_rsr", "net_rtr", "net_rur")),
                "amazon": ("Amazon.mat", "user", ("net_upu", "net_usu", "net_uvu"))}


def read_fraud_mat(root, name, seed=717, train_size=0.7, val_size=0.1):
    """CARE-GNN fraud graphs (fraud.py): one ``.mat`` with sparse ``features``, ``label`` and three relation adjacency
    matrices -> a single-node-type heterograph with ``feature`` / ``label`` and a seeded random split (the first 3305
    Amazon users are unlabeled and left out of every split)."""
    from scipy import io
    fname, ntype, rels = FRAUD_LAYOUT[name]
    data = io.loadmat(os.path.join(root, fname), spmatrix=True)
    feat = data["features"]
    feat = np.asarray(feat.todense() if hasattr(feat, "todense") else feat, dtype=np.float32)
    n = feat.shape[0]
    edges = {}
    for rel in rels:
        coo = data[rel].tocoo()
        edges[(ntype, rel, ntype)] = (torch.as_tensor(coo.row.astype(np.int64)), torch.as_tensor(coo.col.astype(np.int64)))
    g = make_heterograph(edges, num_nodes_dict={ntype: n})
    g.ndata["feature"] = torch.as_tensor(feat)
    g.ndata["label"] = torch.as_tensor(np.asarray(data["label"]).squeeze().astype(np.int64))
    pool = np.arange(3305, n) if name == "amazon" else np.arange(n)
    pool = np.random.RandomState(seed).permutation(pool)
    n_tr, n_va = int(train_size * len(pool)), int(val_size * len(pool))
    g.ndata["train_mask"] = _index_mask(pool[:n_tr], n)
    g.ndata["val_mask"] = _index_mask(pool[len(pool) - n_va:], n)
    g.ndata["test_mask"] = _index_mask(pool[n_tr: len(pool) - n_va], n)
    return RawGraph(g, 2)


# ------------------------------------------------------------------------------------------------ knowledge graphs ----
def _read_id_map(path):
    """``id \\t name`` lines -> {name: id}."""
    out = {}
    with open(path, encoding="utf-8") as f:
        for line in f:
            parts = line.rstrip("\n").split("\t")
            if len(parts) >= 2:
                out[parts[1]] = int(parts[0])
    return out


def _read_triples(path, ent, rel):
    rows = []
    with open(path, encoding="utf-8") as f:
        for line in f:
            parts = line.rstrip("\n").split("\t")
            if len(parts) >= 3:
                rows.append((ent[parts[0]], rel[parts[1]], ent[parts[2]]))
    return np.asarray(rows, dtype=np.int64).reshape(-1, 3)


def knowledge_graph_from_triples(num_nodes, num_rels, train, valid, test, reverse=True):
    """Assemble the link-prediction multigraph the reference builds (knowledge_graph.py:build_knowledge_graph).

    Forward edges are grouped by relation in order of first appearance over train -> valid -> test and, inside a
    relation, keep that order; with ``reverse`` the inverse edges (relation id + ``num_rels``) follow in the same
    grouping. ``train_edge_mask`` / ``valid_edge_mask`` / ``test_edge_mask`` mark forward edges only, while
    ``train_mask`` / ``val_mask`` / ``test_mask`` also cover the inverse copies."""
    trip = np.concatenate([np.asarray(x, dtype=np.int64).reshape(-1, 3) for x in (train, valid, test)], 0)
    which = np.repeat(np.array([1, 2, 3]), [len(train), len(valid), len(test)])
    if len(trip) and trip[:, 1].max() >= num_rels:
        raise DGLError("relation id out of range")
    _, first = np.unique(trip[:, 1], return_index=True)
    rank = np.empty(max(num_rels, 1), dtype=np.int64)
    rank[trip[np.sort(first), 1]] = np.arange(len(first))
    order = np.argsort(rank[trip[:, 1]], kind="stable")
    h, r, t, which = trip[order, 0], trip[order, 1], trip[order, 2], which[order]
    fwd_set = which
    if reverse:
        h, t = np.concatenate([h, t]), np.concatenate([t, h])
        r = np.concatenate([r, r + num_rels])
        fwd_set = np.concatenate([which, np.zeros_like(which)])
        which = np.concatenate([which, which])
    g = make_graph((torch.as_tensor(h), torch.as_tensor(t)), num_nodes=int(num_nodes))
    g.edata["etype"] = torch.as_tensor(r)
    g.ndata["ntype"] = torch.zeros(int(num_nodes), dtype=torch.int64)
    for key, code in (("train", 1), ("valid", 2), ("test", 3)):
        g.edata[key + "_edge_mask"] = _bool(fwd_set == code)
        g.edata[("val" if key == "valid" else key) + "_mask"] = _bool(which == code)
    return g


def read_kg_triplets(root, reverse=True):
    """FB15k / FB15k-237 / WN18 layout (knowledge_graph.py): ``entities.dict``, ``relations.dict`` and
    ``train.txt`` / ``valid.txt`` / ``test.txt`` with ``head \\t relation \\t tail`` names."""
    ent = _read_id_map(os.path.join(root, "entities.dict"))
    rel = _read_id_map(os.path.join(root, "relations.dict"))
    train, valid, test = (_read_triples(os.path.join(root, f"{s}.txt"), ent, rel) for s in ("train", "valid", "test"))
    g = knowledge_graph_from_triples(len(ent), len(rel), train, valid, test, reverse)
    return RawGraph(g, None, {"num_nodes": len(ent), "num_rels": len(rel), "train": train, "valid": valid, "test": test})


# ----------------------------------------------------------------------------------------------- graph collections ----
def _split_by_indicator(u, v, indicator, node_data, edge_data):
    """Cut one big graph into per-graph pieces by a node -> graph id vector. Only edges with both ends in the same
    piece survive (TU files never contain others); node / edge ids of each piece keep file order."""
    num_graphs = int(indicator.max()) + 1 if len(indicator) else 0
    order = np.argsort(indicator, kind="stable")
    counts = np.bincount(indicator, minlength=num_graphs)
    start = np.concatenate([[0], np.cumsum(counts)])
    local = np.empty(len(indicator), dtype=np.int64)
    local[order] = np.arange(len(indicator)) - np.repeat(start[:-1], counts)
    gid_e = indicator[u]
    keep = gid_e == indicator[v]
    eorder = np.argsort(np.where(keep, gid_e, num_graphs), kind="stable")[: int(keep.sum())]
    ecount = np.bincount(gid_e[keep], minlength=num_graphs)
    estart = np.concatenate([[0], np.cumsum(ecount)])
    graphs, node_ids = [], []
    for i in range(num_graphs):
        es = eorder[estart[i]: estart[i + 1]]
        ns = order[start[i]: start[i + 1]]
        g = make_graph((torch.as_tensor(local[u[es]]), torch.as_tensor(local[v[es]])), num_nodes=int(counts[i]))
        for k, val in node_data.items():
            g.ndata[k] = torch.as_tensor(val[ns])
        for k, val in edge_data.items():
            g.edata[k] = torch.as_tensor(val[es])
        g.ndata["_ID"] = torch.as_tensor(ns)
        g.edata["_ID"] = torch.as_tensor(es)
        graphs.append(g)
        node_ids.append(ns)
    return graphs, node_ids


def read_tu(root, name, legacy=False):
    """TU Dortmund kernel-benchmark text files (tu.py). ``<name>_A.txt`` (1-based ``a, b`` pairs),
    ``_graph_indicator``, ``_graph_labels`` or ``_graph_attributes`` and optional ``_node_labels`` /
    ``_node_attributes`` / ``_edge_labels`` / ``_edge_attributes``. All label files are shifted to start at zero.

    Field names follow ``TUDataset`` (``node_labels``, ``node_attr``, ``edge_labels``; edge attributes land in
    ``edata['node_labels']`` as in the reference). ``legacy=True`` follows ``LegacyTUDataset``: ``node_label`` plus a
    ``feat`` matrix that is the one-hot node label, the node attributes, or a constant, in that order of preference."""
    def path(kind):
        return os.path.join(root, f"{name}_{kind}.txt")

    def zero_based(a):
        return a - a.min() if a.size else a

    edges = zero_based(_int_table(path("A")))
    indicator = zero_based(_int_table(path("graph_indicator")).ravel())
    extra = {}
    if os.path.exists(path("graph_labels")):
        raw = _int_table(path("graph_labels")).ravel()
        labels = zero_based(raw) if legacy else np.unique(raw, return_inverse=True)[1]
        extra["num_labels"] = int(labels.max()) + 1
        labels = torch.as_tensor(labels if legacy else labels.reshape(-1, 1))
    elif os.path.exists(path("graph_attributes")):
        labels = torch.as_tensor(_float_table(path("graph_attributes")))
        extra["num_labels"] = None
    else:
        raise DGLError(f"{name}: neither graph_labels nor graph_attributes found under {root}")
    node_data, edge_data = {}, {}
    n = len(indicator)
    if legacy:
        if os.path.exists(path("node_labels")):
            nl = zero_based(_int_table(path("node_labels")).ravel())
            node_data["node_label"] = nl
            node_data["feat"] = np.eye(int(nl.max()) + 1, dtype=np.float32)[nl]
            extra["data_mode"] = "node_label"
        if os.path.exists(path("node_attributes")):
            node_data["feat"] = _float_table(path("node_attributes")).astype(np.float32)
            extra["data_mode"] = "node_attr"
        if "feat" not in node_data:
            node_data["feat"] = np.ones((n, 1), dtype=np.float32)
            extra["data_mode"] = "constant"
    else:
        for kind, (where, key) in {"node_labels": ("n", "node_labels"), "node_attributes": ("n", "node_attr"),
                                   "edge_labels": ("e", "edge_labels"), "edge_attributes": ("e", "node_labels")}.items():
            if not os.path.exists(path(kind)):
                continue
            val = zero_based(_int_table(path(kind))) if "label" in kind else _float_table(path(kind))
            (node_data if where == "n" else edge_data)[key] = val
    graphs, _ = _split_by_indicator(edges[:, 0], edges[:, 1], indicator, node_data, edge_data)
    extra["max_num_node"] = int(np.bincount(indicator).max())
    return RawCollection(graphs, labels, extra)


def read_gin_text(path, self_loop=False, degree_as_nlabel=False):
    """"How powerful are GNNs" text format (gindt.py): first line = number of graphs; per graph ``n label`` and then
    one line per node ``tag deg nb_1 .. nb_deg [attributes..]``. Graph and node tags are renumbered in order of first
    appearance; without attributes the node feature is the one-hot tag (or one-hot in-degree)."""
    with open(path) as f:
        tokens = f.read().split("\n")
    num = int(tokens[0].strip())
    pos = 1
    glabel, ntag = {}, {}
    graphs, labels, has_attr = [], [], False
    for _ in range(num):
        n, lab = (int(t) for t in tokens[pos].split())
        pos += 1
        labels.append(glabel.setdefault(lab, len(glabel)))
        src, dst, tags, attrs = [], [], [], []
        for j in range(n):
            row = tokens[pos].split()
            pos += 1
            deg = int(row[1])
            if len(row) < deg + 2:
                raise DGLError(f"{path}: node line announces {deg} neighbours but lists {len(row) - 2}")
            tags.append(ntag.setdefault(int(row[0]), len(ntag)))
            nbrs = [int(t) for t in row[2: 2 + deg]]
            src += [j] * deg
            dst += nbrs
            if self_loop:
                src.append(j)
                dst.append(j)
            if len(row) > deg + 2:
                attrs.append([float(t) for t in row[2 + deg:]])
        g = make_graph((torch.tensor(src, dtype=torch.int64), torch.tensor(dst, dtype=torch.int64)), num_nodes=n)
        g.ndata["label"] = torch.tensor(tags, dtype=torch.int64)
        if attrs:
            g.ndata["attr"] = torch.tensor(attrs, dtype=torch.float32)
            has_attr = True
        graphs.append(g)
    if not has_attr:
        if degree_as_nlabel:
            for g in graphs:
                g.ndata["label"] = g.in_degrees().to(torch.int64)
        present = torch.unique(torch.cat([g.ndata["label"] for g in graphs]))
        contiguous = int(present[-1]) + 1 == len(present) and not degree_as_nlabel
        lut = torch.full((int(present[-1]) + 1,), -1, dtype=torch.int64)
        lut[present] = present if contiguous else torch.arange(len(present))
        for g in graphs:
            g.ndata["attr"] = torch.nn.functional.one_hot(lut[g.ndata["label"]], len(present)).float()
    extra = {"gclasses": len(glabel), "nclasses": len(ntag), "eclasses": 0, "dim_nfeats": int(graphs[0].ndata["attr"].shape[1]),
             "glabel_dict": glabel, "nlabel_dict": ntag, "N": num,
             "n": sum(g.num_nodes() for g in graphs), "m": sum(g.num_edges() for g in graphs)}
    return RawCollection(graphs, torch.tensor(labels, dtype=torch.int64), extra)


PPI_RANGES = {"train": (1, 21), "valid": (21, 23), "test": (23, 25)}


def read_ppi(root, mode):
    """GraphSAGE protein-protein interaction graphs (ppi.py): ``<mode>_graph.json`` node-link JSON,
    ``<mode>_feats.npy`` [N, 50], ``<mode>_labels.npy`` [N, 121], ``<mode>_graph_id.npy``; graph ids 1-20 train,
    21-22 valid, 23-24 test. Every undirected link becomes two directed edges."""
    with open(os.path.join(root, f"{mode}_graph.json")) as f:
        blob = json.load(f)
    feats = np.load(os.path.join(root, f"{mode}_feats.npy")).astype(np.float32)
    labels = np.load(os.path.join(root, f"{mode}_labels.npy")).astype(np.float32)
    gid = np.load(os.path.join(root, f"{mode}_graph_id.npy")).astype(np.int64)
    n = len(gid)
    ids = {node["id"]: i for i, node in enumerate(blob["nodes"])}
    links = blob.get("links", blob.get("edges", []))
    u = np.fromiter((ids.get(e["source"], e["source"]) for e in links), dtype=np.int64, count=len(links))
    v = np.fromiter((ids.get(e["target"], e["target"]) for e in links), dtype=np.int64, count=len(links))
    if not blob.get("directed", False):
        u, v = _symmetric(u, v, n)
    lo, hi = PPI_RANGES[mode]
    inside = (gid >= lo) & (gid < hi)
    remap = np.where(inside, gid - lo, hi - lo)               # nodes of other ids fall into a discarded bucket
    graphs, _ = _split_by_indicator(u, v, remap, {"feat": feats, "label": labels}, {})
    return RawCollection(graphs[: hi - lo], None, {"num_labels": 121})


def parse_bracketed_trees(text):
    """Penn-treebank style ``(label child child ...)`` / ``(label word)`` sentences -> list of nested tuples
    ``(label, [children])`` with string leaves. Replaces nltk's BracketParseCorpusReader for the SST files."""
    trees, stack = [], []
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if c == "(":
            j = i + 1
            while j < n and not text[j].isspace() and text[j] not in "()":
                j += 1
            stack.append((text[i + 1: j], []))
            i = j
        elif c == ")":
            node = stack.pop()
            if stack:
                stack[-1][1].append(node)
            else:
                trees.append(node)
            i += 1
        elif c.isspace():
            i += 1
        else:
            j = i
            while j < n and not text[j].isspace() and text[j] not in "()":
                j += 1
            stack[-1][1].append(text[i: j])
            i = j
    if stack:
        raise DGLError("unbalanced brackets in tree file")
    return trees


SST_PAD, SST_UNK = -1, -1


def read_sst(root, mode, vocab_file=None):
    """Stanford sentiment treebank (tree.py): ``vocab.txt`` (one token per line) + ``<mode>.txt`` bracketed trees.
    Each tree becomes a graph with child -> parent edges, nodes numbered in pre-order from the root; ``x`` is the
    vocabulary id of a leaf word (``PAD_WORD`` = -1 for internal nodes, ``UNK_WORD`` = -1 for unknown words), ``y``
    the 5-way sentiment label and ``mask`` = 1 on leaves."""
    vocab = {}
    with open(vocab_file or os.path.join(root, "vocab.txt"), encoding="utf-8") as f:
        for line in f:
            vocab.setdefault(line.strip(), len(vocab))
    with open(os.path.join(root, f"{mode}.txt"), encoding="utf-8") as f:
        trees = parse_bracketed_trees(f.read())
    graphs = []
    for root_node in trees:
        xs, ys, masks, src, dst = [SST_PAD], [int(root_node[0])], [0], [], []
        # a child is numbered when it is met, then its subtree is numbered completely before the next sibling

        def visit(pid, node):
            for child in node[1]:
                cid = len(xs)
                leaf = len(child[1]) == 1 and isinstance(child[1][0], str)
                xs.append(vocab.get(child[1][0].lower(), SST_UNK) if leaf else SST_PAD)
                ys.append(int(child[0]))
                masks.append(1 if leaf else 0)
                src.append(cid)
                dst.append(pid)
                if not leaf:
                    visit(cid, child)
        visit(0, root_node)
        g = make_graph((torch.tensor(src, dtype=torch.int64), torch.tensor(dst, dtype=torch.int64)), num_nodes=len(xs))
        g.ndata["x"] = torch.tensor(xs, dtype=torch.int64)
        g.ndata["y"] = torch.tensor(ys, dtype=torch.int64)
        g.ndata["mask"] = torch.tensor(masks, dtype=torch.int64)
        graphs.append(g)
    return RawCollection(graphs, None, {"vocab": vocab, "num_classes": 5})


def read_qm7b(path):
    """QM7b (qm7b.py): ``X`` [G, 23, 23] Coulomb matrices, ``T`` [G, 14] targets. Non-zero entries become edges with
    the Coulomb value in ``edata['h']``."""
    from scipy import io
    data = io.loadmat(path)
    graphs = []
    for mat in data["X"]:
        r, c = np.nonzero(mat)
        g = make_graph((torch.as_tensor(r.astype(np.int64)), torch.as_tensor(c.astype(np.int64))))
        g.edata["h"] = torch.as_tensor(mat[r, c].reshape(-1, 1).astype(np.float32))
        graphs.append(g)
    return RawCollection(graphs, torch.as_tensor(data["T"].astype(np.float32)), {"num_tasks": int(data["T"].shape[1])})


class QM9Molecules:
    """Lazy QM9 (qm9.py): atoms of all molecules are stored concatenated (``N`` counts, ``R`` coordinates, ``Z``
    atomic numbers); molecule ``i`` becomes a graph joining atoms closer than ``cutoff`` (both directions, no self
    loops) when it is indexed."""

    def __init__(self, path, label_keys, cutoff=5.0):
        z = np.load(path, allow_pickle=True)
        self.N, self.R, self.Z = z["N"], z["R"], z["Z"]
        self.label = np.stack([z[k] for k in label_keys], axis=1)
        self.offsets = np.concatenate([[0], np.cumsum(self.N)])
        self.cutoff = cutoff

    def __len__(self):
        return len(self.N)

    def __getitem__(self, i):
        lo, hi = self.offsets[i], self.offsets[i + 1]
        pos = self.R[lo:hi]
        near = np.linalg.norm(pos[:, None] - pos[None], axis=-1) <= self.cutoff
        np.fill_diagonal(near, False)
        u, v = np.nonzero(near | near.T)
        g = make_graph((torch.as_tensor(u.astype(np.int64)), torch.as_tensor(v.astype(np.int64))), num_nodes=hi - lo)
        g.ndata["R"] = torch.as_tensor(pos.astype(np.float32))
        g.ndata["Z"] = torch.as_tensor(self.Z[lo:hi].astype(np.int64))
        return g, torch.as_tensor(self.label[i].astype(np.float32))


def _open_text(path):
    return gzip.open(path, "rt") if path.endswith(".gz") else open(path)


class CumulativeSnapshots:
    """Lazy sequence of growing graphs over one event table: snapshot ``i`` holds every event whose time bucket is
    ``<= first_bucket + i``, in file order (what the reference's temporal datasets hand out). Graphs are built on access -
    the sequence costs one table, not one graph per time step. Node counts follow the events included so far
    (``max id + 1``), as in the reference, unless ``num_nodes`` fixes them."""

    def __init__(self, src, dst, bucket, edge_data, first, count, num_nodes=None):
        self.src, self.dst, self.bucket, self.edge_data = src, dst, bucket, edge_data
        self.first, self.count, self.num_nodes = int(first), int(count), num_nodes

    def __len__(self):
        return self.count

    def __iter__(self):
        return (self[i] for i in range(self.count))

    def __getitem__(self, i):
        if i < 0:
            i += self.count
        if not 0 <= i < self.count:
            raise IndexError(i)
        keep = self.bucket <= self.first + i
        g = make_graph((torch.as_tensor(self.src[keep]), torch.as_tensor(self.dst[keep])), num_nodes=self.num_nodes)
        for k, v in self.edge_data.items():
            g.edata[k] = torch.as_tensor(v[keep])
        return g


def read_bitcoinotc(path):
    """Bitcoin-OTC trust network (bitcoinotc.py): csv rows ``source, target, rating, unix time``. Ids are shifted so
    that the smallest one is 0, time is bucketed in 14-day steps (rounded), and snapshot ``i`` accumulates every rating
    up to bucket ``i`` with the rating in ``edata['h']`` ([E, 1], int64); the last bucket closes the range and is not
    emitted, like in the reference."""
    with _open_text(path) as f:
        rows = np.array([ln.split(",") for ln in f.read().split("\n") if ln.strip()], dtype=np.float64).astype(np.int64)
    pairs = rows[:, :2] - rows[:, :2].min()
    bucket = np.around((rows[:, 3] - rows[:, 3].min()) / (14 * 86400.0)).astype(np.int64)
    graphs = CumulativeSnapshots(pairs[:, 0], pairs[:, 1], bucket, {"h": rows[:, 2].reshape(-1, 1)}, 0, int(bucket.max()))
    return RawCollection(graphs, None, {"num_nodes": int(pairs.max()) + 1})


def read_temporal_quadruples(path, time_step=1):
    """Event knowledge graphs (gdelt.py / icews18.py): rows ``subject relation object time [..]``. Time is bucketed by
    ``floor(time / time_step)`` (15 minutes-units for GDELT, 24 hours for ICEWS18); snapshot ``i`` accumulates every
    event up to bucket ``start + i`` with the relation id in ``edata['rel_type']`` ([E, 1], int64)."""
    table = _int_table(path)[:, :4]
    bucket = np.floor(table[:, 3] / time_step).astype(np.int64)
    start, end = int(bucket.min()), int(bucket.max())
    graphs = CumulativeSnapshots(table[:, 0].copy(), table[:, 2].copy(), bucket, {"rel_type": table[:, 1].reshape(-1, 1).copy()},
                                 start, end - start + 1)
    return RawCollection(graphs, None, {"start_time": start, "end_time": end, "time_step": time_step})


def read_movielens_100k(root, test_ratio=0.1, valid_ratio=0.1, seed=0):
    """MovieLens-100K tables (movielens.py): ``u.data`` (user, item, rating, time; tab separated), ``u.user``
    (id|age|gender|occupation|zip), ``u.item`` (id|title|date||url|19 genre flags). Returns a bipartite heterograph
    ``user -rates-> movie`` + reverse with ``rate`` and split masks on the edges and simple numeric side features."""
    ratings = _int_table(os.path.join(root, "u.data"), delimiter="\t")
    users, ui = np.unique(ratings[:, 0], return_inverse=True)
    items, ii = np.unique(ratings[:, 1], return_inverse=True)
    u, v = torch.as_tensor(ui.astype(np.int64)), torch.as_tensor(ii.astype(np.int64))
    g = make_heterograph({("user", "user-movie", "movie"): (u, v), ("movie", "movie-user", "user"): (v, u)},
                         num_nodes_dict={"user": len(users), "movie": len(items)})
    rate = torch.as_tensor(ratings[:, 2].astype(np.float32))
    perm = np.random.RandomState(seed).permutation(len(ratings))
    n_te, n_va = int(test_ratio * len(perm)), int(valid_ratio * len(perm))
    split = {"test_mask": perm[:n_te], "valid_mask": perm[n_te: n_te + n_va], "train_mask": perm[n_te + n_va:]}
    for et in ("user-movie", "movie-user"):
        g.edges[et].data["rate"] = rate
        for key, idx in split.items():
            g.edges[et].data[key] = _index_mask(idx, len(ratings))
    user_file, item_file = os.path.join(root, "u.user"), os.path.join(root, "u.item")
    if os.path.exists(user_file):
        with open(user_file, encoding="latin1") as f:
            rows = [ln.split("|") for ln in f.read().split("\n") if ln.strip()]
        by_id = {int(r[0]): r for r in rows}
        occ = {o: i for i, o in enumerate(sorted({r[3] for r in rows}))}
        feat = np.zeros((len(users), 2 + len(occ)), dtype=np.float32)
        for k, uid in enumerate(users):
            r = by_id.get(int(uid))
            if r is not None:
                feat[k, 0] = float(r[1]) / 100.0
                feat[k, 1] = 1.0 if r[2] == "M" else 0.0
                feat[k, 2 + occ[r[3]]] = 1.0
        g.nodes["user"].data["feat"] = torch.as_tensor(feat)
    if os.path.exists(item_file):
        with open(item_file, encoding="latin1") as f:
            rows = [ln.split("|") for ln in f.read().split("\n") if ln.strip()]
        by_id = {int(r[0]): r for r in rows}
        feat = np.zeros((len(items), 19), dtype=np.float32)
        for k, iid in enumerate(items):
            r = by_id.get(int(iid))
            if r is not None:
                feat[k] = [float(x) for x in r[-19:]]
        g.nodes["movie"].data["feat"] = torch.as_tensor(feat)
    return RawGraph(g, None, {"user_ids": users, "movie_ids": items})


# --------------------------------------------------------------------------------------------------------- RDF ----
def parse_ntriples(path):
    """Line-oriented N-Triples / simple N3 (``<s> <p> <o> .`` with IRIs, blank nodes and literals) -> list of
    (subject, predicate, object) strings; literals keep their quotes so they can be told from IRIs. Enough for the
    ``.nt`` dumps of AIFB / MUTAG / BGS / AM (rdf.py parses the same files with rdflib, which is not installed here)."""
    out = []
    with _open_text(path) as f:
        for line in f:
            line = line.strip()
            if not line or line[0] == "#" or line.startswith("@"):
                continue
            if line.endswith("."):
                line = line[:-1].rstrip()
            terms, i = [], 0
            while i < len(line) and len(terms) < 3:
                c = line[i]
                if c.isspace():
                    i += 1
                elif c == "<":
                    j = line.index(">", i)
                    terms.append(line[i + 1: j])
                    i = j + 1
                elif c == '"':
                    j = i + 1
                    while j < len(line) and (line[j] != '"' or line[j - 1] == "\\"):
                        j += 1
                    k = j + 1
                    while k < len(line) and not line[k].isspace():
                        k += 1
                    terms.append(line[i: k])
                    i = k
                else:
                    j = i
                    while j < len(line) and not line[j].isspace():
                        j += 1
                    terms.append(line[i: j])
                    i = j
            if len(terms) == 3:
                out.append(tuple(terms))
    return out


def rdf_heterograph(triples, type_predicate="http://www.w3.org/1999/02/22-rdf-syntax-ns#type", drop=()):
    """Entity-typed heterograph from RDF triples (the construction of rdf.py:RDFGraphDataset.process_raw_tuples):
    the node type of an IRI is its ``rdf:type`` local name (``_Thing`` when untyped, ``_Literal`` for literals), the
    edge type is the predicate's local name, and every relation gets a ``rev-`` twin. Returns (graph, {ntype: names})."""
    def local(iri):
        return iri.rstrip("/#").rsplit("#", 1)[-1].rsplit("/", 1)[-1]

    kind = {}
    for s, p, o in triples:
        if p == type_predicate:
            kind.setdefault(s, local(o))
    members, index = {}, {}

    def node(term):
        nt = "_Literal" if term.startswith('"') else kind.get(term, "_Thing")
        slot = index.get((nt, term))
        if slot is None:
            slot = index[(nt, term)] = len(members.setdefault(nt, []))
            members[nt].append(term)
        return nt, slot

    edges = {}
    for s, p, o in triples:
        if p == type_predicate or p in drop:
            continue
        (st, si), (ot, oi) = node(s), node(o)
        rel = local(p)
        edges.setdefault((st, rel, ot), ([], []))
        edges[(st, rel, ot)][0].append(si)
        edges[(st, rel, ot)][1].append(oi)
        edges.setdefault((ot, "rev-" + rel, st), ([], []))
        edges[(ot, "rev-" + rel, st)][0].append(oi)
        edges[(ot, "rev-" + rel, st)][1].append(si)
    for term in kind:
        node(term)
    data = {k: (torch.tensor(a, dtype=torch.int64), torch.tensor(b, dtype=torch.int64)) for k, (a, b) in edges.items()}
    g = make_heterograph(data, num_nodes_dict={nt: len(v) for nt, v in members.items()})
    return g, members


# ------------------------------------------------------------------------------------------- source table (nodes) ----
def _npz_source(name):
    return ([name], [name + ".npz"], lambda root, **kw: read_attributed_npz(os.path.join(root, name + ".npz")))


def _hetero_source(name):
    # heterophilous_graphs.py lower-cases the name and turns '-' into '_' for both the directory and the archive
    under = name.replace("-", "_")
    return ([os.path.join(under, under + ".npz"), os.path.join(name, name + ".npz")], [],
            lambda root, **kw: read_heterophilous_npz(root))


def _geom_source(name, bow=False):
    return ([name], ["out1_node_feature_label.txt", "out1_graph_edges.txt"],
            lambda root, **kw: read_geom_gcn(root, name, bag_of_words=bow))


def _planetoid_source(dirname, name):
    return ([dirname], [f"ind.{name}.x", f"ind.{name}.graph"],
            lambda root, reverse_edge=True, reorder=False, **kw: read_planetoid(root, name, reverse_edge, reorder))


# dataset key -> (candidate raw directory names, files that must be there, reader(root, **options))
NODE_SOURCES = {
    "cora": _planetoid_source("cora_v2", "cora_v2"), "citeseer": _planetoid_source("citeseer", "citeseer"),
    "pubmed": _planetoid_source("pubmed", "pubmed"),
    "reddit": (["reddit"], ["reddit_data.npz", "reddit_graph.npz"], lambda root, **kw: read_reddit(root, False)),
    "reddit_self_loop": (["reddit_self_loop"], ["reddit_data.npz", "reddit_self_loop_graph.npz"],
                         lambda root, **kw: read_reddit(root, True)),
    "cora_full": _npz_source("cora_full"), "amazon_co_buy_computer": _npz_source("amazon_co_buy_computer"),
    "amazon_co_buy_photo": _npz_source("amazon_co_buy_photo"), "coauthor_cs": _npz_source("coauthor_cs"),
    "coauthor_physics": _npz_source("coauthor_physics"),
    "wiki_cs": (["wiki_cs"], ["data.json"], lambda root, **kw: read_wikics(root)),
    "flickr": (["flickr"], ["adj_full.npz", "feats.npy", "class_map.json", "role.json"],
               lambda root, reorder=False, **kw: read_graphsaint_dir(root, reorder)),
    "yelp": (["yelp"], ["adj_full.npz", "feats.npy", "class_map.json", "role.json"],
             lambda root, reorder=False, **kw: read_graphsaint_dir(root, reorder)),
    "chameleon": _geom_source("chameleon"), "squirrel": _geom_source("squirrel"), "cornell": _geom_source("cornell"),
    "texas": _geom_source("texas"), "wisconsin": _geom_source("wisconsin"), "actor": _geom_source("actor", bow=True),
    "roman_empire": _hetero_source("roman-empire"), "amazon_ratings": _hetero_source("amazon-ratings"),
    "minesweeper": _hetero_source("minesweeper"), "tolokers": _hetero_source("tolokers"),
    "questions": _hetero_source("questions"),
}


def find_node_source(key, raw_dir):
    """Raw directory of node dataset ``key`` under ``raw_dir`` (None when the files are not there)."""
    spec = NODE_SOURCES.get(key)
    return None if spec is None else locate(raw_dir, spec[0], spec[1])


def load_node_source(key, root, **options):
    return NODE_SOURCES[key][2](root, **options)
