"""Dataset utilities (parity: /root/reference/python/dgl/data/utils.py): split_dataset, Subset, add_nodepred_split,
mask helpers, makedirs. Download helpers raise: there is no network in this environment."""
from __future__ import annotations

import os

import numpy as np
import torch

from ..base import DGLError
from .graph_serialize import load_graphs, load_info, load_labels, save_graphs, save_info  # noqa: F401


def download(url, path=None, overwrite=True, sha1_hash=None, retries=5, verify_ssl=True, log=True):
    raise DGLError("network access is not available in this environment; place the file manually")


def extract_archive(file, target_dir, overwrite=True):
    import tarfile
    import zipfile
    if file.endswith(".zip"):
        with zipfile.ZipFile(file) as z:
            z.extractall(target_dir)
    else:
        with tarfile.open(file) as t:
            t.extractall(target_dir)


def makedirs(path):
    os.makedirs(os.path.expanduser(os.path.normpath(path)), exist_ok=True)


def generate_mask_tensor(mask):
    return torch.as_tensor(np.asarray(mask), dtype=torch.bool)


class Subset:
    def __init__(self, dataset, indices):
        self.dataset, self.indices = dataset, indices

    def __getitem__(self, item):
        return self.dataset[self.indices[item]]

    def __len__(self):
        return len(self.indices)


def split_dataset(dataset, frac_list=None, shuffle=False, random_state=None):
    frac_list = frac_list or [0.8, 0.1, 0.1]
    n = len(dataset)
    idx = np.random.RandomState(random_state).permutation(n) if shuffle else np.arange(n)
    lens = [int(f * n) for f in frac_list]
    lens[-1] = n - sum(lens[:-1])
    out, off = [], 0
    for l in lens:
        out.append(Subset(dataset, idx[off: off + l]))
        off += l
    return out


def add_nodepred_split(dataset, ratio, ntype=None):
    g = dataset[0]
    n = g.num_nodes(ntype)
    perm = torch.randperm(n)
    a, b = int(ratio[0] * n), int((ratio[0] + ratio[1]) * n)
    for name, idx in (("train_mask", perm[:a]), ("val_mask", perm[a:b]), ("test_mask", perm[b:])):
        m = torch.zeros(n, dtype=torch.bool)
        m[idx] = True
        g.nodes[ntype].data[name] = m


# ------------------------------------------------------------------------------------------------------
# Further helpers of the reference's data/utils.py
# ------------------------------------------------------------------------------------------------------
def check_sha1(filename, sha1_hash):
    """True iff the file's SHA-1 hex digest equals ``sha1_hash`` (utils.py:check_sha1)."""
    import hashlib
    h = hashlib.sha1()
    with open(filename, "rb") as f:
        for block in iter(lambda: f.read(1 << 20), b""):
            h.update(block)
    return h.hexdigest() == sha1_hash


def loadtxt(path, delimiter, dtype=None):
    """Delimited text -> 2-D numpy array (pandas' C parser when available)."""
    import numpy as np
    try:
        import pandas as pd
        return pd.read_csv(path, delimiter=delimiter, header=None).values
    except ImportError:
        return np.loadtxt(path, delimiter=delimiter)


def idx2mask(idx, len):  # noqa: A002 - reference argument name
    """0/1 float vector of length ``len`` with ones at ``idx``."""
    import numpy as np
    mask = np.zeros(len)
    mask[idx] = 1
    return mask


def deprecate_property(old, new):
    import warnings
    warnings.warn(f"Property {old} will be deprecated, please use {new} instead.")


def deprecate_function(old, new):
    import warnings
    warnings.warn(f"Function {old} will be deprecated, please use {new} instead.")


def deprecate_class(old, new):
    import warnings
    warnings.warn(f"Class {old} will be deprecated, please use {new} instead.")


def _undirected_simple(g):
    """Symmetrised, loop-free, duplicate-free adjacency of ``g`` as a host ``dgl_b200.sparse`` matrix."""
    import torch

    from .. import sparse as dglsp
    u, v = [t.cpu().long() for t in g.edges()]
    keep = u != v
    u, v = u[keep], v[keep]
    n = g.num_nodes()
    key = torch.unique(torch.cat([u * n + v, v * n + u]))
    return dglsp.from_coo(key // n, key % n, torch.ones(key.shape[0]), (n, n))


def _pagerank(adj, personalization=None, alpha=0.85, iters=100, tol=1e-6):
    """Power iteration on the column-stochastic walk matrix of an undirected adjacency (dangling mass and teleport
    follow ``personalization``; uniform when ``None``) — the quantity networkx's ``pagerank`` converges to."""
    import torch

    from .. import sparse as dglsp
    n = adj.shape[0]
    deg = dglsp.sum(adj, 1)
    tele = torch.full((n,), 1.0 / n) if personalization is None else personalization / personalization.sum()
    x = torch.full((n,), 1.0 / n)
    inv = torch.where(deg > 0, 1.0 / deg.clamp(min=1), torch.zeros_like(deg))
    for _ in range(iters):
        spread = dglsp.spmm(adj, (x * inv).unsqueeze(1)).squeeze(1)   # symmetric: A^T == A
        dangling = x[deg == 0].sum()
        new = alpha * (spread + dangling * tele) + (1 - alpha) * tele
        if float((new - x).abs().sum()) < n * tol:
            x = new
            break
        x = new
    return x


def _clustering(adj):
    """Local clustering coefficient: triangles through v / (deg choose 2). Triangles are the row sums of
    (A @ A) masked by A, evaluated only at A's non-zeros (native masked row-dot)."""
    import torch

    from ..ops import cpu_struct
    ip, idx, _ = adj.csr()
    idx, _ = cpu_struct.sort_rows(ip, idx.to(ip.dtype))
    rows = torch.repeat_interleave(torch.arange(adj.shape[0], dtype=ip.dtype), (ip[1:] - ip[:-1]).long())
    common = cpu_struct.csr_rowdot(rows, idx, ip, idx, torch.ones(idx.shape[0]), ip, idx, torch.ones(idx.shape[0]))
    tri2 = torch.zeros(adj.shape[0]).index_add_(0, rows.long(), common)   # = 2 * triangles(v)
    deg = (ip[1:] - ip[:-1]).float()
    denom = deg * (deg - 1)
    return torch.where(denom > 0, tri2 / denom.clamp(min=1), torch.zeros_like(denom))


def _node_property(g, property_name, ascending):
    import torch
    adj = _undirected_simple(g)
    if property_name == "popularity":
        val = _pagerank(adj)
    elif property_name == "locality":
        pr = _pagerank(adj)
        pers = torch.zeros(adj.shape[0])
        pers[int(torch.argmax(pr))] = 1.0
        val = _pagerank(adj, pers)
    else:
        val = _clustering(adj)
    return (-val if ascending else val).numpy()


def mask_nodes_by_property(property_values, part_ratios, random_seed=None):
    """Five disjoint boolean masks (in_train / in_valid / in_test / out_valid / out_test): nodes are ordered by
    ``property_values``; the first three parts (in-distribution) are shuffled among themselves, the tail parts keep the
    order so they hold the extreme values (utils.py:mask_nodes_by_property)."""
    import numpy as np
    property_values = np.asarray(property_values)
    n = len(property_values)
    sizes = np.round(n * np.asarray(part_ratios, dtype=float)).astype(int)
    sizes[-1] -= sizes.sum() - n
    rng = np.random.RandomState(random_seed)
    perm = rng.permutation(n)
    ordered = perm[np.argsort(property_values[perm], kind="stable")]
    n_in = int(sizes[:3].sum())
    ordered[:n_in] = rng.permutation(ordered[:n_in])
    names = ["in_train_mask", "in_valid_mask", "in_test_mask", "out_valid_mask", "out_test_mask"]
    parts = np.split(ordered, np.cumsum(sizes))[:-1]
    return {name: generate_mask_tensor(idx2mask(idx, n)) for name, idx in zip(names, parts)}


def add_node_property_split(dataset, part_ratios, property_name, ascending=True, random_seed=None):
    """Attach the five structural-shift masks of :func:`mask_nodes_by_property` to every graph of ``dataset`` using
    'popularity' (PageRank), 'locality' (PageRank personalised at the most popular node) or 'density' (clustering
    coefficient) — Bazhenov et al. 2023 (utils.py:add_node_property_split)."""
    assert property_name in ("popularity", "locality", "density"), \
        "The name of property has to be 'popularity', 'locality', or 'density'"
    assert len(part_ratios) == 5, "part_ratios must contain 5 values"
    for i in range(len(dataset)):
        g = dataset[i]
        masks = mask_nodes_by_property(_node_property(g, property_name, ascending), part_ratios, random_seed)
        for k, m in masks.items():
            g.ndata[k] = m


def __getattr__(name):
    if name == "get_download_dir":
        from .dgl_dataset import get_download_dir
        return get_download_dir
    if name in ("save_tensors", "load_tensors"):
        from . import compat
        return getattr(compat, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
