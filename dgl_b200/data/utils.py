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
