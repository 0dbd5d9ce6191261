"""File helpers of the on-disk dataset pipeline (parity: graphbolt/internal/utils.py)."""
from __future__ import annotations

import hashlib
import json
import os
import shutil

import numpy as np
import torch

__all__ = ["numpy_save_aligned", "read_data", "save_data", "get_npy_dim", "copy_or_convert_data", "read_edges",
           "calculate_file_hash", "calculate_dir_hash", "check_dataset_change"]


def numpy_save_aligned(path, arr, alignment=4096):
    from .. import numpy_save_aligned as impl
    return impl(path, arr, alignment)


def read_data(path, fmt, in_memory=True):
    """Tensor from a ``torch.save`` file or a ``.npy`` file (memory-mapped unless ``in_memory``)."""
    if fmt == "torch":
        return torch.load(path, weights_only=False)
    if fmt == "numpy":
        if in_memory:
            return torch.from_numpy(np.load(path))
        return torch.from_numpy(np.load(path, mmap_mode="r+"))
    raise RuntimeError(f"Unsupported format: {fmt}")


def save_data(data, path, fmt):
    """Write a tensor / ndarray as ``torch`` or page-aligned ``numpy``."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    if fmt == "numpy":
        arr = data.cpu().numpy() if torch.is_tensor(data) else np.asarray(data)
        numpy_save_aligned(path, np.ascontiguousarray(arr))
    elif fmt == "torch":
        t = data if torch.is_tensor(data) else torch.from_numpy(np.asarray(data))
        torch.save(t.cpu().contiguous(), path)
    else:
        raise RuntimeError(f"Unsupported format: {fmt}")


def get_npy_dim(npy_path):
    """Number of dimensions of a ``.npy`` array, read from its header only."""
    from numpy.lib import format as npf
    with open(npy_path, "rb") as f:
        try:
            version = npf.read_magic(f)
            shape = (npf.read_array_header_1_0(f) if version == (1, 0) else npf.read_array_header_2_0(f))[0]
        except ValueError as exc:
            raise ValueError("Invalid file format") from exc
    return len(shape)


def copy_or_convert_data(input_path, output_path, input_format, output_format="numpy", in_memory=True,
                         is_feature=False, within_int32=False):
    """Bring one input file into the preprocessed directory as ``.npy``: plain copy when nothing has to change, otherwise
    load, (optionally) narrow ids to int32, make 1-D features ``[N, 1]`` and re-save aligned."""
    assert output_format == "numpy", "The output format of the data should be numpy."
    os.makedirs(os.path.dirname(os.path.abspath(output_path)), exist_ok=True)
    data = read_data(input_path, input_format, in_memory)
    changed = False
    if within_int32:
        data, changed = data.to(torch.int32), True
    if is_feature and data.dim() == 1:
        data, changed = data.reshape(-1, 1), True
    if input_format == "numpy" and not changed and data.is_contiguous():
        shutil.copyfile(input_path, output_path)
        return
    save_data(data, output_path, output_format)


def read_edges(dataset_dir, edge_fmt, edge_path):
    """(src, dst) numpy arrays from a ``[2, E]`` ``.npy`` file or a two-column header-less CSV."""
    assert edge_fmt in ("numpy", "csv"), f"`numpy` or `csv` is expected when reading edges but got `{edge_fmt}`."
    full = os.path.join(dataset_dir, edge_path)
    if edge_fmt == "numpy":
        edges = read_data(full, "numpy")
        assert edges.dim() == 2 and edges.shape[0] == 2, f"The shape of edges should be (2, N), but got {tuple(edges.shape)}."
        src, dst = edges.numpy()
        return src, dst
    import pandas as pd
    df = pd.read_csv(full, names=["src", "dst"])
    return df["src"].to_numpy(), df["dst"].to_numpy()


_ALGOS = ("md5", "sha1", "sha224", "sha256", "sha384", "sha512")


def calculate_file_hash(file_path, hash_algo="md5"):
    if hash_algo not in _ALGOS:
        raise ValueError(f"Hash algorithm must be one of: {list(_ALGOS)}, but got `{hash_algo}`.")
    h = getattr(hashlib, hash_algo)()
    with open(file_path, "rb") as f:
        for block in iter(lambda: f.read(1 << 20), b""):
            h.update(block)
    return h.hexdigest()


def calculate_dir_hash(dir_path, hash_algo="md5", ignore=None):
    """{file path: digest} of every file under ``dir_path`` except names in ``ignore``."""
    ignore = {ignore} if isinstance(ignore, str) else set(ignore or ())
    out = {}
    for root, _, files in os.walk(dir_path):
        for name in files:
            if name not in ignore:
                p = os.path.join(root, name)
                out[p] = calculate_file_hash(p, hash_algo)
    return out


def check_dataset_change(dataset_dir, processed_dir):
    """True when the raw dataset no longer matches the digest recorded at preprocessing time (or none was recorded)."""
    record = os.path.join(dataset_dir, processed_dir, "dataset_hash_value.txt")
    if not os.path.exists(record):
        return True
    with open(record) as f:
        old = json.load(f)
    return old != calculate_dir_hash(dataset_dir, ignore="dataset_hash_value.txt")
