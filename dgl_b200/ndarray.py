"""``dgl.ndarray`` surface (parity: /root/reference/python/dgl/ndarray.py).

The reference carries its own TVM-derived ``NDArray`` and converts to/from framework tensors through DLPack. This
package has no second tensor type: every array *is* a ``torch.Tensor`` (device memory from PyTorch's caching
allocator, shared memory from ``utils.shared_mem``). The names below keep code written against ``dgl.ndarray``
working; conversions are identities or DLPack round trips."""
from __future__ import annotations

import numpy as np
import torch

from .utils.shared_mem import create_shared_mem_array, get_shared_mem_array, shm_exists  # noqa: F401

__all__ = ["NDArray", "array", "cpu", "gpu", "zerocopy_from_numpy", "cast_to_signed", "SparseFormat", "SparseMatrix",
           "create_shared_mem_array", "get_shared_mem_array", "exist_shared_mem_array", "from_dlpack", "to_dlpack"]

NDArray = torch.Tensor


def cpu(dev_id=0):
    return torch.device("cpu")


def gpu(dev_id=0):
    return torch.device("cuda", dev_id)


def array(arr, ctx=None):
    """Array-like -> tensor on ``ctx`` (default CPU)."""
    t = arr if torch.is_tensor(arr) else torch.as_tensor(np.asarray(arr))
    return t.to(ctx) if ctx is not None else t


def zerocopy_from_numpy(np_data):
    """Tensor sharing memory with a numpy array."""
    return torch.from_numpy(np_data)


def cast_to_signed(arr):
    """uint8/uint16/uint32/uint64 views reinterpreted as the signed type of the same width (ids are signed here)."""
    table = {torch.uint8: torch.int8}
    for u, s in (("uint16", "int16"), ("uint32", "int32"), ("uint64", "int64")):
        if hasattr(torch, u):
            table[getattr(torch, u)] = getattr(torch, s)
    return arr.view(table[arr.dtype]) if arr.dtype in table else arr


def exist_shared_mem_array(name):
    return bool(shm_exists(name))


def from_dlpack(capsule):
    return torch.utils.dlpack.from_dlpack(capsule)


def to_dlpack(tensor):
    return torch.utils.dlpack.to_dlpack(tensor)


class SparseFormat:
    """Format codes of the legacy sparse-matrix C API."""
    ANY = 0
    COO = 1
    CSR = 2
    CSC = 3
    FORMAT2STR = {0: "ANY", 1: "COO", 2: "CSR", 3: "CSC"}


class SparseMatrix:
    """Legacy container ``(format, indices arrays, flags)`` returned by adjacency getters of the C API."""

    def __init__(self, fmt, indices, flags=()):
        self._format, self._indices, self._flags = fmt, list(indices), list(flags)

    @property
    def format(self):
        return self._format

    @property
    def num_indices(self):
        return len(self._indices)

    @property
    def indices(self):
        return self._indices

    @property
    def num_flags(self):
        return len(self._flags)

    @property
    def flags(self):
        return [bool(f) for f in self._flags]
