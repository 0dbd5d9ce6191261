"""GraphBolt base types and tensor utilities (parity: /root/reference/python/dgl/graphbolt/base.py,
internal/utils.py; native graphbolt/src/{expand_indptr,isin,index_select}.cu replaced by torch ops +
the sm_100a gather kernels)."""
from __future__ import annotations

from dataclasses import dataclass

import torch

ORIGINAL_EDGE_ID = "_ORIGINAL_EDGE_ID"
CANONICAL_ETYPE_DELIMITER = ":"
CSCFormatBaseStr = "CSCFormatBase"


def etype_tuple_to_str(c_etype):
    if isinstance(c_etype, str):
        return c_etype
    assert isinstance(c_etype, tuple) and len(c_etype) == 3, "Passed-in canonical etype should be in format of (str, str, str)."
    return CANONICAL_ETYPE_DELIMITER.join(c_etype)


def etype_str_to_tuple(c_etype):
    if isinstance(c_etype, tuple):
        return c_etype
    ret = tuple(c_etype.split(CANONICAL_ETYPE_DELIMITER))
    assert len(ret) == 3, f"Passed-in canonical etype should be in format of 'str:str:str'. But got {c_etype}."
    return ret


def seed_type_str_to_ntypes(seed_type, seed_size):
    ntypes = seed_type.split(CANONICAL_ETYPE_DELIMITER)
    if len(ntypes) == 3:
        return [ntypes[0], ntypes[2]] if seed_size == 2 else ntypes
    return ntypes


def isin(elements, test_elements):
    """Membership test (sorted + binary search; reference graphbolt/src/cuda/isin.cu)."""
    assert elements.dim() == 1 and test_elements.dim() == 1, "Elements and test_elements should be 1-D tensors."
    if test_elements.numel() == 0:
        return torch.zeros_like(elements, dtype=torch.bool)
    s = torch.sort(test_elements)[0]
    pos = torch.searchsorted(s, elements).clamp(max=s.numel() - 1)
    return s[pos] == elements


def expand_indptr(indptr, dtype=None, node_ids=None, output_size=None):
    """CSR row pointer -> per-nonzero row ids (reference graphbolt/src/cuda/expand_indptr.cu)."""
    deg = (indptr[1:] - indptr[:-1]).long()
    base = torch.arange(deg.shape[0], device=indptr.device, dtype=dtype or indptr.dtype) if node_ids is None else node_ids
    return torch.repeat_interleave(base, deg, output_size=output_size)


def index_select(tensor, index):
    """Row gather; UVA kernel when ``tensor`` is pinned host memory and ``index`` is on the GPU."""
    from ..ops.index import index_select as _is
    if index.is_cuda or tensor.is_cuda:
        if not tensor.is_cuda and not tensor.is_pinned():
            return tensor[index.cpu().long()].to(index.device)
        return _is(tensor, index.to("cuda") if not index.is_cuda else index)
    return tensor[index.long()]


def apply_to(x, device, non_blocking=False):
    return x.to(device, non_blocking=non_blocking) if hasattr(x, "to") else x


@dataclass
class CSCFormatBase:
    """Sampled bipartite block in CSC: ``indptr`` over destination (seed) nodes, ``indices`` = source ids."""
    indptr: torch.Tensor = None
    indices: torch.Tensor = None

    def __post_init__(self):
        if self.indptr is not None and self.indices is not None:
            assert self.indptr.dim() == 1 and self.indices.dim() == 1

    def to(self, device, non_blocking=False):
        return CSCFormatBase(self.indptr.to(device, non_blocking=non_blocking), self.indices.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return CSCFormatBase(self.indptr.pin_memory(), self.indices.pin_memory())

    def is_pinned(self):
        return self.indptr.is_pinned() and self.indices.is_pinned()

    def __repr__(self):
        return f"CSCFormatBase(indptr={self.indptr},\n              indices={self.indices},\n)"


def __getattr__(name):
    # names the reference keeps in this module but that live elsewhere in this package
    if name.startswith("__"):
        raise AttributeError(name)
    import importlib
    for modname in ('dgl_b200.graphbolt.dataloader', 'dgl_b200.graphbolt.compat'):
        mod = importlib.import_module(modname)
        if name in vars(mod):
            return vars(mod)[name]
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
