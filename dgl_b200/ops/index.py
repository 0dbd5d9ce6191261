"""Row gather / scatter incl. UVA (GPU threads reading pinned host memory).

Parity: /root/reference/src/array/cuda/uvm/array_index_select_uvm.cu,
graphbolt/src/cuda/index_select_impl.cu:34-205.
"""
from __future__ import annotations

import torch

from ..base import DGLError


def index_select(tensor, index):
    """out[i] = tensor[index[i]]. If ``tensor`` is pinned host memory and ``index`` is a CUDA
    tensor the gather runs as a UVA kernel and returns a CUDA tensor."""
    if index.is_cuda and not tensor.is_cuda:
        if not tensor.is_pinned():
            raise DGLError("UVA index_select requires the source tensor to be pinned (pin_memory_inplace)")
        from .cuda_struct import uva_index_select
        return uva_index_select(tensor, index)
    if index.is_cuda and tensor.is_cuda:
        from .cuda_struct import gather_rows
        return gather_rows(tensor, index)
    return tensor[index.long()]


def index_scatter(dest, index, source):
    if index.is_cuda and not dest.is_cuda:
        if not dest.is_pinned():
            raise DGLError("UVA index_scatter requires the destination tensor to be pinned")
        from .cuda_struct import uva_index_scatter
        return uva_index_scatter(dest, index, source)
    dest[index.long()] = source
