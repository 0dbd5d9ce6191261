"""Pinned-host (UVA) helpers.

Parity: /root/reference/python/dgl/utils/pin_memory.py and the UVA gather kernels
(/root/reference/src/array/cuda/uvm/array_index_select_uvm.cu,
graphbolt/src/cuda/index_select_impl.cu). ``gather_pinned_tensor_rows`` runs an
sm_100a kernel whose threads read host-pinned rows directly over PCIe/C2C.
"""
from __future__ import annotations

import torch

from ..base import DGLError


def pin_memory_inplace(tensor):
    """cudaHostRegister ``tensor``'s memory in place; returns the tensor."""
    if tensor.is_cuda:
        raise DGLError("only CPU tensors can be pinned")
    if tensor.is_pinned() or tensor.numel() == 0 or not torch.cuda.is_available():
        return tensor
    if not tensor.is_contiguous():
        raise DGLError("only contiguous tensors can be pinned in place")
    rc = torch.cuda.cudart().cudaHostRegister(tensor.data_ptr(), tensor.numel() * tensor.element_size(), 0)
    if int(rc) != 0:
        raise DGLError(f"cudaHostRegister failed: {int(rc)}")
    return tensor


def unpin_memory_inplace(tensor):
    if tensor.numel() == 0 or not torch.cuda.is_available():
        return tensor
    torch.cuda.cudart().cudaHostUnregister(tensor.data_ptr())
    return tensor


def gather_pinned_tensor_rows(tensor, rows):
    """out[i] = tensor[rows[i]] where ``tensor`` lives in pinned host memory and ``rows``
    on the GPU. Returns a CUDA tensor."""
    from ..ops.index import index_select
    return index_select(tensor, rows)


def scatter_pinned_tensor_rows(dest, rows, source):
    """dest[rows[i]] = source[i]; ``dest`` pinned host, rows/source on GPU."""
    from ..ops.index import index_scatter
    index_scatter(dest, rows, source)


def host_device_pointer(tensor):
    """Device-visible address of a host-pinned tensor (UVA): what kernels dereference to read host memory directly."""
    import ctypes

    from .._native import loader
    if not tensor.is_pinned():
        raise DGLError("tensor must be in pinned (page-locked) host memory for UVA access")
    out = ctypes.c_void_p()
    loader.cuda_lib().fn("dglb200_cuda_host_device_ptr", "pp")(tensor.data_ptr(), ctypes.addressof(out))
    return out.value
