"""Python bindings of the sm_100a structure kernels (csrc/cuda/structure.cu)."""
from __future__ import annotations

import ctypes

import torch

from .._native import loader
from ..base import DGLError, id_code

_INT_MAX = 2 ** 31 - 1


def _pow2_cap(n):
    cap = 16
    while cap < 2 * n + 2:
        cap <<= 1
    return cap


def _scan_bytes(n):
    lib = loader.cuda_lib()
    f = lib.raw("dglb200_cuda_scan_temp_bytes", ctypes.c_int64, [ctypes.c_int64])
    return int(f(int(n)))


def cuda_unique_compact(first, ids, return_device_count=False):
    """Order-preserving unique of [first ++ ids]; returns (unique, local_ids_of_ids)."""
    dev, idt = ids.device, ids.dtype
    if first is None:
        first = torch.empty(0, dtype=idt, device=dev)
    first = first.to(idt).contiguous()
    ids = ids.contiguous()
    nfirst, n = first.shape[0], ids.shape[0]
    N = nfirst + n
    if N >= _INT_MAX:
        raise DGLError("unique_compact supports fewer than 2^31 ids per call")
    cap = _pow2_cap(N)
    keys = torch.empty(cap, dtype=idt, device=dev)
    minpos = torch.empty(cap, dtype=torch.int32, device=dev)
    local = torch.empty(cap, dtype=torch.int32, device=dev)
    work = torch.empty((3, max(N, 1)), dtype=torch.int32, device=dev)
    sb = _scan_bytes(max(N, 1))
    temp = torch.empty(max(sb, 1), dtype=torch.uint8, device=dev)
    out_unique = torch.empty(max(N, 1), dtype=idt, device=dev)
    out_local = torch.empty(n, dtype=idt, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    loader.cuda_lib().fn("dglb200_cuda_unique_compact", "iplplppppl" + "pppplppplpp")(
        id_code(idt), first.data_ptr(), nfirst, ids.data_ptr(), n, 0, keys.data_ptr(), minpos.data_ptr(),
        local.data_ptr(), cap, work[0].data_ptr(), work[1].data_ptr(), work[2].data_ptr(), temp.data_ptr(), sb,
        out_unique.data_ptr(), out_local.data_ptr(), count.data_ptr(), 0, 0, loader.stream_ptr())
    if return_device_count:
        return out_unique, out_local, count
    return out_unique[: int(count.item())], out_local


def cuda_map_ids(universe, ids):
    dev, idt = ids.device, ids.dtype
    universe = universe.to(device=dev, dtype=idt).contiguous()
    nu, n = universe.shape[0], ids.shape[0]
    cap = _pow2_cap(nu)
    keys = torch.empty(cap, dtype=idt, device=dev)
    vals = torch.empty(cap, dtype=torch.int32, device=dev)
    out = torch.empty(n, dtype=idt, device=dev)
    loader.cuda_lib().fn("dglb200_cuda_map_ids", "iplplpplpp")(
        id_code(idt), universe.data_ptr(), nu, ids.contiguous().data_ptr(), n, keys.data_ptr(), vals.data_ptr(), cap,
        out.data_ptr(), loader.stream_ptr())
    return out


_uva_sort_min_rows = [1 << 14]


def _gather(src, index, uva, max_blocks=0):
    if index.dtype not in (torch.int32, torch.int64):
        index = index.long()
    index = index.contiguous()
    dst_pos = None
    if uva and index.shape[0] >= _uva_sort_min_rows[0]:
        # host reads in ascending address order (DRAM page / IOMMU-TLB locality, the reference sorts too:
        # graphbolt/src/cuda/index_select_impl.cu:129-134); rows are written straight to their original positions
        index, dst_pos = torch.sort(index)
    if not src.is_contiguous():
        raise DGLError("gather source must be contiguous")
    n = index.shape[0]
    out = torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=index.device)
    row_bytes = src.element_size()
    for s in src.shape[1:]:
        row_bytes *= s
    if n > 0 and row_bytes > 0:
        loader.cuda_lib().fn("dglb200_cuda_gather_rows", "ipplpllii" + "pp")(
            id_code(index.dtype), src.data_ptr(), index.data_ptr(), n, out.data_ptr(), row_bytes, src.shape[0],
            1 if uva else 0, max_blocks, 0 if dst_pos is None else dst_pos.data_ptr(), loader.stream_ptr())
    return out


def gather_rows(src, index):
    """out[i] = src[index[i]] on the device (vectorised sub-warp-per-row gather)."""
    return _gather(src, index, False)


_max_uva_blocks = [0]


def set_max_uva_threads(n):
    """Cap the number of threads used by UVA gathers (reference: graphbolt ``max_uva_threads``)."""
    _max_uva_blocks[0] = max(1, int(n) // 256) if n else 0


def uva_index_select(src, index):
    """Gather rows of a host-pinned tensor from the GPU."""
    return _gather(src, index, True, _max_uva_blocks[0])


def uva_index_scatter(dest, index, source):
    index = index.contiguous()
    source = source.contiguous()
    row_bytes = dest.element_size()
    for s in dest.shape[1:]:
        row_bytes *= s
    loader.cuda_lib().fn("dglb200_cuda_scatter_rows", "ipplplip")(
        id_code(index.dtype), dest.data_ptr(), index.data_ptr(), index.shape[0], source.data_ptr(), row_bytes, 1,
        loader.stream_ptr())


def scatter_rows(dest, index, source):
    """dest[index[i]] = source[i] on the device (index int32/int64, unique)."""
    index = index.contiguous()
    source = source.contiguous()
    row_bytes = dest.element_size()
    for s_ in dest.shape[1:]:
        row_bytes *= s_
    if index.shape[0] and row_bytes:
        loader.cuda_lib().fn("dglb200_cuda_scatter_rows", "ipplplip")(
            id_code(index.dtype), dest.data_ptr(), index.data_ptr(), index.shape[0], source.data_ptr(), row_bytes, 0,
            loader.stream_ptr())
    return dest
