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


# ---------------------------------------------------------------------------------------------------------------------
# Structure construction kernels (csrc/cuda/structure_build.cu): radix-sort COO -> CSR with the edge-id permutation,
# CSR transpose, coalesce with value reduction, hash-table SpGEMM. No host synchronisation except where the OUTPUT size
# is data dependent (coalesce / SpGEMM read one device counter, like torch.unique does).
# ---------------------------------------------------------------------------------------------------------------------
def cuda_coo_to_csr(row, col, num_rows):
    """(indptr, indices, eids): rows sorted stably, ``eids[p]`` = input position of output entry p."""
    idt = row.dtype
    row, col = row.contiguous(), col.to(idt).contiguous()
    nnz = row.shape[0]
    dev = row.device
    indptr = torch.empty(num_rows + 1, dtype=idt, device=dev)
    indices = torch.empty(nnz, dtype=idt, device=dev)
    eids = torch.empty(nnz, dtype=idt, device=dev)
    loader.cuda_lib().fn("dglb200_cuda_coo2csr", "ippllpppp")(id_code(idt), row.data_ptr(), col.data_ptr(), nnz, int(num_rows),
                                                             indptr.data_ptr(), indices.data_ptr(), eids.data_ptr(),
                                                             loader.stream_ptr())
    return indptr, indices, eids


def cuda_csr_transpose(indptr, indices, num_rows, num_cols):
    """CSR of the transpose + for each transposed entry its position in the input."""
    idt = indptr.dtype
    nnz = indices.shape[0]
    dev = indptr.device
    out_p = torch.empty(num_cols + 1, dtype=idt, device=dev)
    out_i = torch.empty(nnz, dtype=idt, device=dev)
    out_e = torch.empty(nnz, dtype=idt, device=dev)
    loader.cuda_lib().fn("dglb200_cuda_csr_transpose", "ipplllpppp")(
        id_code(idt), indptr.contiguous().data_ptr(), indices.contiguous().data_ptr(), nnz, int(num_rows), int(num_cols),
        out_p.data_ptr(), out_i.data_ptr(), out_e.data_ptr(), loader.stream_ptr())
    return out_p, out_i, out_e


def cuda_coalesce(row, col, val, num_rows, num_cols, want_first=False):
    """Entries sorted by (row, col) with duplicates merged (values summed in fp32).
    Returns (row, col, val or None[, first]) trimmed to the number of distinct coordinates."""
    idt = row.dtype
    nnz = row.shape[0]
    dev = row.device
    vdim = 0
    v32 = None
    if val is not None:
        v32 = val.reshape(nnz, -1).float().contiguous()
        vdim = v32.shape[1]
    out_r = torch.empty(nnz, dtype=idt, device=dev)
    out_c = torch.empty(nnz, dtype=idt, device=dev)
    out_v = torch.zeros((nnz, max(vdim, 1)), dtype=torch.float32, device=dev) if val is not None else None
    first = torch.empty(nnz, dtype=torch.int64, device=dev) if want_first else None
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    loader.cuda_lib().fn("dglb200_cuda_coo_coalesce", "ipppilllpppppp")(
        id_code(idt), row.contiguous().data_ptr(), col.to(idt).contiguous().data_ptr(), 0 if v32 is None else v32.data_ptr(), vdim,
        nnz, int(num_rows), int(num_cols), out_r.data_ptr(), out_c.data_ptr(), 0 if out_v is None else out_v.data_ptr(),
        0 if first is None else first.data_ptr(), count.data_ptr(), loader.stream_ptr())
    n = int(count.item())
    res = [out_r[:n], out_c[:n], None if out_v is None else out_v[:n].reshape((n,) + tuple(val.shape[1:])).to(val.dtype)]
    if want_first:
        res.append(first[:n])
    return tuple(res)


def cuda_spgemm(ap, ai, av, bp, bi, bv, num_rows, num_cols):
    """C = A @ B (CSR x CSR, fp32 accumulation in per-row hash tables). Returns canonical COO (row, col, val) sorted by
    (row, col). ``av`` / ``bv`` may be None (pattern product)."""
    idt = ap.dtype
    dev = ap.device
    lib = loader.cuda_lib()
    cap = torch.empty(num_rows, dtype=torch.int64, device=dev)
    off = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
    lib.fn("dglb200_cuda_spgemm_plan", "ippplppp")(id_code(idt), ap.data_ptr(), ai.data_ptr(), bp.data_ptr(), int(num_rows),
                                                 cap.data_ptr(), off.data_ptr(), loader.stream_ptr())
    total = int(off[-1].item())
    keys = torch.full((max(total, 1),), -1, dtype=torch.int64, device=dev)
    vals = torch.zeros(max(total, 1), dtype=torch.float32, device=dev)
    row_nnz = torch.empty(num_rows, dtype=torch.int64, device=dev)
    out_off = torch.empty(num_rows + 1, dtype=torch.int64, device=dev)
    f32 = lambda t: None if t is None else t.float().contiguous()  # noqa: E731
    av32, bv32 = f32(av), f32(bv)
    lib.fn("dglb200_cuda_spgemm_accumulate", "ipppppplpppppp" + "p")(
        id_code(idt), ap.data_ptr(), ai.data_ptr(), 0 if av32 is None else av32.data_ptr(), bp.data_ptr(), bi.data_ptr(),
        0 if bv32 is None else bv32.data_ptr(), int(num_rows), off.data_ptr(), cap.data_ptr(), keys.data_ptr(), vals.data_ptr(),
        row_nnz.data_ptr(), out_off.data_ptr(), loader.stream_ptr())
    nnz = int(out_off[-1].item())
    out_r = torch.empty(nnz, dtype=idt, device=dev)
    out_c = torch.empty(nnz, dtype=idt, device=dev)
    out_v = torch.empty(nnz, dtype=torch.float32, device=dev)
    if nnz:
        lib.fn("dglb200_cuda_spgemm_compact", "ilpppppppp" + "p")(id_code(idt), int(num_rows), off.data_ptr(), cap.data_ptr(),
                                                                out_off.data_ptr(), keys.data_ptr(), vals.data_ptr(),
                                                                out_r.data_ptr(), out_c.data_ptr(), out_v.data_ptr(),
                                                                loader.stream_ptr())
        out_r, out_c, out_v = cuda_coalesce(out_r, out_c, out_v, num_rows, num_cols)   # canonical (row, col) order
    return out_r, out_c, out_v
