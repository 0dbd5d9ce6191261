"""segment_mm / gather_mm with autograd.

Parity: /root/reference/python/dgl/backend/pytorch/sparse.py:968-1020 (SEGMENTMM, GATHERMM) and
/root/reference/src/array/cuda/gather_mm.cu. The reference loops over relations on the host (one cuBLAS call each,
plus a D2H copy of ``seglen``). Here 16-bit CUDA inputs run ONE persistent tcgen05 kernel per product
(csrc/cuda/grouped_mm_tc.cu): the segment table and the tile list stay on the device, so there is no host
synchronisation, and the three products of a training step are

    out = segment_mm(A, W)          seg_mm_tc_kernel     (A rows optionally gathered through an index)
    dA  = segment_mm(dOut, W^T)     seg_mm_tc_kernel     (W is already the K-major operand of this product)
    dW  = A_seg^T @ dOut_seg        seg_mm_dw_tc_kernel  (MN-major operands, fp32 red.add into dW)

fp32 / CPU inputs and shapes the kernel does not take (K or M not a multiple of 8) use the torch formulation
(sort by relation + padded bmm), which is also the numerical reference of the tests.
"""
from __future__ import annotations

import torch

from .._native import loader
from ..base import dtype_code

_TILE_M = 128
_DW_ROWS = 2048


def _segment_ids(seglen, n, device):
    seglen = seglen.to(device)
    return torch.repeat_interleave(torch.arange(seglen.shape[0], device=device), seglen.long(), output_size=n)


def _tc_ok(a, b):
    return (a.is_cuda and b.is_cuda and a.dtype in (torch.bfloat16, torch.float16) and b.dtype == a.dtype
            and a.dim() == 2 and b.dim() == 3 and a.shape[1] % 8 == 0 and b.shape[2] % 8 == 0 and a.shape[0] > 0)


def _offsets(seglen, per):
    """[R+1] int64 prefix sums of seglen and of ceil(seglen / per) - device ops only."""
    seglen = seglen.long()
    z = seglen.new_zeros(1)
    seg_off = torch.cat([z, torch.cumsum(seglen, 0)])
    unit_off = torch.cat([z, torch.cumsum((seglen + per - 1) // per, 0)])
    return seg_off, unit_off


def _seg_mm_tc(a, idx_a, n_rows, bt, seg_off, tile_off, out_f32=False):
    """out[i] = a[idx_a[i] or i] @ bt[r(i)]^T with bt: [R, Nout, K] contiguous."""
    R, Nout, K = bt.shape
    out = torch.empty((n_rows, Nout), dtype=torch.float32 if out_f32 else a.dtype, device=a.device)
    max_tiles = (n_rows + _TILE_M - 1) // _TILE_M + R
    loader.cuda_lib().fn("dglb200_cuda_segment_mm_tc", "iplllppppliiipip")(
        dtype_code(a.dtype), a.data_ptr(), a.shape[0], n_rows, a.stride(0), 0 if idx_a is None else idx_a.data_ptr(), bt.data_ptr(),
        seg_off.data_ptr(), tile_off.data_ptr(), max_tiles, R, K, Nout, out.data_ptr(), 1 if out_f32 else 0,
        loader.stream_ptr())
    return out


def _seg_mm_dw_tc(a, idx_a, dout, seg_off, item_off, R):
    n_rows, Nout = dout.shape
    K = a.shape[1]
    dw = torch.zeros((R, K, Nout), dtype=torch.float32, device=a.device)
    max_items = (n_rows + _DW_ROWS - 1) // _DW_ROWS + R
    loader.cuda_lib().fn("dglb200_cuda_segment_mm_dw_tc", "iplppllppliiipp")(
        dtype_code(a.dtype), a.data_ptr(), a.stride(0), 0 if idx_a is None else idx_a.data_ptr(), dout.data_ptr(),
        n_rows, dout.stride(0), seg_off.data_ptr(), item_off.data_ptr(), max_items, R, K, Nout, dw.data_ptr(),
        loader.stream_ptr())
    return dw


class _SegmentMMTC(torch.autograd.Function):
    """Rows of the product are in relation-sorted order; ``idx_a`` (optional, int64) gathers the A rows."""

    @staticmethod
    def forward(ctx, a, b, seglen, idx_a):
        a = a if a.stride(1) == 1 and a.stride(0) % 8 == 0 and a.data_ptr() % 16 == 0 else a.contiguous()
        seglen = seglen.to(a.device)
        n_rows = a.shape[0] if idx_a is None else idx_a.shape[0]
        seg_off, tile_off = _offsets(seglen, _TILE_M)
        out = _seg_mm_tc(a, idx_a, n_rows, b.transpose(1, 2).contiguous(), seg_off, tile_off)
        ctx.save_for_backward(a, b, seglen, seg_off, tile_off, idx_a if idx_a is not None else torch.empty(0))
        ctx.has_idx = idx_a is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b, seglen, seg_off, tile_off, idx_a = ctx.saved_tensors
        idx_a = idx_a if ctx.has_idx else None
        dout = dout.contiguous()
        R = b.shape[0]
        da = db = None
        if ctx.needs_input_grad[0]:
            # dA_sorted = dOut @ W[r]^T: the K-major "B^T" operand of this product is W itself ([R, K, M])
            da_sorted = _seg_mm_tc(dout, None, dout.shape[0], b.contiguous(), seg_off, tile_off)
            if idx_a is None:
                da = da_sorted
            else:
                da = torch.zeros_like(a).index_add_(0, idx_a, da_sorted)
        if ctx.needs_input_grad[1]:
            _, item_off = _offsets(seglen, _DW_ROWS)
            db = _seg_mm_dw_tc(a, idx_a, dout, seg_off, item_off, R).to(b.dtype)
        return da, db, None, None


def _segment_mm_torch(a, b, seglen_a):
    if not a.is_cuda:
        outs, off = [], 0
        for r, n in enumerate(seglen_a.tolist()):
            outs.append(a[off:off + n] @ b[r])
            off += n
        return torch.cat(outs, 0) if outs else a.new_zeros((0, b.shape[-1]))
    idx = _segment_ids(seglen_a, a.shape[0], a.device)
    return _gather_mm_torch(a, b, idx)


def _gather_mm_torch(a, b, idx_b):
    N, K = a.shape
    R, _, M = b.shape
    idx_b = idx_b.long()
    if R * M <= 4096 and N * R * M * K < 2 ** 33:
        full = torch.einsum("nk,rkm->nrm", a, b)
        return full.gather(1, idx_b.view(-1, 1, 1).expand(-1, 1, M)).squeeze(1)
    order = torch.sort(idx_b, stable=True)[1]
    counts = torch.bincount(idx_b, minlength=R)
    maxc = int(counts.max())
    starts = torch.cumsum(counts, 0) - counts
    pos = torch.arange(N, device=a.device) - torch.repeat_interleave(starts, counts, output_size=N)
    sorted_rel = idx_b[order]
    padded = a.new_zeros((R, maxc, K))
    padded[sorted_rel, pos] = a[order]
    outp = torch.bmm(padded, b)
    out = a.new_empty((N, M))
    out[order] = outp[sorted_rel, pos]
    return out


def segment_mm(a, b, seglen_a):
    """out[seg_r] = a[seg_r] @ b[r];  a:[N,K] b:[R,K,M] seglen_a:[R] (host or device tensor)."""
    from ..ops.sparse_autograd import _autocast_args
    (a, b), ctx = _autocast_args(a, b)       # under autocast both operands run in the autocast dtype -> tcgen05 path
    with ctx:
        return _segment_mm(a, b, seglen_a)


def _segment_mm(a, b, seglen_a):
    if a.shape[0] == 0:
        return a.new_zeros((0, b.shape[-1]))
    if _tc_ok(a, b):
        return _SegmentMMTC.apply(a, b, seglen_a, None)
    return _segment_mm_torch(a, b, seglen_a)


def gather_mm(a, b, idx_b):
    """out[i] = a[i] @ b[idx_b[i]];  a:[N,K] b:[R,K,M]. Unsorted relation ids: rows are bucketed by relation on the
    device (stable sort + bincount), the A rows are gathered inside the GEMM kernel, and the result is scattered back."""
    from ..ops.sparse_autograd import _autocast_args
    (a, b), ctx = _autocast_args(a, b)
    with ctx:
        return _gather_mm(a, b, idx_b)


def _gather_mm(a, b, idx_b):
    N, K = a.shape
    R, _, M = b.shape
    if N == 0:
        return a.new_zeros((0, M))
    if R == 1:
        return a @ b[0]
    if _tc_ok(a, b):
        idx_b = idx_b.long()
        order = torch.sort(idx_b, stable=True)[1]
        seglen = torch.bincount(idx_b, minlength=R)
        out_sorted = _SegmentMMTC.apply(a, b, seglen, order)
        return torch.empty_like(out_sorted).index_copy_(0, order, out_sorted)
    return _gather_mm_torch(a, b, idx_b)
