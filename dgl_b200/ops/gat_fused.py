"""Fused GAT attention + aggregation with autograd (csrc/cuda/gat_fused.cu).

out[v, h, :] = sum_{u -> v} softmax_u(LeakyReLU(el[u, h] + er[v, h])) * ft[u, h, :]

Forward: one online-softmax pass over the CSC rows (no [E, H] score / attention tensors). Backward: one source-parallel
pass over the CSR rows that recomputes the attention weights from the saved per-(v, h) max and normaliser.
The unfused formulation (reference gatconv.py:300-340) stays available as the fallback and the test oracle."""
from __future__ import annotations

import torch

from .._native import loader
from ..base import dtype_code, id_code


_ENABLED = [__import__("os").environ.get("DGLB200_GAT_FUSED", "1") != "0"]      # A/B switch


def supported(el, ft, num_heads, head_dim):
    if not _ENABLED[0]:
        return False
    if not (ft.is_cuda and ft.dtype in (torch.float32, torch.bfloat16, torch.float16) and el.dtype == ft.dtype):
        return False
    vec = 16 // ft.element_size()
    if head_dim % vec:
        return False
    lph = head_dim // vec
    if num_heads * lph <= 32:
        return num_heads == 1 or (lph & (lph - 1)) == 0
    # wider rows run as column tiles of 32 vectors: heads must not straddle a tile
    return num_heads > 1 and lph <= 32 and (lph & (lph - 1)) == 0 and num_heads * lph <= 512


class _FusedGAT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ug, el, er, ft, slope):
        # ug: UnitGraph whose CSC has no edge-id indirection (a destination-sorted view)
        indptr, indices, eids = ug.csc()
        assert eids is None, "fused GAT needs a destination-sorted relation"
        n_dst, H = er.shape[0], er.shape[1]
        D = ft.shape[2]
        el, er, ft = el.contiguous(), er.contiguous(), ft.contiguous()
        out = torch.empty((n_dst, H, D), dtype=ft.dtype, device=ft.device)
        m = torch.empty((n_dst, H), dtype=torch.float32, device=ft.device)
        l = torch.empty((n_dst, H), dtype=torch.float32, device=ft.device)
        loader.cuda_lib().fn("dglb200_cuda_gat_fwd", "iipplpppiifpppp")(
            dtype_code(ft.dtype), id_code(indptr.dtype), indptr.data_ptr(), indices.data_ptr(), n_dst, el.data_ptr(),
            er.data_ptr(), ft.data_ptr(), H, D, float(slope), out.data_ptr(), m.data_ptr(), l.data_ptr(), loader.stream_ptr())
        ctx.ug, ctx.slope = ug, float(slope)
        ctx.save_for_backward(el, er, ft, out, m, l)
        return out

    @staticmethod
    def backward(ctx, dout):
        el, er, ft, out, m, l = ctx.saved_tensors
        dout = dout.contiguous()
        H, D = er.shape[1], ft.shape[2]
        indptr_r, indices_r, _ = ctx.ug.csr()          # rows = sources; built once, cached on the view
        delta = (out.float() * dout.float()).sum(-1)   # sum_u a_uv <ft_u, dOut_v> = <out_v, dOut_v>
        aux = torch.stack([er.float(), m, 1.0 / l, delta], -1).contiguous()
        n_src = ft.shape[0]
        d_ft = torch.empty_like(ft)
        d_el = torch.empty((n_src, H), dtype=torch.float32, device=ft.device)
        d_er = torch.zeros((er.shape[0], H), dtype=torch.float32, device=ft.device)
        loader.cuda_lib().fn("dglb200_cuda_gat_bwd", "iipplppppiifpppp")(
            dtype_code(ft.dtype), id_code(indptr_r.dtype), indptr_r.data_ptr(), indices_r.data_ptr(), n_src, el.data_ptr(),
            ft.data_ptr(), dout.data_ptr(), aux.data_ptr(), H, D, ctx.slope, d_ft.data_ptr(), d_el.data_ptr(),
            d_er.data_ptr(), loader.stream_ptr())
        return None, d_el.to(el.dtype), d_er.to(er.dtype), d_ft, None


def fused_gat(ug, el, er, ft, negative_slope):
    """el: [N_src, H], er: [N_dst, H], ft: [N_src, H, D] -> [N_dst, H, D]."""
    from .sparse_autograd import _autocast_args
    (el, er, ft), ctx = _autocast_args(el, er, ft)
    with ctx:
        return _FusedGAT.apply(ug, el, er, ft, negative_slope)
