"""tcgen05 fused SAGE layer for the GENERIC API: ``dgl.nn.SAGEConv(mean)`` on a sampled block (message-flow graph).

The block's CSC (every destination has at most ``fanout`` in-edges) is turned into the ELL layout of the engine kernel
and the layer runs as ONE kernel (csrc/cuda/sage_tc.cu): gather + mean + [self | neigh] concat in shared memory ->
tcgen05.mma against [W_self | W_neigh]^T (TMA) -> bias (+ReLU) epilogue. The backward is two more tcgen05 kernels
(csrc/cuda/sage_bwd_tc.cu): the split-K weight gradient with MN-major operands and the input gradient whose epilogue
scatters through the block (``red.global.add.v4.f32``). These are the kernels ``FusedSAGETrainer`` replays inside its
CUDA graph, made available to models written against ``dgl.nn``.

Requirements (otherwise SAGEConv keeps its g-SpMM formulation): CUDA, bf16 activations (e.g. under
``torch.autocast(dtype=torch.bfloat16)``) or fp32 activations (3xTF32 split products, fp32-level accuracy), block whose
destination nodes are the first rows of the source nodes, max in-degree <= 28, out_feats <= 256, in_feats <= 256,
no edge weights.
"""
from __future__ import annotations

import os

import torch

from ..ops import fused as fused_ops

MAX_FANOUT = 28
ENABLED = [os.environ.get("DGLB200_FUSED_SAGECONV", "1") != "0"]   # switch for A/B measurements


def _pad_to(n, m):
    return (n + m - 1) // m * m


def block_to_ell(block):
    """(nbr int32 [n_dst, fanout], cnt int32 [n_dst], fanout) of a single-relation block, cached on the block."""
    cached = getattr(block, "_ell_cache", None)
    if cached is not None:                                         # attached by the one-sync sampler, or computed below
        return cached if cached[0] is None or cached[2] <= MAX_FANOUT else (None, None, cached[2])
    indptr, indices, _ = block._graphs[0].csc()
    deg = (indptr[1:] - indptr[:-1]).to(torch.int32)
    n_dst = deg.shape[0]
    fan = int(deg.max()) if n_dst else 0                       # one host sync per block (cached)
    if fan == 0 or fan > MAX_FANOUT:
        res = (None, None, fan)
    else:
        nbr = torch.zeros((n_dst, fan), dtype=torch.int32, device=indices.device)
        mask = torch.arange(fan, device=indices.device, dtype=torch.int32).unsqueeze(0) < deg.unsqueeze(1)
        nbr[mask] = indices.to(torch.int32)                    # CSC rows are contiguous: row-major mask order matches
        res = (nbr, deg.contiguous(), fan)
    try:
        block._ell_cache = res
    except AttributeError:
        pass
    return res


def _np_pad(out_f):
    """Output width padded to a size every engine kernel accepts as an MMA N (and the dW kernel as an N-side block)."""
    return 64 if out_f <= 64 else (128 if out_f <= 128 else 256)


class _FusedSAGETC(torch.autograd.Function):
    """Forward: ONE sage_layer_tc launch. Backward: sage_bwd_dw (weight gradient, MN-major tcgen05 split-K) and
    sage_bwd_da (input gradient whose epilogue scatters through the ELL block) - no cuBLAS call in either direction."""

    @staticmethod
    def forward(ctx, x, w_self, w_neigh, bias, nbr, cnt, fan, relu, n_dev, self_ids=None):
        # self_ids (int32 [n_dst]) + nbr in the row space of ``x``: ``x`` is then a feature TABLE read in place (no gathered
        # copy of the input frontier exists); without it row d of the output reads x[d] as its own row
        n_dst = nbr.shape[0]
        F = x.shape[1]
        out_f = w_self.shape[0]
        Np = _np_pad(out_f)
        dev, dt = x.device, x.dtype
        # [W_self | W_neigh], K-major: one cat when the width needs no padding (few torch ops: this path is host-bound)
        Wt = torch.cat((w_self, w_neigh), dim=1)
        if Np != out_f:
            Wt = torch.nn.functional.pad(Wt, (0, 0, 0, Np - out_f))
        if bias is None:
            b_pad = torch.zeros(Np, dtype=torch.float32, device=dev)
        else:
            b_pad = bias.float() if Np == out_f else torch.nn.functional.pad(bias.float(), (0, Np - out_f))
        if n_dev is None:
            n_dev = torch.tensor([n_dst], dtype=torch.int64, device=dev)
        planes = fused_ops.tf32_split(Wt) if dt == torch.float32 else (Wt.contiguous(), None)
        out, A = fused_ops.sage_layer_tc(x, self_ids, nbr, cnt, n_dev, planes[0], b_pad, relu=relu, save_A=True,
                                         Wt_lo=planes[1])                        # A = [x_self | mean x_nbr], kept for dW
        ctx.save_for_backward(planes[0], planes[1] if planes[1] is not None else _EMPTY.get(dev), A, nbr, cnt, n_dev,
                              out if relu else _EMPTY.get(dev))
        if self_ids is not None and x.requires_grad:
            raise RuntimeError("a feature table read in place cannot require grad")
        ctx.meta = (x.shape[0], F, out_f, Np, fan, relu, bias is not None)
        return out if Np == out_f else out[:, :out_f]

    @staticmethod
    def backward(ctx, dout):
        Wt, Wt_lo, A, nbr, cnt, n_dev, out = ctx.saved_tensors
        n_src, F, out_f, Np, fan, relu, has_bias = ctx.meta
        dt = A.dtype
        dZ = dout.to(dt)
        if Np != out_f:
            dZ = torch.nn.functional.pad(dZ, (0, Np - out_f))
        if relu:
            dZ = dZ * (out > 0)
        dZ = dZ.contiguous()
        # dWt [Np, 2F] = dZ^T A. The M side of the MMA must be a multiple of 128: A's columns (2F, F % 64 == 0) always are.
        dWt = torch.zeros((Np, 2 * F), dtype=torch.float32, device=dout.device)
        fused_ops.sage_bwd_dw(dWt, dZ, n_dev, P=A, transposed_out=True)
        dX = None
        if ctx.needs_input_grad[0]:
            dX = torch.zeros((n_src, F), dtype=torch.float32, device=dout.device)
            fused_ops.sage_bwd_da(dZ, Wt, nbr, cnt, n_dev, dX, Wt_lo=Wt_lo if Wt_lo.numel() else None)
            dX = dX.to(dt)
        d_bias = dZ.float().sum(0)[:out_f] if has_bias else None
        return (dX, dWt[:out_f, :F].to(dt), dWt[:out_f, F:].to(dt), d_bias, None, None, None, None, None, None)


class _EmptyCache:
    """One shared zero-length tensor per device (placeholders in save_for_backward)."""

    def __init__(self):
        self._t = {}

    def get(self, dev):
        t = self._t.get(dev)
        if t is None:
            t = self._t[dev] = torch.empty(0, device=dev)
        return t


_EMPTY = _EmptyCache()


def usable(block, feat, out_feats, edge_weight):
    if not ENABLED[0] or edge_weight is not None or not torch.is_tensor(feat) or not feat.is_cuda or feat.dim() != 2:
        return False
    if not getattr(block, "is_block", False) or len(block.canonical_etypes) != 1 or out_feats > 256:
        return False
    dt = feat.dtype
    if torch.is_autocast_enabled():
        dt = torch.get_autocast_dtype("cuda") if hasattr(torch, "get_autocast_dtype") else torch.get_autocast_gpu_dtype()
    return dt in (torch.bfloat16, torch.float32) and feat.shape[1] <= 256 and block.num_src_nodes() >= block.num_dst_nodes() > 0


def fused_sage_mean(block, x, w_self, w_neigh, bias, relu=False):
    """SAGE-mean layer on a block; returns None when the block does not fit the kernel (caller falls back)."""
    nbr, cnt, fan = block_to_ell(block)
    if nbr is None:
        return None
    F = x.shape[1]
    Fp = _pad_to(F, 64)
    dt = torch.bfloat16 if (x.dtype == torch.bfloat16 or torch.is_autocast_enabled()) else torch.float32
    self_ids = None
    table = getattr(block, "_global_input", None)          # (self ids, neighbour ids) in the row space of a feature table
    if table is not None:
        if Fp != F or x.dtype != dt or x.requires_grad:
            return None                                     # padding / casting a whole table per step would defeat the point
        self_ids, nbr = table
    x = x.to(dt)
    ws, wn = w_self.to(dt), w_neigh.to(dt)
    if Fp != F:                                                   # zero columns: mathematically identical
        x = torch.nn.functional.pad(x, (0, Fp - F))
        ws = torch.nn.functional.pad(ws, (0, Fp - F))
        wn = torch.nn.functional.pad(wn, (0, Fp - F))
    with torch.autocast("cuda", enabled=False):
        return _FusedSAGETC.apply(x.contiguous(), ws, wn, bias, nbr, cnt, fan, relu, getattr(block, "_n_dst_dev", None),
                                  self_ids)
