"""tcgen05 fused SAGE layer for the GENERIC API: ``dgl.nn.SAGEConv(mean)`` on a sampled block (message-flow graph).

The block's CSC (every destination has at most ``fanout`` in-edges) is turned into the ELL layout of the engine kernel
and the layer runs as ONE kernel (csrc/cuda/sage_tc.cu): gather + mean + [self | neigh] concat in shared memory ->
tcgen05.mma against [W_self | W_neigh]^T (TMA) -> bias (+ReLU) epilogue. The backward reuses the engine's pieces: two
cuBLAS GEMMs (dW, dA) and the vector-``red.add`` scatter of the mean aggregation. This is the same kernel
``FusedSAGETrainer`` replays inside its CUDA graph, made available to models written against ``dgl.nn``.

Requirements (otherwise SAGEConv keeps its g-SpMM + cuBLAS formulation): CUDA, bf16 activations (e.g. under
``torch.autocast(dtype=torch.bfloat16)``), block whose destination nodes are the first rows of the source nodes,
max in-degree <= 28, out_feats <= 256, no edge weights.
"""
from __future__ import annotations

import os

import torch

from .._native import loader
from ..base import dtype_code, id_code

MAX_FANOUT = 28
ENABLED = [os.environ.get("DGLB200_FUSED_SAGECONV", "1") != "0"]   # switch for A/B measurements


def _pad_to(n, m):
    return (n + m - 1) // m * m


def block_to_ell(block):
    """(nbr int32 [n_dst, fanout], cnt int32 [n_dst], fanout) of a single-relation block, cached on the block."""
    cached = getattr(block, "_ell_cache", None)
    if cached is not None:
        return cached
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


class _FusedSAGETC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_self, w_neigh, bias, nbr, cnt, fan, relu):
        n_dst = nbr.shape[0]
        F = x.shape[1]
        out_f = w_self.shape[0]
        Np = _pad_to(out_f, 32)
        dev = x.device
        Wt = torch.zeros((Np, 2 * F), dtype=torch.bfloat16, device=dev)          # [W_self | W_neigh], K-major
        Wt[:out_f, :F] = w_self
        Wt[:out_f, F:] = w_neigh
        b_pad = torch.zeros(Np, dtype=torch.float32, device=dev)
        if bias is not None:
            b_pad[:out_f] = bias.float()
        out = torch.empty((n_dst, Np), dtype=torch.bfloat16, device=dev)
        A = torch.empty((n_dst, 2 * F), dtype=torch.bfloat16, device=dev)        # [x_self | mean x_nbr], kept for dW
        n_dev = torch.tensor([n_dst], dtype=torch.int64, device=dev)
        loader.cuda_lib().fn("dglb200_engine_sage_layer_tc", "ppppp" + "l" + "iii" + "pppp" + "iflpipp")(
            x.data_ptr(), 0, nbr.data_ptr(), cnt.data_ptr(), n_dev.data_ptr(), n_dst, fan, F, Np, Wt.data_ptr(),
            b_pad.data_ptr(), out.data_ptr(), A.data_ptr(), 1 if relu else 0, 0.0, 0, 0, 0, 0, loader.stream_ptr())
        ctx.save_for_backward(Wt, A, nbr, cnt, n_dev, out if relu else torch.empty(0))
        ctx.meta = (x.shape[0], F, out_f, Np, fan, relu, bias is not None)
        return out[:, :out_f]

    @staticmethod
    def backward(ctx, dout):
        Wt, A, nbr, cnt, n_dev, out = ctx.saved_tensors
        n_src, F, out_f, Np, fan, relu, has_bias = ctx.meta
        n_dst = nbr.shape[0]
        dZ = torch.zeros((n_dst, Np), dtype=torch.bfloat16, device=dout.device)
        dZ[:, :out_f] = dout
        if relu:
            dZ = dZ * (out > 0)
        dWt = torch.mm(dZ.t(), A).float()                                        # [Np, 2F]
        dA = torch.mm(dZ, Wt)                                                    # [n_dst, 2F] bf16
        dX = torch.zeros((n_src, F), dtype=torch.float32, device=dout.device)
        loader.cuda_lib().fn("dglb200_engine_scatter_grad", "iippppliipp")(
            dtype_code(torch.bfloat16), id_code(torch.int32), dA.data_ptr(), nbr.data_ptr(), cnt.data_ptr(),
            n_dev.data_ptr(), n_dst, fan, F, dX.data_ptr(), loader.stream_ptr())
        d_bias = dZ.float().sum(0)[:out_f] if has_bias else None
        return (dX.to(torch.bfloat16), dWt[:out_f, :F].to(torch.bfloat16), dWt[:out_f, F:].to(torch.bfloat16),
                d_bias, None, None, None, None)


def usable(block, feat, out_feats, edge_weight):
    if not ENABLED[0] or edge_weight is not None or not torch.is_tensor(feat) or not feat.is_cuda or feat.dim() != 2:
        return False
    if not getattr(block, "is_block", False) or len(block.canonical_etypes) != 1 or out_feats > 256:
        return False
    dt = feat.dtype
    if torch.is_autocast_enabled():
        dt = torch.get_autocast_dtype("cuda") if hasattr(torch, "get_autocast_dtype") else torch.get_autocast_gpu_dtype()
    return dt == torch.bfloat16 and block.num_src_nodes() >= block.num_dst_nodes() > 0


def fused_sage_mean(block, x, w_self, w_neigh, bias, relu=False):
    """SAGE-mean layer on a block; returns None when the block does not fit the kernel (caller falls back)."""
    nbr, cnt, fan = block_to_ell(block)
    if nbr is None:
        return None
    F = x.shape[1]
    Fp = _pad_to(F, 64)
    x = x.to(torch.bfloat16)
    ws, wn = w_self.to(torch.bfloat16), w_neigh.to(torch.bfloat16)
    if Fp != F:                                                   # zero columns: mathematically identical
        x = torch.nn.functional.pad(x, (0, Fp - F))
        ws = torch.nn.functional.pad(ws, (0, Fp - F))
        wn = torch.nn.functional.pad(wn, (0, Fp - F))
    with torch.autocast("cuda", enabled=False):
        return _FusedSAGETC.apply(x.contiguous(), ws, wn, bias, nbr, cnt, fan, relu)
