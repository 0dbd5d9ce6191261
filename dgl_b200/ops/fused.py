"""Python entry points of the fused tensor-core kernels (csrc/cuda/sage_tc.cu)."""
from __future__ import annotations

import torch

from .._native import loader
from ..base import DGLError, dtype_code


def _p(t):
    return 0 if t is None else t.data_ptr()


def tf32_split(t):
    """x = hi + lo with hi exactly representable in TF32 (round-to-nearest on the 10-bit mantissa) - the operand
    planes of the fp32 tensor-core kernels (csrc/cuda/tc_common.cuh: tf32_split)."""
    t = t.contiguous().float()
    hi = (t.view(torch.int32) + 0x1000).bitwise_and(-8192).view(torch.float32)
    return hi, t - hi


def sage_layer_tc(X, self_ids, nbr, cnt, n_dev, Wt, bias, relu=True, p_drop=0.0, seed=0, step_dev=None, out=None,
                  save_A=False, A_out=None, trace=None, out_f32=False, Wt_lo=None):
    """out[d] = act([X[self(d)] | mean_k X[nbr[d,k]]] @ Wt^T + bias) for d < n_dev (tcgen05 + TMA + TMEM kernel).

    X: [rows, Fp] bf16 (Fp % 64 == 0) or fp32 (Fp % 32 == 0); self_ids: int32 [n_cap] or None (identity);
    nbr: int32 [n_cap, fanout]; cnt: int32 [n_cap]; n_dev: int64 [1] device tensor; Wt: [Np, 2*Fp] in X's dtype
    (K-major); bias: fp32 [Np]. fp32 operands run as 3xTF32 split products: pass the unsplit ``Wt`` (the planes are
    derived here) or the hi plane as ``Wt`` together with ``Wt_lo``.
    Returns (out [n_cap, Np] in X's dtype (fp32 if out_f32), A or None)."""
    if X.dtype not in (torch.bfloat16, torch.float32) or Wt.dtype != X.dtype:
        raise DGLError("sage_layer_tc needs bf16 or fp32 operands of one dtype")
    if X.dtype == torch.float32 and Wt_lo is None:
        Wt, Wt_lo = tf32_split(Wt)
    n_cap, fanout = nbr.shape
    Fp = X.shape[1]
    Np = Wt.shape[0]
    if Wt.shape[1] != 2 * Fp:
        raise DGLError("Wt must be [Np, 2*Fp]")
    if out is None:
        out = torch.empty((n_cap, Np), dtype=torch.float32 if out_f32 else X.dtype, device=X.device)
    if save_A and A_out is None:
        A_out = torch.zeros((n_cap, 2 * Fp), dtype=X.dtype, device=X.device)
    loader.cuda_lib().fn("dglb200_engine_sage_layer_tc", "ippppp" + "l" + "iii" + "ppppp" + "iflpipp")(
        dtype_code(X.dtype), _p(X), _p(self_ids), _p(nbr), _p(cnt), _p(n_dev), n_cap, fanout, Fp, Np, _p(Wt), _p(Wt_lo),
        _p(bias), _p(out), _p(A_out), 1 if relu else 0, float(p_drop), int(seed), _p(step_dev), 1 if out_f32 else 0,
        _p(trace), loader.stream_ptr())
    return out, A_out


def sage_bwd_dw(out, Q, n_dev, P=None, dH=None, Hmask=None, scale=1.0, transposed_out=False, gb=None, dz_out=None,
                dz_out_lo=None, nbs=0):
    """Weight gradient on tcgen05 (csrc/cuda/sage_bwd_tc.cu): ``out += P^T @ Q`` over the first ``n_dev`` rows.

    P [n_cap, CP] (CP % 128 == 0) is given, or built on the fly as ``(Hmask > 0) * dH * scale`` from the incoming fp32
    gradient ``dH`` and the layer output ``Hmask`` (then ``gb`` += column sums and ``dz_out`` (+ ``dz_out_lo`` for fp32)
    receive dZ). Q [n_cap, CQ]. ``out`` fp32 [CP, CQ], or [CQ, CP] when ``transposed_out``."""
    src = Hmask if P is None else P
    n_cap, CP = src.shape
    CQ = Q.shape[1]
    ldi, ldj = (1, CP) if transposed_out else (CQ, 1)
    loader.cuda_lib().fn("dglb200_engine_sage_bwd_dw", "ipppiifpiipl" + "pllpppi" + "p")(
        dtype_code(Q.dtype), _p(dH), _p(Hmask), _p(P), src.stride(0), CP, float(scale), _p(Q), Q.stride(0), CQ, _p(n_dev),
        n_cap, _p(out), ldi, ldj, _p(gb), _p(dz_out), _p(dz_out_lo), int(nbs), loader.stream_ptr())
    return out


def sage_bwd_da(dz, Wt, nbr, cnt, n_dev, dH_below, dz_lo=None, Wt_lo=None, cb=0):
    """Input gradient with the scatter in the epilogue: dA = dz [n_cap, Np] @ Wt [Np, 2F];
    ``dH_below[d] += dA[d, :F]``, ``dH_below[nbr[d, k]] += dA[d, F:] / cnt[d]`` (fp32 [*, F], caller zeroes it)."""
    n_cap, fanout = nbr.shape
    Np, F2 = Wt.shape
    if dz.dtype == torch.float32 and dz_lo is None:
        dz, dz_lo = tf32_split(dz)
    if Wt.dtype == torch.float32 and Wt_lo is None:
        Wt, Wt_lo = tf32_split(Wt)
    loader.cuda_lib().fn("dglb200_engine_sage_bwd_da", "ipppppppliiipi" + "p")(
        dtype_code(dz.dtype), _p(dz), _p(dz_lo), _p(Wt), _p(Wt_lo), _p(nbr), _p(cnt), _p(n_dev), n_cap, fanout, Np, F2 // 2,
        _p(dH_below), int(cb), loader.stream_ptr())
    return dH_below


def sample_ell(indptr, indices, seeds, fanout, seed=0, n_dev=None, step_dev=None, return_eids=False, pad_rows=True):
    """Uniform neighbour sampling without replacement into a dense ELL block (no host synchronisation).

    indptr/indices: CSC of the graph (int32 or int64, on the device); seeds: [n_cap] ids (-1 = empty slot).
    Returns (nbr [n_cap, fanout] padded with -1, cnt int32 [n_cap], eids or None, total int64 [1]).
    Reference semantics: sampling/neighbor.py sample_neighbors(replace=False) / rowwise_sampling.cu:43-229."""
    idt = indptr.dtype
    n_cap = seeds.shape[0]
    dev = seeds.device
    nbr = torch.empty((n_cap, fanout), dtype=idt, device=dev)
    eid = torch.empty((n_cap, fanout), dtype=idt, device=dev) if return_eids else None
    cnt = torch.empty(n_cap, dtype=torch.int32, device=dev)
    total = torch.zeros(16, dtype=torch.int64, device=dev)  # kTotalSlots partial sums
    from ..base import id_code
    loader.cuda_lib().fn("dglb200_cuda_sample_ell", "ippiplpilpppppi" + "p")(
        id_code(idt), _p(indptr), _p(indices), 0, _p(seeds.to(idt).contiguous()), n_cap, _p(n_dev), fanout, int(seed),
        _p(step_dev), _p(nbr), _p(eid), _p(cnt), _p(total), 1 if pad_rows else 0, loader.stream_ptr())
    return nbr, cnt, eid, total.sum(0, keepdim=True)


class FrontierCompactor:
    """Workspace + launcher of the cooperative frontier-compaction kernel (csrc/cuda/frontier.cu).

    ``compact(first, n_first_dev, nbr_global)`` returns (nodes [out_cap] = first ++ new ids (order of the new ids is
    unspecified) padded with -1, nbr_local like nbr_global, count int64[1] on the device, overflow int32[1]).
    Same contract as to_block / unique_and_compact (transforms/functional.py to_block) minus the ordering of new ids."""

    def __init__(self, first_cap, fanout, out_cap, device):
        N = first_cap * (1 + fanout)
        hcap = 16
        while hcap < 2 * N + 2:
            hcap <<= 1
        i32 = torch.int32
        self.first_cap, self.fanout, self.out_cap, self.hcap = first_cap, fanout, out_cap, hcap
        self.keys = torch.full((hcap,), -1, dtype=i32, device=device)
        self.vals = torch.full((hcap,), -2, dtype=i32, device=device)
        self.slot = torch.empty(N, dtype=i32, device=device)
        self.counter = torch.zeros(1, dtype=i32, device=device)
        self.overflow = torch.zeros(1, dtype=i32, device=device)

    def compact(self, first, n_first_dev, nbr_global):
        dev = first.device
        nodes = torch.full((self.out_cap,), -1, dtype=torch.int32, device=dev)
        nbr_local = torch.empty_like(nbr_global)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        loader.cuda_lib().fn("dglb200_frontier_compact", "pplpipplpplpppp" + "p")(
            _p(first), _p(n_first_dev), self.first_cap, _p(nbr_global), self.fanout, _p(self.keys), _p(self.vals),
            self.hcap, _p(self.slot), _p(nodes), self.out_cap, _p(nbr_local), _p(count), _p(self.counter),
            _p(self.overflow), loader.stream_ptr())
        return nodes, nbr_local, count, self.overflow


# ------------------------------------------------------------------------------------------- ELL aggregation with autograd
class _EllMeanConcat(torch.autograd.Function):
    """A[d] = [x[self(d)] | mean_{k < cnt[d]} x[nbr[d, k]]] for d < n_dev, zero rows beyond (csrc/cuda/engine.cu
    ``agg_concat_kernel``); backward scatters both halves back with vector ``red.add`` (``scatter_grad_kernel``).
    Everything is indexed by the device-side row count, so the op runs unchanged on capacity-shaped buffers inside a
    captured CUDA graph (dataloading/graphed.py)."""

    @staticmethod
    def forward(ctx, x, nbr, cnt, n_dev):
        n_cap, fanout = nbr.shape
        F = x.shape[1]
        A = torch.empty((n_cap, 2 * F), dtype=x.dtype, device=x.device)
        from ..base import id_code
        loader.cuda_lib().fn("dglb200_engine_agg_concat", "iippppplii" + "pi" + "p")(
            dtype_code(x.dtype), id_code(torch.int32), _p(x), 0, _p(nbr), _p(cnt), _p(n_dev), n_cap, fanout, F, _p(A), 1,
            loader.stream_ptr())
        ctx.save_for_backward(nbr, cnt, n_dev)
        ctx.shape = (x.shape[0], F, x.dtype)
        return A

    @staticmethod
    def backward(ctx, dA):
        nbr, cnt, n_dev = ctx.saved_tensors
        n_src, F, dt = ctx.shape
        from ..base import id_code
        dX = torch.zeros((n_src, F), dtype=torch.float32, device=dA.device)
        dA = dA.contiguous()
        loader.cuda_lib().fn("dglb200_engine_scatter_grad", "iipppplii" + "p" + "p")(
            dtype_code(dA.dtype), id_code(torch.int32), _p(dA), _p(nbr), _p(cnt), _p(n_dev), nbr.shape[0], nbr.shape[1], F,
            _p(dX), loader.stream_ptr())
        return dX.to(dt), None, None, None


def ell_mean_concat(x, nbr, cnt, n_dev):
    """[x_self | mean of the ELL neighbours]; x: [n_src, F] bf16 / fp32 (F % 4 == 0), nbr: int32 [n_dst_cap, fanout] local
    source rows, cnt: int32 [n_dst_cap], n_dev: int64 [1] live destination rows. Differentiable in ``x``."""
    if x.dtype not in (torch.bfloat16, torch.float32) or x.shape[1] % 4:
        raise DGLError("ell_mean_concat needs bf16 / fp32 rows with a width that is a multiple of 4")
    return _EllMeanConcat.apply(x.contiguous(), nbr, cnt, n_dev)
