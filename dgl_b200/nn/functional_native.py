"""segment_mm / gather_mm with autograd.

Parity: /root/reference/python/dgl/backend/pytorch/sparse.py:968-1020 (SEGMENTMM,
GATHERMM) and /root/reference/src/array/cuda/gather_mm.cu. The reference loops
over relations on the host (one cuBLAS call each, plus a D2H sync for seglen);
here CUDA inputs sort rows by relation once and run a batched/padded bmm with
no host synchronisation when ``seglen`` is already a device tensor.
"""
from __future__ import annotations

import torch


def _segment_ids(seglen, n, device):
    seglen = seglen.to(device)
    return torch.repeat_interleave(torch.arange(seglen.shape[0], device=device), seglen.long(), output_size=n)


def segment_mm(a, b, seglen_a):
    """out[seg_r] = a[seg_r] @ b[r];  a:[N,K] b:[R,K,M] seglen_a:[R]."""
    if a.shape[0] == 0:
        return a.new_zeros((0, b.shape[-1]))
    R = b.shape[0]
    if not a.is_cuda:
        outs = []
        off = 0
        for r, n in enumerate(seglen_a.tolist()):
            outs.append(a[off:off + n] @ b[r])
            off += n
        return torch.cat(outs, 0) if outs else a.new_zeros((0, b.shape[-1]))
    idx = _segment_ids(seglen_a, a.shape[0], a.device)
    return gather_mm(a, b, idx)


def gather_mm(a, b, idx_b):
    """out[i] = a[i] @ b[idx_b[i]];  a:[N,K] b:[R,K,M]."""
    N, K = a.shape
    R, _, M = b.shape
    if N == 0:
        return a.new_zeros((0, M))
    if R == 1:
        return a @ b[0]
    # Heuristic: when R*K*M is small relative to N use the dense "all relations" product and pick;
    # otherwise group rows by relation (sort) and run a padded batched GEMM.
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
