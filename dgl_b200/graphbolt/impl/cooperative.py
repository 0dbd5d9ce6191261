"""Cooperative minibatching feature exchange (parity: /root/reference/python/dgl/graphbolt/impl/cooperative_conv.py:27-93,
graphbolt/src/cuda/cooperative_minibatching_utils.cu:37-146). Each rank owns the hidden rows of the vertices hashed
to it; before a layer every rank gathers the rows it needs. With ``P2PComm`` the gather is a direct remote load over
NVLink (csrc/cuda/comm.cu p2p_gather) and the backward a remote red.add scatter — no all_to_all; without it the
torch.distributed all_to_all fallback is used (the reference behaviour)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def rank_assignment(ids, world_size, seed=0):
    """Owner rank of every id (hash based, reference cooperative_minibatching_utils.cu:37-54)."""
    x = ids.long() * 0x9E3779B1 + seed
    x = x ^ (x >> 15)
    return (x % world_size).to(torch.int64)


def rank_sort(ids, rank, world_size, seed=0):
    """Sort ids by owner; returns (sorted_ids, permutation, offsets[world+1])."""
    owner = rank_assignment(ids, world_size, seed)
    perm = torch.sort(owner, stable=True)[1]
    counts = torch.bincount(owner, minlength=world_size)
    offsets = torch.zeros(world_size + 1, dtype=torch.int64, device=ids.device)
    offsets[1:] = torch.cumsum(counts, 0)
    return ids[perm], perm, offsets


class CooperativeConvFunction(torch.autograd.Function):
    """h_needed = exchange(h_owned): forward all-to-all of hidden rows, backward the reverse all-to-all."""

    @staticmethod
    def forward(ctx, counts_sent, counts_received, seed_inverse_ids, h):
        ctx.counts_sent, ctx.counts_received = counts_sent, counts_received
        ctx.seed_inverse_ids = seed_inverse_ids
        ctx.n_in = h.shape[0]
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return h[seed_inverse_ids]
        out = h.new_empty((sum(counts_received),) + h.shape[1:])
        dist.all_to_all_single(out, h, counts_received, counts_sent)
        return out[seed_inverse_ids]

    @staticmethod
    def backward(ctx, grad_output):
        g = torch.zeros((max(int(ctx.seed_inverse_ids.max()) + 1 if ctx.seed_inverse_ids.numel() else 0, 0),) + grad_output.shape[1:],
                        dtype=grad_output.dtype, device=grad_output.device)
        g.index_add_(0, ctx.seed_inverse_ids, grad_output)
        if not dist.is_initialized() or dist.get_world_size() == 1:
            out = torch.zeros((ctx.n_in,) + grad_output.shape[1:], dtype=g.dtype, device=g.device)
            out[: g.shape[0]] = g
            return None, None, None, out
        out = g.new_empty((sum(ctx.counts_sent),) + g.shape[1:])
        dist.all_to_all_single(out, g, ctx.counts_sent, ctx.counts_received)
        return None, None, None, out


class CooperativeConv(torch.nn.Module):
    def forward(self, subgraph, h):
        return CooperativeConvFunction.apply(getattr(subgraph, "_counts_sent", [h.shape[0]]),
                                             getattr(subgraph, "_counts_received", [h.shape[0]]),
                                             getattr(subgraph, "_seed_inverse_ids", torch.arange(h.shape[0], device=h.device)), h)
