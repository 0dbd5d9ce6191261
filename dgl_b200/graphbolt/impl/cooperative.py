"""Cooperative minibatching feature exchange (parity: /root/reference/python/dgl/graphbolt/impl/cooperative_conv.py:27-93,
graphbolt/src/cuda/cooperative_minibatching_utils.cu:37-146). Each rank owns the hidden rows of the vertices hashed
to it; before a layer every rank gathers the rows it needs. ``P2PCooperativeConv`` (below) does the gather as a direct remote
load over NVLink (csrc/cuda/comm.cu p2p_gather) and the backward as a remote red.add scatter - no all_to_all;
``CooperativeConv`` is the torch.distributed all_to_all formulation (the reference behaviour)."""
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
    if ids.is_cuda:
        # radix sort on ceil(log2 world) key bits (csrc/cuda/graph_kernels.cu, reference
        # cooperative_minibatching_utils.cu:56-113 uses a segmented radix sort by rank)
        from ..._graph_kernels import partition_perm
        perm = partition_perm(owner.to(torch.int32).contiguous(), world_size)
    else:
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


class _P2PExchange(torch.autograd.Function):
    """out[i] = h_owner(i)[local(i)] with rows addressed as ``owner * capacity + local``: the forward is a direct remote
    gather out of the owners' symmetric buffers (csrc/cuda/comm.cu ``p2p_gather_kernel``), the backward a remote
    ``red.add`` scatter into the owners' fp32 gradient buffers (``p2p_scatter_add_kernel``) - no all_to_all, no count
    exchange, no host synchronisation."""

    @staticmethod
    def forward(ctx, h, need, ex):
        comm, cap = ex.comm, ex.capacity
        n, F = h.shape
        buf, table = ex._feat(F, h.dtype)
        table[:n].copy_(h)
        comm.barrier(0)
        out = torch.empty((need.shape[0], F), dtype=h.dtype, device=h.device)
        comm.gather(buf, cap, need, out)
        comm.barrier(1)
        ctx.ex, ctx.n, ctx.dtype = ex, n, h.dtype
        ctx.save_for_backward(need)
        return out

    @staticmethod
    def backward(ctx, grad):
        (need,) = ctx.saved_tensors
        ex = ctx.ex
        comm, cap = ex.comm, ex.capacity
        F = grad.shape[1]
        gbuf, gtab = ex._grad(F)
        gtab.zero_()
        comm.barrier(2)
        comm.scatter_add(gbuf, cap, need, grad.float().contiguous())
        comm.barrier(3)
        return gtab[: ctx.n].to(ctx.dtype).clone(), None, None


class P2PCooperativeConv(torch.nn.Module):
    """Cooperative feature exchange over peer memory. Every rank calls ``forward(h_owned, need)`` where ``need`` holds
    the rows it wants as ``owner_rank * capacity + local_row`` (``capacity`` >= the largest number of rows any rank owns,
    identical on all ranks). Drop-in replacement of the all-to-all in ``CooperativeConvFunction`` when all ranks sit on
    one NVLink domain."""

    def __init__(self, comm, capacity):
        super().__init__()
        self.comm, self.capacity = comm, int(capacity)
        self._fb, self._gb = {}, {}

    def _feat(self, F, dtype):
        key = (F, dtype)
        if key not in self._fb:
            es = torch.empty(0, dtype=dtype).element_size()
            b = self.comm.alloc(self.capacity * F * es)
            self._fb[key] = (b, b.view(dtype)[: self.capacity * F].view(self.capacity, F))
        return self._fb[key]

    def _grad(self, F):
        if F not in self._gb:
            b = self.comm.alloc(self.capacity * F * 4)
            self._gb[F] = (b, b.view(torch.float32)[: self.capacity * F].view(self.capacity, F))
        return self._gb[F]

    @staticmethod
    def address(owner, local, capacity):
        return owner.long() * int(capacity) + local.long()

    def forward(self, h, need):
        return _P2PExchange.apply(h.contiguous(), need.contiguous(), self)
