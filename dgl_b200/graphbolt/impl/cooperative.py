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


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


class ExchangePlan:
    """Routing of a rank's id list to the ranks that own those ids, reusable for any number of row transfers.

    ``route(ids)`` sorts the ids by owner, tells every peer how many ids it will get (one all-to-all of counts), ships the
    ids (one all-to-all of ids) and deduplicates what arrived: ``owned`` are the distinct ids this rank is responsible for
    in this round and ``recv_to_owned`` maps every received request to its row in ``owned``. ``pull(h_owned)`` then returns
    the rows of this rank's ORIGINAL id list (one all-to-all of rows, differentiable: the backward pushes the gradients
    back and accumulates duplicates on the owner). The reference spreads the same protocol over four datapipe stages
    (python/dgl/graphbolt/subgraph_sampler.py:199-310, impl/neighbor_sampler.py cooperative branches)."""

    def __init__(self, sort_perm, send_counts, recv_counts, owned, recv_to_owned, group):
        self.sort_perm, self.send_counts, self.recv_counts = sort_perm, send_counts, recv_counts
        self.owned, self.recv_to_owned, self.group = owned, recv_to_owned, group

    @staticmethod
    def route(ids, group=None, seed=0):
        world, rank = _world(group)
        if world == 1:
            owned, inverse = torch.unique(ids, return_inverse=True)
            n = ids.shape[0]
            return ExchangePlan(torch.arange(n, device=ids.device), [n], [n], owned, inverse, group)
        sorted_ids, perm, offsets = rank_sort(ids, rank, world, seed)
        send = (offsets[1:] - offsets[:-1]).to(torch.int64)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        send_counts, recv_counts = send.tolist(), recv.tolist()
        got = sorted_ids.new_empty(sum(recv_counts))
        dist.all_to_all_single(got, sorted_ids.contiguous(), recv_counts, send_counts, group=group)
        owned, inverse = torch.unique(got, return_inverse=True)
        return ExchangePlan(perm, send_counts, recv_counts, owned, inverse, group)

    @property
    def num_requested(self):
        return self.sort_perm.shape[0]

    def pull(self, h_owned):
        """rows of ``h_owned`` (aligned with ``self.owned``) for this rank's original ids, in their original order."""
        return _PullRows.apply(h_owned, self)


class _PullRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h_owned, plan):
        ctx.plan, ctx.n_owned = plan, h_owned.shape[0]
        staged = h_owned[plan.recv_to_owned]                          # one row per received request, requester-major
        world, _ = _world(plan.group)
        if world == 1:
            arrived = staged
        else:
            arrived = staged.new_empty((sum(plan.send_counts),) + tuple(staged.shape[1:]))
            dist.all_to_all_single(arrived, staged.contiguous(), plan.send_counts, plan.recv_counts, group=plan.group)
        out = torch.empty_like(arrived)
        out[plan.sort_perm] = arrived                                 # undo the owner sort
        return out

    @staticmethod
    def backward(ctx, grad):
        plan = ctx.plan
        outgoing = grad[plan.sort_perm].contiguous()
        world, _ = _world(plan.group)
        if world == 1:
            landed = outgoing
        else:
            landed = outgoing.new_empty((sum(plan.recv_counts),) + tuple(outgoing.shape[1:]))
            dist.all_to_all_single(landed, outgoing, plan.recv_counts, plan.send_counts, group=plan.group)
        g = landed.new_zeros((ctx.n_owned,) + tuple(landed.shape[1:]))
        g.index_add_(0, plan.recv_to_owned, landed)
        return g, None


def cooperative_sample(sample_layer, seeds, num_layers, group=None, seed=0):
    """Cooperative multi-layer sampling. ``sample_layer(owned_ids, hop) -> CSCFormatBase`` samples the in-neighbours
    (global ids) of the ids this rank owns. Returns ``(input_nodes, layers, out_plan)``:

    * ``out_plan``  - routes the caller's seeds to their owners; ``out_plan.pull(h_last)`` fetches the final rows;
    * ``layers[i]`` - ``dict(csc=compacted CSC, dst=owned ids (columns), src=needed ids (rows; dst is its prefix),
      plan=ExchangePlan of src)``; before applying layer i the model calls ``plan.pull(h)`` to obtain the rows of ``src``
      from their owners, the layer then produces rows aligned with ``dst``;
    * ``input_nodes`` - ids whose input features this rank must load (= ``layers[0]['plan'].owned``).

    Every vertex is expanded by exactly one rank per layer, so neighbourhoods shared between ranks are sampled, fetched and
    computed once instead of once per rank."""
    from ..utils import unique_and_compact_csc_formats
    out_plan = ExchangePlan.route(seeds, group, seed)
    owned = out_plan.owned
    layers = []
    for hop in reversed(range(num_layers)):
        csc = sample_layer(owned, hop)
        needed, compacted = unique_and_compact_csc_formats(csc, owned)
        plan = ExchangePlan.route(needed, group, seed)
        layers.insert(0, dict(csc=compacted, dst=owned, src=needed, plan=plan))
        owned = plan.owned
    return owned, layers, out_plan


class CooperativeConvFunction:
    """Functional form kept for API compatibility: ``apply(subgraph, h)`` pulls the rows the subgraph's sources need."""

    @staticmethod
    def apply(subgraph, h):
        return subgraph._plan.pull(h)


class CooperativeConv(torch.nn.Module):
    """``forward(subgraph, h)``: ``h`` holds this rank's owned rows for the layer; returns the rows of the subgraph's source
    nodes, fetched from their owners (reference python/dgl/graphbolt/impl/cooperative_conv.py:27-93)."""

    def forward(self, subgraph, h):
        plan = getattr(subgraph, "_plan", None)
        return h if plan is None else plan.pull(h)


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
