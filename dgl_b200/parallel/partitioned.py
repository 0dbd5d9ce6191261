"""Vertex-partitioned full-graph training on one NVLink domain (the GNN analogue of context parallelism).

Every rank owns a contiguous range of ``shard_rows`` vertices: their features / hidden states, and the CSC rows of the
edges that END in them (source ids are global). One layer =

    barrier -> HaloPlan.spmm (ONE kernel: de-duplicated halo pull over NVLink overlapped with interior aggregation)
            -> local dense transform (weights replicated, gradients all-reduced with the fused all-reduce + Adam)

The backward of the aggregation pushes ``dOut[v] / deg(v)`` to the OWNERS of the source rows with vector ``red.add``
over NVLink (csrc/cuda/comm.cu ``halo_spmm_bwd_kernel``) - no all-to-all, no NCCL call, no host synchronisation.
The reference has no such path: DistDGL pulls halo features through the KVStore per mini-batch
(python/dgl/distributed/graph_services.py), GraphBolt's CooperativeConv all-to-alls them
(python/dgl/graphbolt/impl/cooperative_conv.py:27-93).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .p2p import HaloPlan, P2PComm


class _HaloAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, layer):
        plan, comm = layer.plan, layer.comm
        buf, table = layer._feature_buffer(x.shape[1], x.dtype)
        table[: plan.rows].copy_(x)
        comm.barrier(0)                       # every shard is written
        out = torch.empty((plan.rows, x.shape[1]), dtype=x.dtype, device=x.device)
        layer._aggregate(plan, buf, table, out, layer.mean)
        comm.barrier(1)                       # nobody overwrites its shard while a peer is still pulling it
        ctx.layer = layer
        return out

    @staticmethod
    def backward(ctx, dout):
        layer = ctx.layer
        plan, comm = layer.plan, layer.comm
        Fdim = dout.shape[1]
        if layer.rev_plan is not None:
            # pull formulation: dX[u] = sum over out-edges (u -> v) of dOut[v] / indeg(v). Every rank scales ITS rows of
            # dOut, publishes them and aggregates over the out-neighbours of its own vertices with the same one-kernel
            # halo pipeline as the forward - no atomics, every halo row crosses NVLink once.
            rplan = layer.rev_plan
            buf, table = layer._rev_buffer(Fdim, dout.dtype)
            torch.mul(dout, layer.inv_deg.to(dout.dtype), out=table[: rplan.rows]) if layer.mean else table[: rplan.rows].copy_(dout)
            comm.barrier(2)
            dx = torch.empty((rplan.rows, Fdim), dtype=dout.dtype, device=dout.device)
            layer._aggregate(rplan, buf, table, dx, False)
            comm.barrier(3)
            return dx, None
        dsh, dx = layer._grad_buffer(Fdim)
        dx.zero_()
        comm.barrier(2)                       # all gradient shards are zeroed
        comm.halo_spmm_backward(plan.indptr, layer.indices_global, dsh, plan.rows, Fdim, dout.contiguous(), mean=layer.mean)
        comm.barrier(3)                       # all contributions (local and remote) have landed
        return dx.to(dout.dtype).clone(), None


class HaloAggregate(nn.Module):
    """``out[v] = mean / sum over in-neighbours u of x_global[u]`` for the locally owned ``v``.

    indptr: CSC row pointers of the owned destination rows, indices_global: GLOBAL source ids."""

    def __init__(self, comm: P2PComm, indptr, indices_global, shard_rows, mean=True, reverse=None):
        """``reverse`` (optional): ``(indptr_out, indices_out_global)`` - the CSR rows (out-edges, global destination ids) of
        the owned vertices. With it the backward is a second pull-based halo aggregation over the transposed structure;
        without it gradients are pushed to the owners with ``red.add`` over NVLink."""
        super().__init__()
        self.comm, self.mean = comm, mean
        self.indices_global = indices_global.contiguous()
        self.plan = HaloPlan(comm, indptr, indices_global, shard_rows)
        self._fbufs, self._gbufs, self._rbufs = {}, {}, {}
        self.rev_plan = None
        if reverse is not None:
            self.rev_plan = HaloPlan(comm, reverse[0], reverse[1], shard_rows)
            deg = (indptr[1:] - indptr[:-1]).clamp(min=1).to(torch.float32)
            self.inv_deg = (1.0 / deg).unsqueeze(1)

    @staticmethod
    def _aggregate(plan, buf, table, out, mean):
        """two-phase (bulk pull || balanced SpMM) when the rows are 16-byte multiples, else the fused one-kernel pipeline;
        ``DGLB200_HALO_MODE=fused`` forces the latter."""
        import os
        row_bytes = table.shape[1] * table.element_size()
        if os.environ.get("DGLB200_HALO_MODE", "balanced") != "fused" and row_bytes % 16 == 0 and table.element_size() <= 4:
            plan.spmm_balanced(buf, table, out, mean=mean)
        else:
            plan.spmm(buf, table, out, mean=mean)

    def _rev_buffer(self, feat_dim, dtype):
        key = (feat_dim, dtype)
        if key not in self._rbufs:
            self._rbufs[key] = self.rev_plan.alloc_features(feat_dim, dtype)
        return self._rbufs[key]

    def _feature_buffer(self, feat_dim, dtype):
        key = (feat_dim, dtype)
        if key not in self._fbufs:
            self._fbufs[key] = self.plan.alloc_features(feat_dim, dtype)
        return self._fbufs[key]

    def _grad_buffer(self, feat_dim):
        if feat_dim not in self._gbufs:
            b = self.comm.alloc(self.plan.rows * feat_dim * 4)
            self._gbufs[feat_dim] = (b, b.view(torch.float32)[: self.plan.rows * feat_dim].view(self.plan.rows, feat_dim))
        return self._gbufs[feat_dim]

    def forward(self, x):
        return _HaloAggregate.apply(x.contiguous(), self)


class PartitionedSAGE(nn.Module):
    """GraphSAGE-mean over a vertex-partitioned graph: h' = W_self h + W_neigh mean_{N(v)} h, aggregation through
    ``HaloAggregate``; parameters are replicated, call ``allreduce_gradients`` (or ``comm.fused_allreduce_adam``)
    after backward."""

    def __init__(self, comm, indptr, indices_global, shard_rows, in_size, hid_size, out_size, num_layers=2, dropout=0.0,
                 reverse=None):
        super().__init__()
        self.agg = HaloAggregate(comm, indptr, indices_global, shard_rows, mean=True, reverse=reverse)
        sizes = [in_size] + [hid_size] * (num_layers - 1) + [out_size]
        self.self_lin = nn.ModuleList(nn.Linear(sizes[i], sizes[i + 1]) for i in range(num_layers))
        self.neigh_lin = nn.ModuleList(nn.Linear(sizes[i], sizes[i + 1], bias=False) for i in range(num_layers))
        self.dropout = nn.Dropout(dropout)
        self.comm = comm

    def forward(self, x):
        h = x
        for i, (ls, ln) in enumerate(zip(self.self_lin, self.neigh_lin)):
            h = ls(h) + ln(self.agg(h))
            if i < len(self.self_lin) - 1:
                h = self.dropout(F.relu(h))
        return h

    @torch.no_grad()
    def allreduce_gradients(self, average=True):
        """Combine the (replicated) parameter gradients over the ranks with the one-shot P2P all-reduce (mean, or sum)."""
        ps = [p for p in self.parameters() if p.grad is not None]
        flat = torch.cat([p.grad.reshape(-1).float() for p in ps])
        n = flat.numel()
        pad = (-n) % 4
        if getattr(self, "_gsym", None) is None or self._gsym.nbytes < (n + pad) * 4:
            self._gsym = self.comm.alloc((n + pad) * 4)
        g = self._gsym.view(torch.float32)[: n + pad]
        g.zero_()
        g[:n].copy_(flat)
        out = torch.empty(n + pad, device=flat.device)
        self.comm.all_reduce(self._gsym, out, n + pad, scale=1.0 / self.comm.world_size if average else 1.0)
        off = 0
        for p in ps:
            k = p.numel()
            p.grad.copy_(out[off: off + k].view_as(p.grad))
            off += k


class PartitionedSAGETC(PartitionedSAGE):
    """``PartitionedSAGE`` whose dense half runs on the tensor cores: per layer ONE tcgen05 GEMM
    ``[h | mean_N(h)] @ [W_self ; W_neigh]`` in bf16 with fp32 accumulation (``ops.segment_mm`` with a single segment ->
    csrc/cuda/grouped_mm_tc.cu ``seg_mm_tc_kernel``; its backward is the same kernel for dA and the MN-major split-K
    kernel for dW). The aggregation runs in bf16 as well, which halves the halo bytes crossing NVLink. Master weights stay
    fp32 (the cast is part of the autograd graph)."""

    def __init__(self, comm, indptr, indices_global, shard_rows, in_size, hid_size, out_size, num_layers=2, dropout=0.0,
                 reverse=None):
        super().__init__(comm, indptr, indices_global, shard_rows, in_size, hid_size, out_size, num_layers, dropout, reverse)
        self.out_size = out_size
        self._seglen = torch.tensor([int(shard_rows)], device=comm.device)

    @staticmethod
    def _pad8(w):
        pad = (-w.shape[1]) % 8                                   # GEMM output columns in 16-byte units (47 classes -> 48)
        return (F.pad(w, (0, pad)) if pad else w), pad

    def forward(self, x):
        from ..ops import segment_mm
        h = x.to(torch.bfloat16)
        last = len(self.self_lin) - 1
        for i, (ls, ln) in enumerate(zip(self.self_lin, self.neigh_lin)):
            n_in, n_out = ls.in_features, ls.out_features
            if n_in > n_out:
                # mean() is linear: project first, aggregate the NARROW rows (fewer bytes gathered and fewer halo bytes)
                ws, pad = self._pad8(ls.weight.t())
                wn, _ = self._pad8(ln.weight.t())
                width = ws.shape[1]
                z = segment_mm(h, torch.cat([ws, wn], dim=1).to(torch.bfloat16).unsqueeze(0), self._seglen)
                z = z[:, :width] + self.agg(z[:, width:].contiguous())
                z = z[:, :n_out] if pad else z
            else:
                w, pad = self._pad8(torch.cat([ls.weight, ln.weight], dim=1).t())      # [2 * in, out]
                z = segment_mm(torch.cat([h, self.agg(h)], dim=1), w.to(torch.bfloat16).unsqueeze(0), self._seglen)
                z = z[:, :n_out] if pad else z
            z = z + ls.bias.to(z.dtype)
            h = self.dropout(F.relu(z)) if i < last else z
        return h
