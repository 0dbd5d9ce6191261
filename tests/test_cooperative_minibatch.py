"""Cooperative minibatching: id routing, differentiable row exchange and the cooperative NeighborSampler, checked against
a single-process dense oracle on 2 gloo ranks (reference behaviour: python/dgl/graphbolt/subgraph_sampler.py:141-310,
impl/cooperative_conv.py:27-93)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dgl_b200.graphbolt as gb
from dgl_b200.graphbolt.impl.cooperative import ExchangePlan, cooperative_sample, rank_assignment

N, D = 60, 5


def make_graph():
    gen = torch.Generator().manual_seed(0)
    e = 400
    dst = torch.randint(0, N, (e,), generator=gen).sort().values
    src = torch.randint(0, N, (e,), generator=gen)
    indptr = torch.zeros(N + 1, dtype=torch.int64)
    indptr[1:] = torch.bincount(dst, minlength=N).cumsum(0)
    x = torch.randn(N, D, generator=gen)
    w = [torch.randn(D, D, generator=gen) * 0.3 for _ in range(2)]
    return indptr, src, dst, x, w


def dense_oracle(x, w, src, dst):
    a = torch.zeros(N, N)
    a.index_put_((dst, src), torch.ones(src.numel()), accumulate=True)
    deg = a.sum(1, keepdim=True).clamp(min=1)
    h = x
    for wl in w:
        h = torch.tanh(((a @ h) / deg + h) @ wl)
    return h


def layer_apply(csc, h_src, n_dst, wl):
    ptr, idx = csc.indptr.long(), csc.indices.long()
    deg = ptr[1:] - ptr[:-1]
    owner = torch.repeat_interleave(torch.arange(n_dst), deg)
    agg = torch.zeros(n_dst, h_src.shape[1]).index_add_(0, owner, h_src[idx])
    return torch.tanh((agg / deg.clamp(min=1).unsqueeze(1) + h_src[:n_dst]) @ wl)


def test_single_process_plan_is_identity_routing():
    ids = torch.tensor([5, 3, 5, 9])
    plan = ExchangePlan.route(ids)
    assert plan.owned.tolist() == [3, 5, 9]
    h = torch.arange(3.0).unsqueeze(1).requires_grad_()
    rows = plan.pull(h)
    assert rows.squeeze(1).tolist() == [1.0, 0.0, 1.0, 2.0]
    rows.sum().backward()
    assert h.grad.squeeze(1).tolist() == [1.0, 2.0, 1.0]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        indptr, src, dst, x, w = make_graph()
        g = gb.fused_csc_sampling_graph(indptr, src)
        seeds = torch.tensor([1, 7, 20, 33, 7]) if rank == 0 else torch.tensor([7, 40, 41, 59])
        items = gb.ItemSet(seeds, names="seeds")
        pipe = gb.ItemSampler(items, batch_size=16)
        pipe = gb.NeighborSampler(pipe, g, [torch.tensor([-1]), torch.tensor([-1])], cooperative=True)
        mb = next(iter(pipe))
        # every id this rank works on is one it owns
        for ids in [mb.input_nodes] + [sg.original_column_node_ids for sg in mb.sampled_subgraphs]:
            assert (rank_assignment(ids, world) == rank).all()
        feats = x[mb.input_nodes].clone().requires_grad_()
        weights = [wl.clone().requires_grad_() for wl in w]
        conv = gb.impl.cooperative.CooperativeConv() if hasattr(gb, "impl") else None
        h = feats
        for sg, wl in zip(mb.sampled_subgraphs, weights):
            h_src = sg._plan.pull(h)
            h = layer_apply(sg.sampled_csc, h_src, sg.original_column_node_ids.shape[0], wl)
        out = mb._out_plan.pull(h)
        want = dense_oracle(x, w, src, dst)
        assert torch.allclose(out, want[seeds], atol=1e-5), (out - want[seeds]).abs().max()
        # gradients: d(sum of all ranks' outputs) / d weights, summed over ranks, equals the oracle's
        out.sum().backward()
        grads = [wl.grad.clone() for wl in weights]
        for gr in grads:
            dist.all_reduce(gr)
        # feature gradients live on the owners; collect them into a dense [N, D] table
        gx = torch.zeros(N, D)
        gx[mb.input_nodes] = feats.grad
        dist.all_reduce(gx)
        torch.save(dict(grads=grads, gx=gx, seeds=seeds, n_inputs=mb.input_nodes.numel()), os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_cooperative_forward_backward(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt") for r in range(world))
    indptr, src, dst, x, w = make_graph()
    xo = x.clone().requires_grad_()
    wo = [wl.clone().requires_grad_() for wl in w]
    all_seeds = torch.cat([r0["seeds"], r1["seeds"]])
    dense_oracle(xo, wo, src, dst)[all_seeds].sum().backward()
    for got, ref in zip(r0["grads"], wo):
        assert torch.allclose(got, ref.grad, atol=1e-4)
    assert torch.allclose(r0["gx"], xo.grad, atol=1e-4) and torch.allclose(r1["gx"], xo.grad, atol=1e-4)
    # no vertex's input row is loaded twice across the ranks
    assert r0["n_inputs"] + r1["n_inputs"] <= N
