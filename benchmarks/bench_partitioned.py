#!/usr/bin/env python
"""Full-graph GraphSAGE on a vertex-partitioned graph (`bench.py --mode partitioned`).

ours:      the products-shaped RMAT graph is split into N parts by the repo's own edge-cut partitioner
           (csrc/cpu/partition.cc, LDG + balanced label propagation), vertices are relabelled so that every part is a
           contiguous id range, each rank keeps the CSC rows of ITS vertices and one epoch is
             HaloAggregate (one kernel: de-duplicated halo pull over NVLink overlapped with interior aggregation)
             -> tcgen05 GEMM [h | mean] @ [W_self ; W_neigh]  (x 3 layers) -> loss on the local training vertices
             -> backward (halo red.add over NVLink, tcgen05 dA / dW) -> P2P gradient all-reduce -> Adam.
reference: the unmodified reference trains the SAME model on the WHOLE graph on every rank under
           DistributedDataParallel (a full-graph replica: there is no partition-parallel full-graph path in the reference's
           public API; `dgl.nn.SAGEConv` x 3 under torch.autocast(bf16)).
Both arms: bf16 compute / fp32 master weights, identical graph, features, labels and training vertices; one "step" is one
full-graph epoch; the timed region contains forward, loss, backward, gradient reduction and the optimizer step.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRIC = "fullgraph_sage_products_rmat_epoch_time_s"
HIDDEN, LAYERS = 256, 3


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def partition_and_relabel(n, indptr, indices, world, rank, dev, seed=7):
    """part ids from the native partitioner (rank 0, broadcast), then the contiguous relabelling."""
    part = torch.empty(n, dtype=torch.int64, device=dev)
    t0 = time.time()
    if rank == 0:
        if world == 1:
            part.zero_()
        else:
            from dgl_b200._native import loader
            from dgl_b200.base import id_code
            ip, ix = indptr.cpu().long().contiguous(), indices.cpu().long().contiguous()   # symmetric graph: CSC == CSR
            out = torch.empty(n, dtype=torch.int64)
            # balance EDGES (vertex weight = degree): the aggregation dominates the epoch and its cost follows the edges
            vw = (ip[1:] - ip[:-1]).clamp(min=1).contiguous()
            loader.cpu_lib().fn("dglb200_cpu_ldg_partition", "ipplipidlp")(
                id_code(torch.int64), ip.data_ptr(), ix.data_ptr(), n, world, vw.data_ptr(), 4, 1.03, seed, out.data_ptr())
            part.copy_(out)
    if world > 1:
        dist.broadcast(part, 0)
    t_part = time.time() - t0
    sizes = torch.bincount(part, minlength=world)
    shard_rows = (int(sizes.max()) + 7) // 8 * 8
    order = torch.argsort(part, stable=True)                       # old ids grouped by part
    starts = torch.cumsum(sizes, 0) - sizes
    new_id = torch.empty(n, dtype=torch.int64, device=dev)
    ps = part[order]
    new_id[order] = ps * shard_rows + (torch.arange(n, device=dev) - starts[ps])
    mine = order[int(starts[rank]): int(starts[rank] + sizes[rank])]
    return part, new_id, mine, shard_rows, t_part


def local_csc(indptr, indices, mine, new_id, shard_rows):
    dev = indptr.device
    deg = (indptr[mine + 1] - indptr[mine]).long()
    lp = torch.zeros(shard_rows + 1, dtype=torch.int64, device=dev)
    lp[1: mine.numel() + 1] = torch.cumsum(deg, 0)
    lp[mine.numel() + 1:] = lp[mine.numel()]
    total = int(lp[-1])
    owner = torch.repeat_interleave(torch.arange(mine.numel(), device=dev), deg, output_size=total)
    pos = indptr[mine].long()[owner] + (torch.arange(total, device=dev) - lp[:-1][owner])
    return lp.to(torch.int32), new_id[indices[pos].long()].to(torch.int32)


def run_ours(args):
    B = _bench()
    rank, world, dev = B.dist_setup(args.gpus)
    import dgl_b200  # noqa: F401
    from dgl_b200.ops.cuda_struct import cuda_coo_to_csr
    from dgl_b200.parallel.p2p import P2PComm
    from dgl_b200.parallel.partitioned import PartitionedSAGETC
    n, u, v, feat, labels, train_idx = B.make_data(dev, args.scale)
    indptr, indices, _ = cuda_coo_to_csr(v, u, n)                   # CSC: rows = destinations
    del u, v
    part, new_id, mine, shard_rows, t_part = partition_and_relabel(n, indptr, indices, world, rank, dev)
    lp, lidx = local_csc(indptr, indices, mine, new_id, shard_rows)
    # out-edges of the owned vertices for the pull-based backward; the RMAT graph is symmetrised, so CSR == CSC
    reverse = (lp, lidx)
    lo = rank * shard_rows
    remote = ((lidx < lo) | (lidx >= lo + shard_rows))
    cut_local = float(remote.float().mean()) if lidx.numel() else 0.0
    is_train = torch.zeros(n, dtype=torch.bool, device=dev)
    is_train[train_idx] = True
    feat_w = (B.FEAT + 7) // 8 * 8          # 100 -> 104 zero-padded columns: 16-byte rows for the halo copies and the GEMM
    x = torch.zeros(shard_rows, feat_w, device=dev, dtype=torch.bfloat16)
    x[: mine.numel(), : B.FEAT] = feat[mine].to(torch.bfloat16)
    y = torch.zeros(shard_rows, dtype=torch.int64, device=dev)
    y[: mine.numel()] = labels[mine]
    m = torch.zeros(shard_rows, dtype=torch.bool, device=dev)
    m[: mine.numel()] = is_train[mine]
    n_train_global = int(train_idx.numel())
    del indptr, indices, feat, labels, new_id, part
    torch.cuda.empty_cache()
    comm = P2PComm(dist.group.WORLD if world > 1 else None, dev)
    torch.manual_seed(1234)
    model = PartitionedSAGETC(comm, lp, lidx, shard_rows, feat_w, HIDDEN, B.CLASSES, num_layers=LAYERS, dropout=0.5,
                              reverse=reverse).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3, fused=True)
    train_rows = m.nonzero().squeeze(1)

    def step():
        logits = model(x)
        loss = F.cross_entropy(logits[train_rows].float(), y[train_rows], reduction="sum") * (world / n_train_global)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            model.allreduce_gradients(average=True)                 # mean of (world * local share) = global mean loss grad
        opt.step()
        return loss

    clocks = B.ClockSampler(dev.index or 0)
    for _ in range(args.warmup):
        loss = step()
    B.barrier_sync(dev)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    B.barrier_sync(dev)
    ms = B.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    ck = clocks.stop()
    loss_g = B.sum_over_ranks(float(loss.detach()) / world, dev)
    plan = model.agg.plan
    halo_rows = B.sum_over_ranks(plan.n_halo, dev)
    cut = B.sum_over_ranks(cut_local * lidx.numel(), dev) / max(B.sum_over_ranks(lidx.numel(), dev), 1)
    peak = B.max_over_ranks(torch.cuda.max_memory_allocated(dev) / 2 ** 30, dev)
    edges_max = B.max_over_ranks(lidx.numel(), dev)               # collectives: every rank, outside the rank-0 branch
    edges_mean = B.sum_over_ranks(lidx.numel(), dev) / world
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": ms / 1e3, "unit": "s/epoch", "impl": "ours", "mode": "partitioned", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
            "dtype": "bf16", "data": "synthetic", "final_loss": loss_g, "clocks": ck,
            "config": {"model": f"GraphSAGE-mean {LAYERS}x{HIDDEN} full-graph", "nodes": n, "edges": int(B.NUM_EDGES * args.scale) * 2,
                       "parallelism": f"vertex-partition x{world} (LDG edge-cut partitioner, P2P halo)",
                       "edge_cut_fraction": cut, "halo_rows_total": halo_rows, "shard_rows": shard_rows,
                       "partition_time_s": t_part, "peak_mem_gib": peak, "halo_mode": os.environ.get("DGLB200_HALO_MODE", "balanced"),
                       "local_edges_max": edges_max, "local_edges_mean": edges_mean,
                       "halo_bytes_per_layer_fwd": halo_rows * HIDDEN * 2}}), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


def run_reference(args):
    B = _bench()
    rank, world, dev = B.dist_setup(args.gpus)
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    os.environ.setdefault("DGLBACKEND", "pytorch")
    try:
        import dgl
        import dgl.nn as dglnn
    except Exception as ex:  # noqa: BLE001
        if rank == 0:
            print(json.dumps({"impl": "reference", "mode": "partitioned", "unavailable": f"{type(ex).__name__}: {ex}"[:200]}))
        return
    n, u, v, feat, labels, train_idx = B.make_data(dev, args.scale)
    g = dgl.graph((u, v), num_nodes=n, idtype=torch.int32, device=dev)
    del u, v
    g.create_formats_()

    class SAGE(torch.nn.Module):
        def __init__(self):
            super().__init__()
            sizes = [B.FEAT] + [HIDDEN] * (LAYERS - 1) + [B.CLASSES]
            self.layers = torch.nn.ModuleList(dglnn.SAGEConv(sizes[i], sizes[i + 1], "mean") for i in range(LAYERS))
            self.dropout = torch.nn.Dropout(0.5)

        def forward(self, g, h):
            for i, layer in enumerate(self.layers):
                h = layer(g, h)
                if i < LAYERS - 1:
                    h = self.dropout(F.relu(h))
            return h

    torch.manual_seed(1234)
    model = SAGE().to(dev)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], output_device=dev.index)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(g, feat)
        loss = F.cross_entropy(logits[train_idx].float(), labels[train_idx])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    clocks = B.ClockSampler(dev.index or 0)
    for _ in range(args.warmup):
        loss = step()
    B.barrier_sync(dev)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    B.barrier_sync(dev)
    ms = B.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    ck = clocks.stop()
    peak = B.max_over_ranks(torch.cuda.max_memory_allocated(dev) / 2 ** 30, dev)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": ms / 1e3, "unit": "s/epoch", "impl": "reference", "mode": "partitioned", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
            "dtype": "bf16", "data": "synthetic", "final_loss": float(loss), "clocks": ck,
            "config": {"model": f"GraphSAGE-mean {LAYERS}x{HIDDEN} full-graph", "nodes": n,
                       "parallelism": f"DistributedDataParallel full-graph replica x{world}", "peak_mem_gib": peak}}), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


def run(args):
    (run_reference if args.impl == "reference" else run_ours)(args)
