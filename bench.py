#!/usr/bin/env python
"""Headline benchmark: 3-layer GraphSAGE (mean) mini-batch training on an ogbn-products-shaped RMAT graph.

Config (reference: /root/reference/examples/multigpu/node_classification_sage.py, mode=puregpu):
  2,449,029 nodes, 61,859,140 RMAT edges symmetrised (123.7 M directed, as DGL stores ogbn-products),
  feat 100, hidden 256, 47 classes, fan-out [10,10,10], batch 1024 seeds PER GPU, Adam(1e-3, wd 5e-4),
  dropout 0.5, data parallel (one process per GPU). Synthetic graph/features/labels, random-init weights.

  python bench.py --gpus N --steps K --warmup W                 # this framework (fused engine)
  python bench.py --impl reference --gpus N --steps K --warmup W  # unmodified reference (baseline/_ref)
For N > 1 launch with torchrun (one rank per GPU). Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
NUM_NODES = 2449029
NUM_EDGES = 61859140
FEAT = 100
HIDDEN = 256
CLASSES = 47
FANOUTS = [10, 10, 10]
BATCH = 1024
TRAIN_NODES = 196615
METRIC = "graphsage_products_rmat_train_sampled_edges_per_sec"


def rmat(num_nodes, num_edges, device, seed=0, a=0.57, b=0.19, c=0.19, chunk=1 << 24):
    """Plain-torch R-MAT generator shared by both arms (identical graph for identical seed)."""
    scale = max(1, (num_nodes - 1).bit_length())
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    srcs, dsts = [], []
    done = 0
    while done < num_edges:
        m = min(chunk, num_edges - done)
        src = torch.zeros(m, dtype=torch.int64, device=device)
        dst = torch.zeros(m, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(m, device=device, generator=gen)
            src = (src << 1) | (r >= a + b).long()
            dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c)).long()
        src = ((src * 0x9E3779B1 + 0x7F4A7C15) % (1 << scale)) % num_nodes
        dst = ((dst * 0x9E3779B1 + 0x7F4A7C15) % (1 << scale)) % num_nodes
        srcs.append(src.to(torch.int32))
        dsts.append(dst.to(torch.int32))
        done += m
    u, v = torch.cat(srcs), torch.cat(dsts)
    return torch.cat([u, v]), torch.cat([v, u])  # symmetrise


def make_data(device, scale, seed=0):
    n = max(1000, int(NUM_NODES * scale))
    e = max(10000, int(NUM_EDGES * scale))
    u, v = rmat(n, e, device, seed)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1)
    feat = torch.randn(n, FEAT, device=device, generator=gen)
    labels = torch.randint(0, CLASSES, (n,), device=device, generator=gen)
    ntrain = max(BATCH, int(TRAIN_NODES * scale))
    train_idx = torch.randperm(n, device=device, generator=gen)[:ntrain]
    return n, u, v, feat, labels, train_idx


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [x.strip() for x in r.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != n_gpus and world > 1:
        raise SystemExit(f"--gpus {n_gpus} but WORLD_SIZE={world}")
    if n_gpus > 1 and world == 1:
        raise SystemExit("for --gpus > 1 launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                         "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a mismatched collective must fail fast, not sit in the default 10-minute watchdog on a multi-GPU box
        import datetime
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev,
                                timeout=datetime.timedelta(seconds=int(os.environ.get("BENCH_NCCL_TIMEOUT_S", "600"))))
    return rank, world, dev


def max_over_ranks(x, dev):
    if dist.is_initialized():
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x


def sum_over_ranks(x, dev):
    if dist.is_initialized():
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    return x


def barrier_sync(dev):
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize(dev)


def timed(dev, fn_loop):
    """barrier+sync, CUDA-event timed region, barrier+sync; returns ms (max over ranks)."""
    barrier_sync(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn_loop()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    barrier_sync(dev)
    return max_over_ranks(ms, dev)


def seed_batches(train_idx_cpu, rank, world, nsteps, batch, seed):
    """Per-rank shard of a seeded permutation, cycled; CPU int32 tensors."""
    g = torch.Generator()
    g.manual_seed(seed)
    out = []
    perm = train_idx_cpu[torch.randperm(train_idx_cpu.numel(), generator=g)]
    shard = perm[rank::world]
    pos = 0
    for _ in range(nsteps):
        if pos + batch > shard.numel():
            perm = train_idx_cpu[torch.randperm(train_idx_cpu.numel(), generator=g)]
            shard = perm[rank::world]
            pos = 0
        out.append(shard[pos:pos + batch].to(torch.int32).contiguous())
        pos += batch
    return out


# ----------------------------------------------------------------------------------------- shared protocol
DTYPES = {"fp32": ["fp32"], "bf16": ["bf16"], "both": ["fp32", "bf16"]}


def measure(dev, args, step_async_loop, e2e_loop, edge_counter):
    """The timing protocol BOTH arms run (identical warm-up, identical regions):

    region A1 (K steps, device-timed, host does not read results) -> region E (K steps through the public training
    loop, host reads every loss) -> region A2 (K steps again). ``value`` uses min(A1, A2): the second region is the
    steady state of whichever arm still had lazy initialisation left after the W warm-up steps."""
    regions = []
    for name, fn in (("A1", step_async_loop), ("E", e2e_loop), ("A2", step_async_loop)):
        e0 = edge_counter()
        ms = timed(dev, fn)
        edges = sum_over_ranks(float(edge_counter() - e0), dev)
        regions.append((name, ms, edges))
    a = min((r for r in regions if r[0] != "E"), key=lambda r: r[1])
    e = regions[1]
    return {"ms": a[1], "edges": a[2], "ms_e2e": e[1], "edges_e2e": e[2],
            "regions_ms": {r[0]: round(r[1], 4) for r in regions}}


# ----------------------------------------------------------------------------------------- ours
def run_ours(args):
    sys.path.insert(0, ROOT)
    rank, world, dev = dist_setup(args.gpus)
    import dgl_b200 as dgl
    from dgl_b200.models import FusedSAGETrainer
    from dgl_b200.parallel import P2PComm
    n, u, v, feat, labels, train_idx = make_data(dev, args.scale)
    g = dgl.graph((u, v), num_nodes=n, idtype=torch.int32, device=dev)
    del u, v
    # --comm nccl: BASELINE communication path (NCCL all-reduce + the same Adam kernel) for A/B against the fused P2P path
    W = max(args.warmup, 3)
    train_cpu = train_idx.cpu()
    steps_per_epoch = -(-TRAIN_NODES // (BATCH * world))
    results = {}
    clk = None
    for di, dname in enumerate(DTYPES[args.dtype]):
        comm = ("nccl" if args.comm == "nccl" else P2PComm()) if world > 1 else None
        dtype = torch.bfloat16 if dname == "bf16" else torch.float32
        tr = FusedSAGETrainer(g, feat, labels, CLASSES, hidden=HIDDEN, fanouts=FANOUTS, batch_size=BATCH, lr=1e-3,
                              weight_decay=5e-4, dropout=0.5, dtype=dtype, seed=1234, comm=comm,
                              use_cuda_graph=not args.no_graph, use_tc=(False if args.no_tc else None))
        batches = seed_batches(train_cpu, rank, world, W + 3 * args.steps + steps_per_epoch + 8, BATCH, seed=99)
        it = iter(batches)
        for _ in range(W):
            tr.train_step(next(it))

        def loop_async():
            for _ in range(args.steps):
                tr.step_async(next(it))  # pinned ring slot -> H2D -> CUDA-graph replay; host never waits

        def loop_e2e():
            # public API, pipelined: per step pinned seeds -> H2D -> step -> D2H loss into pinned memory -> host float
            losses = tr.fit(next(it) for _ in range(args.steps))
            assert len(losses) == args.steps and all(l == l for l in losses)

        clocks = ClockSampler(dev.index)
        if rank == 0 and di == 0:
            clocks.start()
        torch.cuda.profiler.start()  # no-op unless run under `ncu --profile-from-start off`
        m = measure(dev, args, loop_async, loop_e2e, tr.total_sampled_edges)
        torch.cuda.profiler.stop()
        if rank == 0 and di == 0:
            clk = clocks.stop()
        # one full epoch through the public loop (measured, not extrapolated)
        ms_epoch = timed(dev, lambda: tr.fit(next(it) for _ in range(steps_per_epoch))) if args.epoch else None
        loss = tr.last_loss()
        overflow = bool(int(tr.overflow.item()) != 0) or tr.overflow_steps > 0
        overflow = bool(max_over_ranks(float(overflow), dev))
        if comm is not None and not isinstance(comm, str):
            comm.check_error()
        results[dname] = dict(m, loss=loss, overflow=overflow, ms_epoch=ms_epoch, use_tc=bool(tr.use_tc),
                              own=tr.gpu_launches_per_step, nodes=tr.total_launches_per_step,
                              precision=tr.precision_note, caps=tr.caps[1:tr.L], calibrated=getattr(tr, "calibrated", None))
        del tr, comm
        torch.cuda.empty_cache()
    if rank == 0:
        def pack(d, r):
            return {"dtype": d, "value": r["edges"] / (r["ms"] / 1e3), "ms_per_step": r["ms"] / args.steps,
                    "e2e_value": r["edges_e2e"] / (r["ms_e2e"] / 1e3), "e2e_ms_per_step": r["ms_e2e"] / args.steps,
                    "regions_ms": r["regions_ms"], "final_loss": r["loss"], "overflow": r["overflow"],
                    "tcgen05_path": r["use_tc"], "precision": r["precision"],
                    "epoch_time_s_measured": None if r["ms_epoch"] is None else r["ms_epoch"] / 1e3}
        first = DTYPES[args.dtype][0]
        r = results[first]
        out = {
            "metric": METRIC, "value": r["edges"] / (r["ms"] / 1e3), "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": r["ms"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": first, "data": "synthetic", "impl": "dgl_b200",
            "config": {"model": "GraphSAGE-mean 3-layer 100-256-256-47", "graph": "ogbn-products-shaped RMAT "
                       f"(nodes={n}, directed_edges={g.num_edges()})", "global_batch": BATCH * world,
                       "per_gpu_batch": BATCH, "fanouts": FANOUTS, "parallelism": f"dp{world}",
                       "optimizer": "Adam lr=1e-3 wd=5e-4", "dropout": 0.5,
                       "l2_policy": "inputs larger than L2 (feature table + CSC >> 126 MB); no flush",
                       "timing": "min of two K-step device-timed regions (A1, A2) around the e2e region; same protocol in both arms",
                       "regions_ms": r["regions_ms"], "precision": r["precision"], "tcgen05_path": r["use_tc"],
                       "epoch_time_s": steps_per_epoch * r["ms"] / args.steps / 1e3,
                       "epoch_time_s_measured": None if r["ms_epoch"] is None else r["ms_epoch"] / 1e3,
                       "steps_per_epoch": steps_per_epoch,
                       "seeds_per_s": BATCH * world * args.steps / (r["ms"] / 1e3),
                       "sampled_edges_per_step_per_gpu": r["edges"] / args.steps / world,
                       "frontier_caps": r["caps"], "overflow": r["overflow"],
                       "cuda_graph": not args.no_graph, "final_loss": r["loss"], "scale": args.scale,
                       "grad_allreduce": ("none (1 GPU)" if world == 1 else "nccl" if args.comm == "nccl" else
                                          "NVLS multimem.ld_reduce fused with Adam" if getattr(comm, "nvls", False)
                                          else "P2P peer loads fused with Adam (CUDA IPC)")},
            "clocks": clk,
            "e2e": {"value": r["edges_e2e"] / (r["ms_e2e"] / 1e3), "unit": "edges/s",
                    "ms_per_step": r["ms_e2e"] / args.steps,
                    "h2d_bytes_per_step": (BATCH + 2) * 4, "d2h_bytes_per_step": 8,
                    "note": "FusedSAGETrainer.fit(): per step the host writes seeds to pinned memory, H2D copy, fused step, "
                            "D2H of loss + overflow flag to pinned memory, host reads every loss (3 steps behind the enqueue)"},
            "gpu_launches": r["own"] * args.steps,
            "gpu_launches_per_step": {"own_kernels": r["own"], "all_nodes": r["nodes"]},
            "pairs": {d: pack(d, rr) for d, rr in results.items()},
        }
        print(json.dumps(out))
    if dist.is_initialized():
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------ reference
def run_reference(args):
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "dgl")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/dgl not installed (see DESIGN.md)"}))
        return
    sys.path.insert(0, ref)
    os.environ.setdefault("DGLBACKEND", "pytorch")
    try:
        import dgl  # the UNMODIFIED reference
        import dgl.nn as dglnn
        from dgl.dataloading import DataLoader, NeighborSampler
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": f"import failed: {type(e).__name__}: {e}"[:300]}))
        return
    import torch.nn as nn
    import torch.nn.functional as F
    rank, world, dev = dist_setup(args.gpus)
    n, u, v, feat, labels, train_idx = make_data(dev, args.scale)
    g = dgl.graph((u.long(), v.long()), num_nodes=n, device=dev)  # idtype int64 as in the reference example
    del u, v
    g.ndata["feat"] = feat
    g.ndata["label"] = labels

    class SAGE(nn.Module):  # examples/multigpu/node_classification_sage.py:61-80
        def __init__(self, in_size, hid_size, out_size):
            super().__init__()
            self.layers = nn.ModuleList([dglnn.SAGEConv(in_size, hid_size, "mean"),
                                         dglnn.SAGEConv(hid_size, hid_size, "mean"),
                                         dglnn.SAGEConv(hid_size, out_size, "mean")])
            self.dropout = nn.Dropout(0.5)

        def forward(self, blocks, x):
            h = x
            for l, (layer, block) in enumerate(zip(self.layers, blocks)):
                h = layer(block, h)
                if l != len(self.layers) - 1:
                    h = F.relu(h)
                    h = self.dropout(h)
            return h

    W = max(args.warmup, 3)
    steps_per_epoch = -(-TRAIN_NODES // (BATCH * world))
    results = {}
    clk = None
    for di, dname in enumerate(DTYPES[args.dtype]):
        torch.manual_seed(1234)
        model = SAGE(FEAT, HIDDEN, CLASSES).to(dev)
        if world > 1:
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev], output_device=dev)
        sampler = NeighborSampler(FANOUTS, prefetch_node_feats=["feat"], prefetch_labels=["label"])
        # the multi-GPU example's loader (examples/multigpu/node_classification_sage.py:197-212): the DDP shard and
        # the per-epoch reshuffle are handled inside DataLoader.__iter__ (no set_epoch on the node DataLoader)
        loader = DataLoader(g, train_idx.to(dev), sampler, device=dev, batch_size=BATCH, shuffle=True, drop_last=True,
                            num_workers=0, use_ddp=world > 1, use_uva=False)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
        state = {"it": iter(loader), "edges": 0, "loss": None}
        amp = dname == "bf16"  # the reference's stock AMP path (python/dgl/backend/pytorch/sparse.py custom_fwd casts)

        def next_batch():
            try:
                return next(state["it"])
            except StopIteration:
                state["it"] = iter(loader)
                return next(state["it"])

        def step(sync_loss):
            input_nodes, output_nodes, blocks = next_batch()
            x = blocks[0].srcdata["feat"]
            y = blocks[-1].dstdata["label"].to(torch.int64)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                y_hat = model(blocks, x)
                loss = F.cross_entropy(y_hat.float(), y)
            opt.zero_grad()
            loss.backward()
            opt.step()
            state["edges"] += sum(b.num_edges() for b in blocks)
            state["loss"] = loss.item() if sync_loss else loss
            return loss

        model.train()
        for _ in range(W):
            step(True)
        clocks = ClockSampler(dev.index)
        if rank == 0 and di == 0:
            clocks.start()
        m = measure(dev, args, lambda: [step(False) for _ in range(args.steps)],
                    lambda: [step(True) for _ in range(args.steps)], lambda: state["edges"])
        if rank == 0 and di == 0:
            clk = clocks.stop()
        ms_epoch = timed(dev, lambda: [step(True) for _ in range(steps_per_epoch)]) if args.epoch else None
        results[dname] = dict(m, loss=float(state["loss"]), ms_epoch=ms_epoch)
        del model, opt, loader, sampler
        torch.cuda.empty_cache()
    if rank == 0:
        def pack(d, r):
            return {"dtype": d, "value": r["edges"] / (r["ms"] / 1e3), "ms_per_step": r["ms"] / args.steps,
                    "e2e_value": r["edges_e2e"] / (r["ms_e2e"] / 1e3), "e2e_ms_per_step": r["ms_e2e"] / args.steps,
                    "regions_ms": r["regions_ms"], "final_loss": r["loss"],
                    "precision": "fp32 (torch default, TF32 off)" if d == "fp32" else "torch.autocast(bfloat16)",
                    "epoch_time_s_measured": None if r["ms_epoch"] is None else r["ms_epoch"] / 1e3}
        first = DTYPES[args.dtype][0]
        r = results[first]
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["edges"] / (r["ms"] / 1e3), "unit": "edges/s",
            "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": r["ms"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": first, "data": "synthetic",
            "config": {"model": "GraphSAGE-mean 3-layer 100-256-256-47 (dgl.nn.SAGEConv)",
                       "graph": f"ogbn-products-shaped RMAT (nodes={n}, directed_edges={g.num_edges()})",
                       "global_batch": BATCH * world, "per_gpu_batch": BATCH, "fanouts": FANOUTS,
                       "parallelism": f"ddp{world}", "optimizer": "Adam lr=1e-3 wd=5e-4", "dropout": 0.5,
                       "mode": "puregpu (graph+features on device, dgl DataLoader)",
                       "timing": "min of two K-step device-timed regions (A1, A2) around the e2e region; same protocol in both arms",
                       "regions_ms": r["regions_ms"],
                       "epoch_time_s": steps_per_epoch * r["ms"] / args.steps / 1e3,
                       "epoch_time_s_measured": None if r["ms_epoch"] is None else r["ms_epoch"] / 1e3,
                       "steps_per_epoch": steps_per_epoch, "final_loss": r["loss"],
                       "sampled_edges_per_step_per_gpu": r["edges"] / args.steps / world, "dgl_version": dgl.__version__,
                       "scale": args.scale},
            "clocks": clk,
            "e2e": {"value": r["edges_e2e"] / (r["ms_e2e"] / 1e3), "unit": "edges/s",
                    "ms_per_step": r["ms_e2e"] / args.steps, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                    "note": "puregpu mode keeps seed ids on the device; loss.item() every step"},
            "pairs": {d: pack(d, rr) for d, rr in results.items()},
        }))
    if dist.is_initialized():
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="both", choices=["both", "fp32", "bf16"],
                    help="both (default): the fp32 pair is `value`, the bf16 pair (reference: torch.autocast) is reported under `pairs`")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (testing only; 1.0 = named config)")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"], help="gradient all-reduce: fused P2P kernel or NCCL baseline")
    ap.add_argument("--no-graph", action="store_true", help="disable CUDA-graph replay (debug)")
    ap.add_argument("--no-tc", action="store_true", help="use the unfused cuBLAS layer path (A/B debug)")
    ap.add_argument("--no-epoch", dest="epoch", action="store_false", help="skip the measured full epoch")
    ap.add_argument("--mode", default="minibatch", choices=["minibatch", "partitioned"],
                    help="minibatch (default): the headline sampled-training step; partitioned: full-graph SAGE on a "
                         "vertex-partitioned graph (benchmarks/bench_partitioned.py), one step = one epoch")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        print(json.dumps({"impl": args.impl, "unavailable": "no CUDA device"}))
        return
    try:
        if args.mode == "partitioned":
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "benchmarks"))
            import bench_partitioned
            if args.steps == 100 and args.warmup == 10:       # the headline defaults are too long for whole-graph epochs
                args.steps, args.warmup = 10, 3
            bench_partitioned.run(args)
        elif args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    except BaseException:  # noqa: BLE001 - every rank prints its own traceback before torchrun reaps it
        import traceback
        sys.stderr.write(f"[bench rank {os.environ.get('RANK', '0')}] FAILED\n{traceback.format_exc()}\n")
        sys.stderr.flush()
        raise


if __name__ == "__main__":
    main()
