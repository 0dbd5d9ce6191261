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
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
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
    comm = ("nccl" if args.comm == "nccl" else P2PComm()) if world > 1 else None
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    tr = FusedSAGETrainer(g, feat, labels, CLASSES, hidden=HIDDEN, fanouts=FANOUTS, batch_size=BATCH, lr=1e-3,
                          weight_decay=5e-4, dropout=0.5, dtype=dtype, seed=1234, comm=comm,
                          use_cuda_graph=not args.no_graph, use_tc=(False if args.no_tc else None))
    del feat
    total = args.warmup + args.steps
    batches = seed_batches(train_idx.cpu(), rank, world, 2 * total + 4, BATCH, seed=99)
    it = iter(batches)
    for _ in range(max(args.warmup, 3)):
        tr.train_step(next(it))
    # ---- device-timed region: K steps enqueued back to back (each step still does its H2D seed copy and
    #      D2H loss copy inside the CUDA graph; the host only waits at the end)
    e_before = tr.total_sampled_edges()
    clocks = ClockSampler(dev.index)
    if rank == 0:
        clocks.start()

    def loop_async():
        for _ in range(args.steps):
            tr.step_async(next(it))  # pinned ring slot -> H2D -> CUDA-graph replay; host never waits

    def loop_e2e():
        # public API, pipelined: per step pinned seeds -> H2D -> step -> D2H loss into pinned memory -> host float
        # (the host reads step i's loss while step i+3 is already enqueued)
        losses = tr.fit(next(it) for _ in range(args.steps))
        assert len(losses) == args.steps and all(l == l for l in losses)
        return losses[-1]

    torch.cuda.profiler.start()  # no-op unless run under `ncu --profile-from-start off`
    ms = timed(dev, loop_async)
    torch.cuda.profiler.stop()
    clk = clocks.stop() if rank == 0 else None
    e_after = tr.total_sampled_edges()
    edges = sum_over_ranks(float(e_after - e_before), dev)
    e_before = tr.total_sampled_edges()
    ms_e2e = timed(dev, loop_e2e)
    edges_e2e = sum_over_ranks(float(tr.total_sampled_edges() - e_before), dev)
    loss = tr.last_loss()
    if comm is not None and not isinstance(comm, str):
        comm.check_error()
    if rank == 0:
        steps_per_epoch = -(-TRAIN_NODES // (BATCH * world))
        value = edges / (ms / 1e3)
        out = {
            "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "impl": "dgl_b200",
            "config": {"model": "GraphSAGE-mean 3-layer 100-256-256-47", "graph": "ogbn-products-shaped RMAT "
                       f"(nodes={n}, directed_edges={g.num_edges()})", "global_batch": BATCH * world,
                       "per_gpu_batch": BATCH, "fanouts": FANOUTS, "parallelism": f"dp{world}",
                       "optimizer": "Adam lr=1e-3 wd=5e-4", "dropout": 0.5,
                       "l2_policy": "inputs larger than L2 (feature table + CSC >> 126 MB); no flush",
                       "epoch_time_s": steps_per_epoch * ms / args.steps / 1e3, "steps_per_epoch": steps_per_epoch,
                       "seeds_per_s": BATCH * world * args.steps / (ms / 1e3),
                       "sampled_edges_per_step_per_gpu": edges / args.steps / world,
                       "cuda_graph": not args.no_graph, "final_loss": loss, "scale": args.scale},
            "clocks": clk,
            "e2e": {"value": edges_e2e / (ms_e2e / 1e3), "unit": "edges/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": (BATCH + 2) * 4, "d2h_bytes_per_step": 4,
                    "note": "FusedSAGETrainer.fit(): per step the host writes seeds to pinned memory, H2D copy, fused step, D2H loss to pinned memory, host reads every loss (3 steps behind the enqueue)"},
            "gpu_launches": tr.gpu_launches_per_step * args.steps,
            "gpu_launches_per_step": {"own_kernels": tr.gpu_launches_per_step, "all_nodes": tr.total_launches_per_step},
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

    model = SAGE(FEAT, HIDDEN, CLASSES).to(dev)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev], output_device=dev)
    sampler = NeighborSampler(FANOUTS, prefetch_node_feats=["feat"], prefetch_labels=["label"])
    loader = DataLoader(g, train_idx.to(dev), sampler, device=dev, batch_size=BATCH, shuffle=True, drop_last=True,
                        num_workers=0, use_ddp=world > 1, use_uva=False)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
    state = {"it": iter(loader), "edges": 0, "loss": None, "epoch": 0}

    def next_batch():
        try:
            return next(state["it"])
        except StopIteration:
            state["epoch"] += 1
            if world > 1:
                loader.set_epoch(state["epoch"])
            state["it"] = iter(loader)
            return next(state["it"])

    def step(sync_loss):
        input_nodes, output_nodes, blocks = next_batch()
        x = blocks[0].srcdata["feat"]
        y = blocks[-1].dstdata["label"].to(torch.int64)
        y_hat = model(blocks, x)
        loss = F.cross_entropy(y_hat, y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        state["edges"] += sum(b.num_edges() for b in blocks)
        state["loss"] = loss.item() if sync_loss else loss
        return loss

    model.train()
    # the reference builds graph formats / cuBLAS handles / allocator pools lazily over its first iterations: give it
    # a generous warm-up so the timed region measures its steady state
    ref_warmup = max(args.warmup, 30)
    for _ in range(ref_warmup):
        step(True)
    clocks = ClockSampler(dev.index)
    if rank == 0:
        clocks.start()
    state["edges"] = 0
    ms = timed(dev, lambda: [step(False) for _ in range(args.steps)])
    clk = clocks.stop() if rank == 0 else None
    edges = sum_over_ranks(float(state["edges"]), dev)
    state["edges"] = 0
    ms_e2e = timed(dev, lambda: [step(True) for _ in range(args.steps)])
    edges_e2e = sum_over_ranks(float(state["edges"]), dev)
    loss = float(state["loss"])
    if rank == 0:
        steps_per_epoch = -(-TRAIN_NODES // (BATCH * world))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": edges / (ms / 1e3), "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": ref_warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"model": "GraphSAGE-mean 3-layer 100-256-256-47 (dgl.nn.SAGEConv)",
                       "graph": f"ogbn-products-shaped RMAT (nodes={n}, directed_edges={g.num_edges()})",
                       "global_batch": BATCH * world, "per_gpu_batch": BATCH, "fanouts": FANOUTS,
                       "parallelism": f"ddp{world}", "mode": "puregpu (graph+features on device, dgl DataLoader)",
                       "epoch_time_s": steps_per_epoch * ms / args.steps / 1e3, "final_loss": loss,
                       "sampled_edges_per_step_per_gpu": edges / args.steps / world, "dgl_version": dgl.__version__,
                       "scale": args.scale},
            "clocks": clk,
            "e2e": {"value": edges_e2e / (ms_e2e / 1e3), "unit": "edges/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                    "note": "puregpu mode keeps seed ids on the device; loss.item() every step"},
        }))
    if dist.is_initialized():
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (testing only; 1.0 = named config)")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"], help="gradient all-reduce: fused P2P kernel or NCCL baseline")
    ap.add_argument("--no-graph", action="store_true", help="disable CUDA-graph replay (debug)")
    ap.add_argument("--no-tc", action="store_true", help="use the unfused cuBLAS layer path (A/B debug)")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        print(json.dumps({"impl": args.impl, "unavailable": "no CUDA device"}))
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
