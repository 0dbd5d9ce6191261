"""Full-graph GraphSAGE on a vertex-partitioned graph across the GPUs of one NVLink domain: per layer the de-duplicated
halo rows are pulled from their owners over NVLink by one gather kernel on a side stream while the load-balanced g-SpMM
aggregates the interior rows; the dense half is one tcgen05 GEMM per layer (``--tc``), parameter gradients use the fused
P2P all-reduce. No NCCL call on the training path. (``bench.py --mode partitioned`` is the measured version of this on the
products-shaped graph with the repo's partitioner; this script builds a synthetic partition with a chosen edge cut.)

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/graphsage_partitioned_fullgraph.py
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-rank", type=int, default=300_000)
    ap.add_argument("--deg", type=int, default=20)
    ap.add_argument("--edge-cut", type=float, default=0.1, help="fraction of edges whose source lives on another rank")
    ap.add_argument("--feat", type=int, default=128)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--classes", type=int, default=47)
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--tc", action="store_true", help="bf16 tcgen05 dense half (PartitionedSAGETC)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from dgl_b200.parallel import P2PComm, PartitionedSAGE
    from dgl_b200.parallel.partitioned import PartitionedSAGETC
    comm = P2PComm()
    rows, deg = args.rows_per_rank, args.deg
    gen = torch.Generator().manual_seed(100 + rank)
    src = torch.randint(0, rows, (rows * deg,), generator=gen) + rank * rows
    if world > 1:
        peer = (rank + 1 + torch.randint(0, world - 1, (rows * deg,), generator=gen)) % world
        remote = peer * rows + torch.randint(0, rows // 5, (rows * deg,), generator=gen)
        src = torch.where(torch.rand(rows * deg, generator=gen) < args.edge_cut, remote, src)
    indptr = (torch.arange(rows + 1, dtype=torch.int32) * deg).to(dev)
    indices = src.to(torch.int32).to(dev)
    x = torch.randn(rows, args.feat, generator=gen).to(dev)
    y = torch.randint(0, args.classes, (rows,), generator=gen).to(dev)
    torch.manual_seed(0)                                           # identical initial weights on every rank
    cls = PartitionedSAGETC if args.tc else PartitionedSAGE
    model = cls(comm, indptr, indices, rows, args.feat, args.hidden, args.classes, num_layers=3).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    for epoch in range(args.epochs):
        torch.cuda.synchronize()
        t0 = time.time()
        loss = F.cross_entropy(model(x), y)
        opt.zero_grad()
        loss.backward()
        model.allreduce_gradients()
        opt.step()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"epoch {epoch:3d} loss {loss.item():.4f} {1e3 * (time.time() - t0):7.2f} ms "
                  f"({world * rows * deg * 3 / (time.time() - t0) / 1e9:.2f} G aggregated edges/s)")
    comm.check_error()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
