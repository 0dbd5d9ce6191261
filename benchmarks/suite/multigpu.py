"""Multi-GPU benchmarks (reference benchmarks/multigpu/: bench_multigpu_sage, bench_multigpu_rgcn): spawn one process
per visible GPU, data-parallel fused GraphSAGE with the P2P all-reduce fused into Adam; returns seconds per step
(device-timed, max over ranks). Needs >= 2 CUDA devices; on smaller machines the benchmark reports an error record."""
import os
import socket

import torch

from . import bench


def _worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import dgl_b200 as dgl
    from dgl_b200.models import FusedSAGETrainer
    from dgl_b200.parallel import P2PComm
    comm = P2PComm()
    dev = comm.device
    g = dgl.data.rmat_graph(200_000, 4_000_000, seed=1).to(dev)
    x = torch.randn(g.num_nodes(), 100, generator=torch.Generator().manual_seed(1)).to(dev)
    y = torch.randint(0, 47, (g.num_nodes(),), generator=torch.Generator().manual_seed(2)).to(dev)
    tr = FusedSAGETrainer(g, x, y, 47, hidden=256, fanouts=[10, 25], batch_size=1000, comm=comm)
    gen = torch.Generator().manual_seed(100 + rank)
    for _ in range(10):
        tr.train_step(torch.randint(0, g.num_nodes(), (1000,), generator=gen))
    dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(40):
        tr.train_step(torch.randint(0, g.num_nodes(), (1000,), generator=gen))
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / 40 / 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put(float(t.item()))
    dist.barrier()
    dist.destroy_process_group()


@bench("multigpu")
def sage_data_parallel_step(device):
    world = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if world < 2:
        raise RuntimeError("needs >= 2 CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    sec = q.get(timeout=600)
    for p in procs:
        p.join(60)
    return sec
