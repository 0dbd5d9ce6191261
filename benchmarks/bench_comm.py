"""Fused compute+collective paths vs their NCCL baselines (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/bench_comm.py

 1. halo g-SpMM: mean aggregation over a range-partitioned feature table; remote rows are read over NVLink inside the
    kernel.  Baseline = what the reference does: NCCL all_gather of the whole table, then a local g-SpMM.
 2. gradient all-reduce fused with Adam (peer loads inside the optimizer kernel) vs NCCL all_reduce + torch.optim.Adam.
 3. sharded-table row gather (P2P) vs NCCL all_to_all of indices + rows (the reference's sparse_all_to_all_pull).
Times are CUDA-event times, max over ranks; rank 0 prints one JSON line per measurement."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def timed(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import dgl_b200 as dgl
    from dgl_b200.parallel import P2PComm
    comm = P2PComm()
    out = []

    def emit(**kw):
        kw.update(n_gpus=world)
        if rank == 0:
            print(json.dumps(kw), flush=True)
            out.append(kw)

    # ---------------------------------------------------------------- 1. halo g-SpMM
    rows, F, deg = 300_000, 128, 20           # per-rank rows; global graph: world*rows nodes, deg in-edges per node
    N = rows * world
    from dgl_b200._sparse_ops import gspmm
    sh = comm.alloc(rows * F * 2)
    x_local = sh.view(torch.bfloat16)[: rows * F].view(rows, F)
    x_local.copy_(torch.randn(rows, F, device=dev))
    outb = torch.empty(rows, F, dtype=torch.bfloat16, device=dev)
    full = torch.empty(N, F, dtype=torch.bfloat16, device=dev)
    indptr = (torch.arange(rows + 1, dtype=torch.int32) * deg).to(dev)
    # cut = fraction of edges whose source lives on another rank: (world-1)/world is a random partition (worst case,
    # what hash partitioning gives), 0.1 is what an edge-cut partitioner (METIS / our LDG) leaves on community graphs;
    # remote sources of the 0.1 case come from a window of the peers' rows (boundary vertices are shared).
    for cut in ((world - 1) / world, 0.1) if world > 1 else (0.0,):
        gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
        local_src = torch.randint(0, rows, (rows * deg,), generator=gen) + rank * rows
        if world > 1:
            peer = (rank + 1 + torch.randint(0, world - 1, (rows * deg,), generator=gen)) % world
            span = rows if cut > 0.3 else rows // 5
            remote_src = peer * rows + torch.randint(0, span, (rows * deg,), generator=gen)
            is_remote = torch.rand(rows * deg, generator=gen) < cut
            src = torch.where(is_remote, remote_src, local_src)
        else:
            src = local_src
        indices = src.to(torch.int32).to(dev)                      # GLOBAL source ids
        plan = comm.halo_plan(indptr, indices, rows)
        pbuf, table = plan.alloc_features(F, torch.bfloat16)
        table[:rows].copy_(x_local)
        outp = torch.empty_like(outb)

        def baseline():  # what the reference does: NCCL all-gather of the table, then the local g-SpMM
            dist.all_gather_into_tensor(full, x_local)
            o, _, _ = gspmm((indptr, indices, None), "copy_lhs", "sum", full, None, rows)
            return o

        def per_edge():
            comm.halo_spmm(indptr, indices, sh, rows, F, torch.bfloat16, outb, mean=True)

        def pipelined():
            plan.spmm(pbuf, table, outp, mean=True)
        comm.barrier(0)
        ref = baseline().float() / deg
        per_edge()
        pipelined()
        torch.cuda.synchronize()
        err_e = (outb.float() - ref).abs().max().item()
        err_p = (outp.float() - ref).abs().max().item()
        comm.barrier(1)
        t_base = timed(baseline)
        comm.barrier(0)
        t_edge = timed(per_edge)
        comm.barrier(1)
        t_pipe = timed(pipelined)
        sweep, default_pw = {}, plan.puller_warps
        for pw in (1, 2, 4, 8):
            plan.puller_warps = pw
            comm.barrier(1)
            sweep[pw] = timed(pipelined)
        plan.puller_warps = default_pw
        comm.barrier(0)
        halo_bytes = plan.n_halo * F * 2
        local_bytes = rows * deg * F * 2 + rows * F * 2
        emit(bench="halo_spmm_mean", rows_per_rank=rows, F=F, deg=deg, dtype="bf16", edge_cut=cut,
             halo_rows=plan.n_halo, interior_rows=plan.n_interior,
             ms_pipelined_one_kernel=t_pipe, ms_per_edge_remote_loads=t_edge, ms_nccl_allgather_plus_spmm=t_base,
             speedup_pipelined_vs_nccl=t_base / t_pipe, ms_by_puller_warps=sweep, max_abs_err_pipelined=err_p, max_abs_err_per_edge=err_e,
             nvlink_in_GBps_pipelined=halo_bytes / t_pipe / 1e6,
             gather_GBps_pipelined=local_bytes / t_pipe / 1e6,
             note="roofline: max(halo bytes / 900 GB/s NVLink in, gathered bytes / HBM-L2); pipelined = HaloPlan.spmm")
        del plan

    # ---------------------------------------------------------------- 2. all-reduce fused with Adam
    # two sizes: the flagship GraphSAGE (0.21 M parameters: latency-bound, one-shot) and a 4 Mi-parameter model
    # (16 MiB of gradients: bandwidth-bound, two-shot = reduce-scatter + Adam on the slice + all-gather of the update)
    for numel in (212_992, 1 << 22):
        gbuf = comm.grad_buffer(numel)
        gbuf.copy_(torch.randn(numel, device=dev))
        param = torch.randn(numel, device=dev)
        dist.broadcast(param, 0)
        m, v = torch.zeros_like(param), torch.zeros_like(param)
        shadow = torch.zeros(numel, dtype=torch.bfloat16, device=dev)
        step = torch.zeros(1, dtype=torch.int64, device=dev)
        p2 = torch.nn.Parameter(param.clone())
        opt = torch.optim.Adam([p2], lr=1e-3, fused=True)
        g2 = gbuf.clone()
        shadow2 = torch.zeros_like(shadow)

        def fused():
            comm.fused_allreduce_adam(param, m, v, shadow, step)

        def nccl():
            g = g2.clone()
            dist.all_reduce(g)
            p2.grad = g.div_(world)
            opt.step()
            shadow2.copy_(p2.data)  # the bf16 copy the fused kernel also produces
        fused()
        nccl()
        torch.cuda.synchronize()
        err = (param - p2.data).abs().max().item()
        t_f = timed(fused)
        t_n = timed(nccl)
        two = world >= 4 and numel * 4 >= comm.TWO_SHOT_MIN_BYTES
        moved = numel * 4 * (2 * (world - 1) / world if two else (world - 1))
        emit(bench="allreduce_adam", numel=numel, algorithm="two-shot" if two else "one-shot", ms_fused=t_f,
             ms_nccl_allreduce_plus_fused_adam=t_n, speedup=t_n / t_f, max_abs_param_diff_after_1_step=err,
             bytes_moved_per_gpu=moved, pull_GBps_per_gpu=moved / t_f / 1e6)

    # ---------------------------------------------------------------- 3. sharded-table gather
    nq = 400_000
    idx = torch.randint(0, N, (nq,), device=dev)
    got = torch.empty(nq, F, dtype=torch.bfloat16, device=dev)
    t_g = timed(lambda: comm.gather(sh, rows, idx, got))

    def nccl_pull():  # the reference's sparse_all_to_all_pull: bucket ids by owner, exchange ids, exchange rows, un-permute
        owner = torch.div(idx, rows, rounding_mode="floor")
        order = torch.argsort(owner)
        counts = torch.bincount(owner, minlength=world)
        send_counts = counts.tolist()                       # host sync, as in python/dgl/cuda/nccl.py:80-83
        recv_counts_t = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts_t, counts)
        recv_counts = recv_counts_t.tolist()
        req = torch.empty(sum(recv_counts), dtype=idx.dtype, device=dev)
        dist.all_to_all_single(req, idx[order], recv_counts, send_counts)
        rows_out = x_local[req - rank * rows]
        back = torch.empty(nq, F, dtype=torch.bfloat16, device=dev)
        dist.all_to_all_single(back, rows_out, send_counts, recv_counts)
        res = torch.empty_like(back)
        res[order] = back
        return res
    chk = nccl_pull()
    assert torch.equal(chk, got)
    t_a = timed(nccl_pull)
    emit(bench="sharded_gather", queries=nq, F=F, ms_p2p=t_g, ms_nccl_all_to_all=t_a, speedup=t_a / t_g,
         GBps_per_gpu=nq * F * 2 / t_g / 1e6)
    comm.check_error()
    if rank == 0 and len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
