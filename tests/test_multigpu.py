"""Multi-GPU (>= 2 B200 on one NVLink domain): P2P barrier / all-reduce / halo g-SpMM / sharded gather+scatter-add,
and data-parallel FusedSAGETrainer keeping replicas bit-identical. Each test spawns one process per GPU."""
import os
import socket
import traceback

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                                                  reason="needs >= 2 GPUs")]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _entry(rank, world, port, fn, q):
    try:
        os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                           "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from dgl_b200.parallel import P2PComm
        comm = P2PComm()
        res = fn(rank, world, comm)
        torch.cuda.synchronize()
        comm.check_error()
        comm.close()                      # collective teardown of the symmetric buffers (unmap peers, then free)
        dist.barrier()
        q.put((rank, "ok", res))
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        q.put((rank, "error", traceback.format_exc()))


def _run(fn, world=2, timeout=240):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, fn, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    try:
        for _ in range(world):
            rank, status, res = q.get(timeout=timeout)
            assert status == "ok", f"rank {rank} failed:\n{res}"
            out[rank] = res
    finally:
        for p in procs:
            p.join(20)
            if p.is_alive():
                p.kill()
    return out


# ---------------------------------------------------------------------------------------------- workers
def _w_allreduce_and_tables(rank, world, comm):
    dev = comm.device
    n = 4096
    buf = comm.alloc(n * 4)
    buf.view(torch.float32)[:n].copy_(torch.arange(n, device=dev, dtype=torch.float32) * (rank + 1))
    out = torch.empty(n, device=dev)
    for _ in range(5):  # repeated use exercises the barrier epochs
        comm.all_reduce(buf, out, n, scale=1.0 / world)
    expect = torch.arange(n, device=dev, dtype=torch.float32) * sum(r + 1 for r in range(world)) / world
    assert torch.allclose(out, expect)
    # range-sharded table: rows [rank*rows, (rank+1)*rows) live on `rank`
    rows, F = 1000, 32
    sh = comm.alloc(rows * F * 2)
    local = sh.view(torch.bfloat16)[: rows * F].view(rows, F)
    g = torch.Generator(device="cpu").manual_seed(5)
    full = torch.randn(world * rows, F, generator=g).bfloat16()
    local.copy_(full[rank * rows:(rank + 1) * rows].to(dev))
    comm.barrier(0)
    idx = torch.randint(0, world * rows, (3000,), device=dev)
    got = torch.empty(3000, F, dtype=torch.bfloat16, device=dev)
    comm.gather(sh, rows, idx, got)
    assert torch.equal(got.cpu(), full[idx.cpu()])
    # scatter-add into an fp32 sharded table from every rank
    acc = comm.alloc(rows * 8 * 4)
    acc.view(torch.float32)[: rows * 8].zero_()
    comm.barrier(1)
    gi = torch.Generator(device="cpu").manual_seed(100 + rank)
    sidx = torch.randint(0, world * rows, (2000,), generator=gi)
    vals = torch.randn(2000, 8, generator=gi)
    comm.scatter_add(acc, rows, sidx.to(dev), vals.to(dev))
    comm.barrier(2)
    mine = acc.view(torch.float32)[: rows * 8].view(rows, 8).cpu()
    ref = torch.zeros(world * rows, 8)
    for r in range(world):
        gr = torch.Generator(device="cpu").manual_seed(100 + r)
        si = torch.randint(0, world * rows, (2000,), generator=gr)
        va = torch.randn(2000, 8, generator=gr)
        ref.index_add_(0, si, va)
    assert torch.allclose(mine, ref[rank * rows:(rank + 1) * rows], atol=1e-4)
    return True


def _w_halo_spmm(rank, world, comm):
    import dgl_b200 as dgl
    dev = comm.device
    rows, F = 1500, 64
    N = world * rows
    g = dgl.rand_graph(N, 40000, idtype=torch.int32)  # same seed on every rank? no: build from a fixed generator
    gen = torch.Generator().manual_seed(1)
    u = torch.randint(0, N, (40000,), generator=gen).int()
    v = torch.randint(0, N, (40000,), generator=gen).int()
    g = dgl.graph((u, v), num_nodes=N, idtype=torch.int32)
    full = torch.randn(N, F, generator=gen)
    ref = dgl.ops.copy_u_mean(g, full)
    # this rank owns destination rows [rank*rows, (rank+1)*rows): local CSC with GLOBAL source ids
    indptr, indices, _ = g._graphs[0].csc()
    lo, hi = rank * rows, (rank + 1) * rows
    ip = (indptr[lo:hi + 1] - indptr[lo]).to(dev)
    ix = indices[int(indptr[lo]): int(indptr[hi])].to(dev)
    sh = comm.alloc(rows * F * 2)
    sh.view(torch.bfloat16)[: rows * F].view(rows, F).copy_(full[lo:hi].to(dev))
    comm.barrier(0)
    out = torch.empty(rows, F, dtype=torch.bfloat16, device=dev)
    comm.halo_spmm(ip, ix, sh, rows, F, torch.bfloat16, out, mean=True)
    comm.barrier(1)
    assert (out.float().cpu() - ref[lo:hi]).abs().max() < 0.06
    # the planned, one-kernel pipelined variant (deduplicated halo pull overlapped with the aggregation)
    plan = comm.halo_plan(ip, ix, rows)
    pbuf, table = plan.alloc_features(F, torch.bfloat16)
    table[:rows].copy_(full[lo:hi].to(dev))
    comm.barrier(0)
    for _ in range(3):  # epochs advance; flags are reused
        out2 = torch.zeros(rows, F, dtype=torch.bfloat16, device=dev)
        plan.spmm(pbuf, table, out2, mean=True)
    comm.barrier(1)
    assert plan.n_halo > 0 and (out2.float().cpu() - ref[lo:hi]).abs().max() < 0.06
    # backward: dX (fp32 shards) accumulates contributions from every rank's destinations
    dsh = comm.alloc(rows * F * 4)
    dsh.view(torch.float32)[: rows * F].zero_()
    comm.barrier(2)
    dout_full = torch.randn(N, F, generator=gen)
    comm.halo_spmm_backward(ip, ix, dsh, rows, F, dout_full[lo:hi].to(dev), mean=False)
    comm.barrier(3)
    dref = torch.zeros(N, F).index_add_(0, u.long(), dout_full[v.long()])
    mine = dsh.view(torch.float32)[: rows * F].view(rows, F).cpu()
    assert torch.allclose(mine, dref[lo:hi], atol=2e-3)
    return True


def _w_trainer(rank, world, comm):
    import dgl_b200 as dgl
    from dgl_b200.models import FusedSAGETrainer
    dev = comm.device
    gen = torch.Generator().manual_seed(3)
    n = 20000
    u = torch.randint(0, n, (300000,), generator=gen).int()
    v = torch.randint(0, n, (300000,), generator=gen).int()
    g = dgl.graph((u, v), num_nodes=n, idtype=torch.int32).to(dev)
    feat = torch.randn(n, 64, generator=gen).to(dev)
    labels = torch.randint(0, 7, (n,), generator=gen).to(dev)
    tr = FusedSAGETrainer(g, feat, labels, 7, hidden=64, fanouts=[5, 5], batch_size=256, lr=1e-2, dropout=0.0,
                          comm=comm, seed=11, use_cuda_graph=True)
    p0 = tr.param.clone()
    losses = []
    for step in range(12):
        gs = torch.Generator().manual_seed(1000 + step * world + rank)  # different seeds per rank (data parallel)
        losses.append(tr.train_step(torch.randint(0, n, (256,), generator=gs)))
    assert all(l == l for l in losses)
    assert not torch.equal(tr.param, p0)
    # replicas must stay bit-identical: every rank applies the same all-reduced gradient
    mine = tr.param.clone()
    import torch.distributed as dist
    ref = mine.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(mine, ref), (mine - ref).abs().max()
    return losses[-1]


def _w_fused_allreduce_adam(rank, world, comm):
    import torch.distributed as dist
    dev = comm.device
    for two_shot, numel in ((False, 4096), (True, 10240), (True, 4100)):
        g = torch.Generator().manual_seed(5)
        param = torch.randn(numel, generator=g).to(dev)
        ref = torch.nn.Parameter(param.clone())
        opt = torch.optim.Adam([ref], lr=1e-2, weight_decay=0.01)
        m, v = torch.zeros_like(param), torch.zeros_like(param)
        shadow = torch.zeros(numel, dtype=torch.bfloat16, device=dev)
        step = torch.zeros(1, dtype=torch.int64, device=dev)
        gbuf = comm.grad_buffer(numel)
        for it in range(3):
            gr = torch.Generator().manual_seed(100 * it + rank)
            grad = torch.randn(numel, generator=gr).to(dev)
            gbuf.copy_(grad)
            comm.fused_allreduce_adam(param, m, v, shadow, step, lr=1e-2, weight_decay=0.01, two_shot=two_shot)
            tot = grad.clone()
            dist.all_reduce(tot)
            ref.grad = tot / world
            opt.step()
        assert int(step.item()) == 3
        assert torch.allclose(param, ref.data, atol=1e-5), (two_shot, (param - ref.data).abs().max())
        assert torch.allclose(shadow.float(), ref.data, atol=2e-2)
        chk = param.clone()
        dist.broadcast(chk, 0)
        assert torch.equal(chk, param)          # replicas identical
    return True


def _w_partitioned_sage(rank, world, comm):
    """Vertex-partitioned full-graph SAGE must reproduce the single-process model: same logits, same gradients."""
    import dgl_b200 as dgl
    from dgl_b200.parallel import PartitionedSAGE
    dev = comm.device
    rows, Fin, C = 1200, 32, 5
    N = rows * world
    gen = torch.Generator().manual_seed(4)
    u = torch.randint(0, N, (30000,), generator=gen).int()
    v = torch.randint(0, N, (30000,), generator=gen).int()
    g = dgl.graph((u, v), num_nodes=N, idtype=torch.int32)
    x = torch.randn(N, Fin, generator=gen)
    y = torch.randint(0, C, (N,), generator=gen)
    indptr, indices, _ = g._graphs[0].csc()
    lo, hi = rank * rows, (rank + 1) * rows
    ip = (indptr[lo:hi + 1] - indptr[lo]).to(dev)
    ix = indices[int(indptr[lo]): int(indptr[hi])].to(dev)
    torch.manual_seed(0)
    model = PartitionedSAGE(comm, ip, ix, rows, Fin, 16, C, num_layers=2).to(dev)
    logits = model(x[lo:hi].to(dev))
    loss = torch.nn.functional.cross_entropy(logits, y[lo:hi].to(dev), reduction="sum") / N
    loss.backward()
    model.allreduce_gradients()
    # single-process oracle with the same weights on the whole graph (the all-reduce averages: oracle grads / world)
    import copy
    sl, nl = copy.deepcopy(model.self_lin).cpu(), copy.deepcopy(model.neigh_lin).cpu()
    for p_ in list(sl.parameters()) + list(nl.parameters()):
        p_.grad = None
    h = x
    for i in range(2):
        h = sl[i](h) + nl[i](dgl.ops.copy_u_mean(g, h))
        if i == 0:
            h = torch.relu(h)
    ref_loss = torch.nn.functional.cross_entropy(h, y, reduction="sum") / N
    ref_loss.backward()
    assert torch.allclose(logits.cpu(), h[lo:hi], atol=1e-4)
    for a, b in zip(list(model.self_lin.parameters()) + list(model.neigh_lin.parameters()),
                    list(sl.parameters()) + list(nl.parameters())):
        assert torch.allclose(a.grad.cpu() * world, b.grad, atol=1e-4), (a.grad.cpu() * world - b.grad).abs().max()
    return float(loss)


def _w_partitioned_sage_pull_backward(rank, world, comm):
    """Pull-based backward (second halo plan over the out-edges) and the tcgen05 dense half: gradients of the partitioned
    model vs a single-process fp32 oracle; also the fused one-kernel pipeline vs the two-phase aggregation."""
    import os
    import dgl_b200 as dgl
    from dgl_b200.parallel import PartitionedSAGE
    from dgl_b200.parallel.partitioned import PartitionedSAGETC
    dev = comm.device
    rows, Fin, C = 1024, 32, 8
    N = rows * world
    gen = torch.Generator().manual_seed(5)
    u = torch.cat([torch.randint(0, N, (30000,), generator=gen), torch.randint(0, N, (4000,), generator=gen)]).int()
    v = torch.cat([torch.randint(0, N, (30000,), generator=gen), torch.full((4000,), 17)]).int()      # one hub row
    g = dgl.graph((u, v), num_nodes=N, idtype=torch.int32)
    x = torch.randn(N, Fin, generator=gen)
    y = torch.randint(0, C, (N,), generator=gen)
    lo, hi = rank * rows, (rank + 1) * rows

    def rows_of(fmt):
        indptr, indices, _ = fmt
        return (indptr[lo:hi + 1] - indptr[lo]).to(dev), indices[int(indptr[lo]): int(indptr[hi])].to(dev)

    ip, ix = rows_of(g._graphs[0].csc())
    rp, rx = rows_of(g._graphs[0].csr())
    outs = {}
    for mode in ("balanced", "fused"):
        os.environ["DGLB200_HALO_MODE"] = mode
        torch.manual_seed(0)
        model = PartitionedSAGE(comm, ip, ix, rows, Fin, 16, C, num_layers=2, reverse=(rp, rx)).to(dev)
        xin = x[lo:hi].to(dev).requires_grad_()
        logits = model(xin)
        (torch.nn.functional.cross_entropy(logits, y[lo:hi].to(dev), reduction="sum") / N).backward()
        outs[mode] = (logits.detach().cpu(), xin.grad.cpu(), [p_.grad.cpu().clone() for p_ in model.parameters()])
    os.environ.pop("DGLB200_HALO_MODE")
    import copy
    sl, nl = copy.deepcopy(model.self_lin).cpu(), copy.deepcopy(model.neigh_lin).cpu()
    for p_ in list(sl.parameters()) + list(nl.parameters()):
        p_.grad = None
    xo = x.clone().requires_grad_()
    h = xo
    for i in range(2):
        h = sl[i](h) + nl[i](dgl.ops.copy_u_mean(g, h))
        if i == 0:
            h = torch.relu(h)
    (torch.nn.functional.cross_entropy(h, y, reduction="sum") / N).backward()
    for mode, (lg, gx, _) in outs.items():
        assert torch.allclose(lg, h[lo:hi].detach(), atol=1e-4), mode
        assert torch.allclose(gx, xo.grad[lo:hi], atol=1e-5), (mode, (gx - xo.grad[lo:hi]).abs().max())
    # tensor-core variant (bf16): same function within bf16 tolerance, including the transform-first last layer (16 -> 8)
    torch.manual_seed(0)
    tc = PartitionedSAGETC(comm, ip, ix, rows, Fin, 16, C, num_layers=2, reverse=(rp, rx)).to(dev)
    lt = tc(x[lo:hi].to(dev)).float()
    assert (lt.cpu() - h[lo:hi].detach()).abs().max() < 0.08 * h.detach().abs().max().clamp(min=1.0)
    (torch.nn.functional.cross_entropy(lt, y[lo:hi].to(dev), reduction="sum") / N).backward()
    tc.allreduce_gradients(average=False)
    for a, b in zip(list(tc.self_lin.parameters()) + list(tc.neigh_lin.parameters()),
                    list(sl.parameters()) + list(nl.parameters())):
        scale = b.grad.abs().max().clamp(min=1e-3)
        assert (a.grad.cpu() - b.grad).abs().max() < 0.08 * scale, ((a.grad.cpu() - b.grad).abs().max(), scale)
    return True


def _w_p2p_cooperative(rank, world, comm):
    import torch.distributed as dist
    from dgl_b200.graphbolt.impl.cooperative import P2PCooperativeConv
    dev = comm.device
    cap, F = 700, 32
    n_local = 500 + 100 * rank                                  # ranks own different numbers of rows
    gen = torch.Generator().manual_seed(10 + rank)
    h = torch.randn(n_local, F, generator=gen).to(dev).requires_grad_()
    sizes = [500 + 100 * r for r in range(world)]
    owner = torch.randint(0, world, (2000,), generator=gen)
    local = (torch.rand(2000, generator=gen) * torch.tensor(sizes, dtype=torch.float)[owner]).long()
    ex = P2PCooperativeConv(comm, cap)
    need = ex.address(owner, local, cap).to(dev)
    out = ex(h, need)
    wgt = torch.randn(2000, F, generator=gen).to(dev)
    (out * wgt).sum().backward()
    # oracle through NCCL: gather every rank's rows, index; gradient = sum over ranks of the scattered weights
    allh = [torch.zeros(cap, F, device=dev) for _ in range(world)]
    pad = torch.zeros(cap, F, device=dev)
    pad[:n_local] = h.detach()
    dist.all_gather(allh, pad)
    ref = torch.stack(allh)[owner.to(dev), local.to(dev)]
    assert torch.allclose(out, ref)
    gfull = torch.zeros(world, cap, F, device=dev)
    gfull.index_put_((owner.to(dev), local.to(dev)), wgt, accumulate=True)
    dist.all_reduce(gfull)
    assert torch.allclose(h.grad, gfull[rank, :n_local], atol=1e-4)
    return True


def _w_trainer_nccl_baseline(rank, world, comm):
    """comm='nccl' (baseline path) must give the same parameters as the fused P2P path."""
    import dgl_b200 as dgl
    from dgl_b200.models import FusedSAGETrainer
    dev = comm.device
    gen = torch.Generator().manual_seed(3)
    n = 8000
    g = dgl.graph((torch.randint(0, n, (100000,), generator=gen).int(), torch.randint(0, n, (100000,), generator=gen).int()),
                  num_nodes=n, idtype=torch.int32).to(dev)
    feat = torch.randn(n, 64, generator=gen).to(dev)
    labels = torch.randint(0, 7, (n,), generator=gen).to(dev)
    params = []
    for c in (comm, "nccl"):
        tr = FusedSAGETrainer(g, feat, labels, 7, hidden=64, fanouts=[5, 5], batch_size=256, lr=1e-2, dropout=0.0, comm=c,
                              seed=11, use_cuda_graph=False)
        for step in range(4):
            gs = torch.Generator().manual_seed(1000 + step * world + rank)
            tr.train_step(torch.randint(0, n, (256,), generator=gs))
        params.append(tr.param.clone())
    assert torch.allclose(params[0], params[1], atol=2e-3), (params[0] - params[1]).abs().max()
    return True


def _w_nvls(rank, world, comm):
    """Multicast path (VMM symmetric heap + multimem): barrier, in-switch all-reduce fused with Adam, replicas identical.
    Returns 'nvls' or 'unavailable: ...' so the log shows which path the box offers."""
    import ctypes
    import torch.distributed as dist
    from dgl_b200._native import loader
    dev = comm.device
    if not comm.nvls:
        return "unavailable: " + str(getattr(comm, "_nvls_error", getattr(getattr(comm, "heap", None), "_why", "no multicast support")))
    lib = loader.cuda_lib()
    numel = 10240 + 4
    pad = lambda b: (b + 255) // 256 * 256  # noqa: E731
    o_param, o_sh = pad(4 * numel), pad(4 * numel) * 2
    buf = comm.heap.alloc(o_sh + pad(2 * numel))
    assert buf.mc_ptr
    grad, param = buf.view(torch.float32, 0, numel), buf.view(torch.float32, o_param, numel)
    shadow = buf.view(torch.bfloat16, o_sh, numel)
    g = torch.Generator().manual_seed(5)
    init = torch.randn(numel, generator=g).to(dev)
    param.copy_(init)
    ref = torch.nn.Parameter(init.clone())
    opt = torch.optim.Adam([ref], lr=1e-2, weight_decay=0.01)
    m, v = torch.zeros(numel, device=dev), torch.zeros(numel, device=dev)
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    for it in range(3):
        gr = torch.Generator().manual_seed(100 * it + rank)
        mine = torch.randn(numel, generator=gr).to(dev)
        grad.copy_(mine)
        comm.nvls_barrier(0)
        lib.fn("dglb200_symm_allreduce_adam", "pppppppp" + "lii" + "fffff" + "pp")(
            buf.mc_ptr, buf.mc_ptr + o_param, param.data_ptr(), m.data_ptr(), v.data_ptr(), buf.mc_ptr + o_sh, 0, 0, numel,
            world, rank, 1e-2, 0.9, 0.999, 1e-8, 0.01, step.data_ptr(), loader.stream_ptr())
        comm.nvls_barrier(1)
        tot = mine.clone()
        dist.all_reduce(tot)
        ref.grad = tot / world
        opt.step()
        torch.cuda.synchronize()
        assert torch.allclose(param, ref.detach(), atol=1e-5, rtol=1e-5), (it, (param - ref.detach()).abs().max())
        assert torch.equal(shadow.float(), param.bfloat16().float())
        other = param.clone()
        dist.broadcast(other, 0)
        assert torch.equal(other, param)            # one writer per element: replicas bit-identical
    assert int(step.item()) == 3
    return "nvls"


def _w_trainer_tc(rank, world, comm):
    """Data-parallel tcgen05 engine (bf16 and fp32) over the communicator (NVLS when available): replicas bit-identical,
    parameters equal to the NCCL-baseline run."""
    import dgl_b200 as dgl
    import torch.distributed as dist
    from dgl_b200.models import FusedSAGETrainer
    dev = comm.device
    gen = torch.Generator().manual_seed(3)
    n = 12000
    g = dgl.graph((torch.randint(0, n, (200000,), generator=gen).int(), torch.randint(0, n, (200000,), generator=gen).int()),
                  num_nodes=n, idtype=torch.int32).to(dev)
    feat = torch.randn(n, 64, generator=gen).to(dev)
    labels = torch.randint(0, 7, (n,), generator=gen).to(dev)
    out = {}
    for dtype in (torch.bfloat16, torch.float32):
        params = []
        for c in (comm, "nccl"):
            tr = FusedSAGETrainer(g, feat, labels, 7, hidden=128, fanouts=[5, 5], batch_size=256, lr=1e-2, dropout=0.0,
                                  dtype=dtype, comm=c, seed=11, use_cuda_graph=(c is comm))
            assert tr.use_tc
            for step in range(6):
                gs = torch.Generator().manual_seed(1000 + step * world + rank)
                tr.train_step(torch.randint(0, n, (256,), generator=gs))
            mine = tr.param.clone()
            ref = mine.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(mine, ref), (dtype, (mine - ref).abs().max())
            params.append(mine)
        assert torch.allclose(params[0], params[1], atol=2e-3), (dtype, (params[0] - params[1]).abs().max())
        out[str(dtype)] = "nvls" if tr is not None and comm.nvls else "ipc"
    return out


# ---------------------------------------------------------------------------------------------- tests
def test_nvls_multicast_allreduce_adam():
    res = _run(_w_nvls)
    print("NVLS status:", res)
    assert len(set(res.values())) == 1


def test_trainer_tc_data_parallel():
    res = _run(_w_trainer_tc)
    print("trainer comm path:", res)


def test_trainer_nccl_baseline_matches_p2p():
    _run(_w_trainer_nccl_baseline)



def test_p2p_cooperative_exchange():
    _run(_w_p2p_cooperative)



def test_partitioned_fullgraph_sage_matches_single_process():
    _run(_w_partitioned_sage)


def test_partitioned_sage_pull_backward_and_tensor_core_dense():
    _run(_w_partitioned_sage_pull_backward)



def test_fused_allreduce_adam_one_and_two_shot():
    _run(_w_fused_allreduce_adam)



def test_p2p_allreduce_gather_scatter():
    _run(_w_allreduce_and_tables)


def test_halo_spmm_two_ranks():
    _run(_w_halo_spmm)


def test_trainer_replicas_identical():
    res = _run(_w_trainer)
    assert len(res) == 2
