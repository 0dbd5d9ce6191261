"""Host/device time breakdown of the generic (DataLoader + dgl.nn) mini-batch pipeline with torch.profiler:
prints the top CPU-side and CUDA-side entries of a few steady-state steps (device-resident graph + features)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))


def main():
    import bench_uva_pipeline as B

    class A:
        pass
    args = A()
    args.resident, args.cache_mb, args.batch, args.workers, args.fanouts, args.amp = "device", 0, 1024, 0, [10, 10, 10], os.environ.get("AMP", "0") == "1"
    args.model = os.environ.get("MODEL", "sage")
    args.graphed = os.environ.get("GRAPHED", "0") == "1"
    args.table_input = os.environ.get("TABLE", "0") == "1"
    n, u, v, feat, labels, train = B.make_graph(float(os.environ.get("SCALE", "0.02")))
    model, opt, it, _, _ = B.run_ours(args, n, u, v, feat, labels, train)

    def step():
        if args.graphed:
            return args._trainer.step(next(it))
        blocks, x, y = next(it)
        loss = F.cross_entropy(model(blocks, x).float(), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(10):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            step()
        torch.cuda.synchronize()
    if not args.graphed:
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=30, max_name_column_width=90))


if __name__ == "__main__":
    main()
