"""torch.profiler breakdown (device time per kernel) of one full-graph epoch of the partitioned SAGE bench, 1 GPU."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))


def main():
    import bench_partitioned as BP
    holder = {}
    orig_print = print

    # reuse run_ours' setup by intercepting the step function through torch.optim.Adam.step counting
    args = types.SimpleNamespace(gpus=1, steps=3, warmup=2, scale=float(os.environ.get("SCALE", "1.0")), impl="ours")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        BP.run_ours(args)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))


if __name__ == "__main__":
    main()
