"""Benchmark suite (parity: /root/reference/benchmarks/benchmarks/{api,kernel,model_speed,multigpu} ASV suites).
Each benchmark is a function ``f(device, **params) -> seconds_per_iteration`` registered with ``@bench``;
``benchmarks/run.py`` sweeps the parameter grid and writes one JSON record per (benchmark, params)."""
import itertools
import time

REGISTRY = []


def bench(group, **param_grid):
    def deco(fn):
        REGISTRY.append((group, fn.__name__, fn, param_grid))
        return fn
    return deco


def timeit(fn, device, warmup=2, iters=5):
    import torch
    for _ in range(warmup):
        fn()
    if str(device).startswith("cuda"):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / 1e3 / iters
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) / iters


def grid(param_grid):
    keys = list(param_grid)
    for vals in itertools.product(*[param_grid[k] for k in keys]):
        yield dict(zip(keys, vals))


from . import api, kernel, model_acc, model_speed, multigpu  # noqa: E402,F401
