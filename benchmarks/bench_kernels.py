#!/usr/bin/env python
"""Kernel micro-benchmarks with roofline fractions (metric conventions of the reference's ASV kernels:
/root/reference/benchmarks/benchmarks/kernel/bench_gspmm_copy_u.py:10-39, bench_gspmm_u_mul_e_sum.py,
bench_gsddmm_u_dot_v.py: 3 warm-up + 10 timed iterations, CUDA events).

  python benchmarks/bench_kernels.py [--scale 1.0] [--out profiles/kernels.json]
Graphs: synthetic RMAT with the node/edge counts of ogbn-arxiv, reddit (scaled), ogbn-products.
Bytes model for g-SpMM (copy_u, sum): E*4 (indices) + E*F*s (gathered rows, no reuse assumed) + N*F*s (output).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dgl_b200 as dgl  # noqa: E402
from dgl_b200 import ops  # noqa: E402


def timeit(fn, warmup=3, iters=10, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()  # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--out", default="gpurun_out/kernels.json")
    ap.add_argument("--graphs", default="ogbn-arxiv,reddit,ogbn-products")
    args = ap.parse_args()
    peaks = {"hbm_gbs": 6462.4}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    results = []
    shapes = {"ogbn-arxiv": (169343, 1166243, False), "reddit": (232965, 114615892 // 4, False),
              "ogbn-products": (2449029, 61859140, True)}
    for name in args.graphs.split(","):
        n, e, sym = shapes[name]
        n, e = int(n * args.scale), int(e * args.scale)
        g = dgl.data.rmat_graph(n, e, symmetrize=sym, device=dev, idtype=torch.int32)
        g.create_formats_()
        E = g.num_edges()
        for F in (32, 128, 256):
            for dtype in (torch.float32, torch.bfloat16):
                x = torch.randn(n, F, device=dev).to(dtype)
                s = x.element_size()
                t = timeit(lambda: ops.copy_u_sum(g, x), flush=flush)
                nbytes = E * 4 + E * F * s + n * F * s
                results.append({"kernel": "gspmm_copy_u_sum", "graph": name, "N": n, "E": E, "F": F,
                                "dtype": str(dtype).replace("torch.", ""), "ms": t, "GBps": nbytes / t / 1e6,
                                "frac_of_measured_hbm": nbytes / t / 1e6 / peaks["hbm_gbs"],
                                "GFLOPS_ref_convention": E * F / t / 1e6})
                w = torch.rand(E, 1, device=dev).to(dtype)
                t = timeit(lambda: ops.u_mul_e_sum(g, x, w), flush=flush)
                nbytes2 = nbytes + E * s
                results.append({"kernel": "gspmm_u_mul_e_sum", "graph": name, "N": n, "E": E, "F": F,
                                "dtype": str(dtype).replace("torch.", ""), "ms": t, "GBps": nbytes2 / t / 1e6,
                                "frac_of_measured_hbm": nbytes2 / t / 1e6 / peaks["hbm_gbs"],
                                "GFLOPS_ref_convention": 2 * E * F / t / 1e6})
                if E * F * s < 20e9:
                    t = timeit(lambda: ops.u_dot_v(g, x, x), flush=flush)
                    nb = E * 8 + 2 * E * F * s + E * s
                    results.append({"kernel": "gsddmm_u_dot_v", "graph": name, "N": n, "E": E, "F": F,
                                    "dtype": str(dtype).replace("torch.", ""), "ms": t, "GBps": nb / t / 1e6,
                                    "frac_of_measured_hbm": nb / t / 1e6 / peaks["hbm_gbs"],
                                    "GFLOPS_ref_convention": 2 * E * F / t / 1e6})
                del x, w
            print(json.dumps(results[-1]), flush=True)
        # fused edge softmax (8 heads)
        sc = torch.randn(E, 8, 1, device=dev)
        t = timeit(lambda: ops.edge_softmax(g, sc), flush=flush)
        results.append({"kernel": "edge_softmax_fwd_h8", "graph": name, "E": E, "ms": t,
                        "GBps": (E * 8 * 4 * 2 + E * 4) / t / 1e6})
        del g, sc
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump({"peaks": peaks, "results": results}, open(args.out, "w"), indent=1)
    for r in results:
        print(f"{r['kernel']:22s} {r['graph']:14s} F={r.get('F', '-'):>4} {r.get('dtype', ''):9s} {r['ms']:9.3f} ms "
              f"{r['GBps']:9.1f} GB/s  frac={r.get('frac_of_measured_hbm', 0):.3f}")


if __name__ == "__main__":
    main()
