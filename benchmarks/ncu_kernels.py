#!/usr/bin/env python
"""Launch a handful of g-SpMM / g-SDDMM cases once each (after one warm-up) so that ncu can capture them:

  ncu --set full --clock-control none -k regex:'spmm|sddmm|SpMM|SDDMM' -o gpurun_out/kern_ours \
      python benchmarks/ncu_kernels.py --impl ours --cases u_dot_v:128:float32,u_mul_e_sum:602:float32
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
from bench_kernels import SHAPES, rmat_edges  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--graph", default="reddit")
    ap.add_argument("--cases", default="u_dot_v:128:float32,u_mul_e_sum:602:float32,copy_u_sum:602:float32")
    a = ap.parse_args()
    if a.impl == "reference":
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        os.environ.setdefault("DGLBACKEND", "pytorch")
        import dgl
    else:
        sys.path.insert(0, ROOT)
        import dgl_b200 as dgl
    dev = torch.device("cuda")
    n, e, sym = SHAPES[a.graph]
    u, v = rmat_edges(n, e, sym, dev)
    g = dgl.graph((u, v), num_nodes=n, idtype=torch.int32, device=dev)
    g.create_formats_()
    E = g.num_edges()
    for case in a.cases.split(","):
        name, F, dt = case.split(":")
        dtype = getattr(torch, dt)
        x = torch.randn(n, int(F), device=dev).to(dtype)
        w = torch.rand(E, 1, device=dev).to(dtype)
        fn = {"u_dot_v": lambda: dgl.ops.u_dot_v(g, x, x), "u_mul_e_sum": lambda: dgl.ops.u_mul_e_sum(g, x, w),
              "copy_u_sum": lambda: dgl.ops.copy_u_sum(g, x)}[name]
        fn()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push(case)
        fn()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()


if __name__ == "__main__":
    main()
