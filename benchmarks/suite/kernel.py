"""Sparse-kernel benchmarks (reference benchmarks/kernel/: bench_gspmm_copy_u, bench_gspmm_u_mul_e, bench_gsddmm_u_dot_v,
bench_edgesoftmax). ``benchmarks/bench_kernels.py`` is the roofline-oriented version on the BASELINE graph shapes."""
import torch

import dgl_b200 as dgl
from dgl_b200 import ops

from . import bench, timeit

_G = {}


def _graph(device, n=200_000, e=4_000_000):
    key = (str(device), n, e)
    if key not in _G:
        _G[key] = dgl.data.rmat_graph(n, e).to(device)
    return _G[key]


@bench("kernel", F=[32, 128], reducer=["sum", "max"])
def gspmm_copy_u(device, F, reducer):
    g = _graph(device)
    x = torch.randn(g.num_nodes(), F, device=device)
    f = getattr(ops, f"copy_u_{reducer}")
    return timeit(lambda: f(g, x), device)


@bench("kernel", F=[32, 128], heads=[1, 8])
def gspmm_u_mul_e_sum(device, F, heads):
    g = _graph(device)
    x = torch.randn(g.num_nodes(), heads, F // heads if heads > 1 else F, device=device)
    w = torch.rand(g.num_edges(), heads, 1, device=device)
    return timeit(lambda: ops.u_mul_e_sum(g, x, w), device)


@bench("kernel", F=[32, 128])
def gsddmm_u_dot_v(device, F):
    g = _graph(device)
    x = torch.randn(g.num_nodes(), F, device=device)
    return timeit(lambda: ops.u_dot_v(g, x, x), device)


@bench("kernel", heads=[1, 8])
def edge_softmax(device, heads):
    g = _graph(device)
    s = torch.randn(g.num_edges(), heads, device=device)
    return timeit(lambda: ops.edge_softmax(g, s), device)


@bench("kernel", F=[64])
def gspmm_backward(device, F):
    g = _graph(device)
    x = torch.randn(g.num_nodes(), F, device=device, requires_grad=True)
    w = torch.rand(g.num_edges(), 1, device=device, requires_grad=True)

    def step():
        ops.u_mul_e_sum(g, x, w).sum().backward()
    return timeit(step, device)
