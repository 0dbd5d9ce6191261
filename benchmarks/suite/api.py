"""Graph-API micro-benchmarks (reference benchmarks/api/: bench_edge_ids, bench_in_degrees, bench_add_self_loop,
bench_batch, bench_unbatch, bench_to_block, bench_node_subgraph, bench_sample_neighbors, bench_random_walk, ...)."""
import torch

import dgl_b200 as dgl

from . import bench, timeit

_G = {}


def _graph(device, n=100_000, e=1_000_000):
    key = (str(device), n, e)
    if key not in _G:
        _G[key] = dgl.data.rmat_graph(n, e, idtype=torch.int64).to(device)
    return _G[key]


@bench("api", fmt=["coo", "csr", "csc"])
def format_conversion(device, fmt):
    g = _graph(device)
    return timeit(lambda: g.formats(fmt).create_formats_(), device)


@bench("api")
def in_degrees(device):
    g = _graph(device)
    return timeit(lambda: g.in_degrees(), device)


@bench("api")
def add_self_loop(device):
    g = _graph(device)
    return timeit(lambda: dgl.add_self_loop(g), device)


@bench("api")
def reverse(device):
    g = _graph(device)
    return timeit(lambda: dgl.reverse(g), device)


@bench("api", k=[1000, 20000])
def node_subgraph(device, k):
    g = _graph(device)
    nodes = torch.randperm(g.num_nodes(), device=device)[:k]
    return timeit(lambda: dgl.node_subgraph(g, nodes), device)


@bench("api", k=[1000, 20000])
def in_subgraph(device, k):
    g = _graph(device)
    nodes = torch.randperm(g.num_nodes(), device=device)[:k]
    return timeit(lambda: dgl.in_subgraph(g, nodes), device)


@bench("api", fanout=[5, 20])
def sample_neighbors(device, fanout):
    g = _graph(device)
    seeds = torch.randperm(g.num_nodes(), device=device)[:10000]
    return timeit(lambda: dgl.sampling.sample_neighbors(g, seeds, fanout), device)


@bench("api")
def to_block(device):
    g = _graph(device)
    seeds = torch.randperm(g.num_nodes(), device=device)[:10000]
    sg = dgl.sampling.sample_neighbors(g, seeds, 10)
    return timeit(lambda: dgl.to_block(sg, seeds), device)


@bench("api", batch=[64, 512])
def batch_unbatch(device, batch):
    gs = [dgl.rand_graph(30, 120).to(device) for _ in range(batch)]
    return timeit(lambda: dgl.unbatch(dgl.batch(gs)), device)


@bench("api", length=[4, 16])
def random_walk(device, length):
    g = _graph("cpu")
    seeds = torch.randperm(g.num_nodes())[:10000]
    return timeit(lambda: dgl.sampling.random_walk(g, seeds, length=length), "cpu")


@bench("api")
def to_simple_and_bidirected(device):
    g = _graph("cpu", 20000, 200000)
    return timeit(lambda: dgl.to_bidirected(dgl.to_simple(g)), "cpu")
