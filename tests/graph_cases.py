"""Small hand-written graph fixtures tagged by property (idea of
/root/reference/tests/utils/graph_cases.py:9-40, rewritten)."""
import torch

import dgl_b200 as dgl

_CASES = {}


def register(tags):
    def deco(f):
        for t in tags:
            _CASES.setdefault(t, []).append(f)
        return f
    return deco


def get_cases(tags=None, exclude=()):
    tags = list(_CASES) if tags is None else tags
    out = []
    for t in tags:
        for f in _CASES.get(t, []):
            if f not in out and not any(f in _CASES.get(e, []) for e in exclude):
                out.append(f)
    return out


@register(["homo", "path"])
def path5(idtype=torch.int64, device="cpu"):
    return dgl.graph(([0, 1, 2, 3], [1, 2, 3, 4]), idtype=idtype, device=device)


@register(["homo", "zero-degree", "small"])
def star_with_isolated(idtype=torch.int64, device="cpu"):
    return dgl.graph(([1, 2, 3, 4, 0], [0, 0, 0, 0, 5]), num_nodes=8, idtype=idtype, device=device)


@register(["homo", "multigraph"])
def multi(idtype=torch.int64, device="cpu"):
    return dgl.graph(([0, 0, 1, 1, 2, 3, 3], [1, 1, 2, 0, 0, 2, 3]), idtype=idtype, device=device)


@register(["homo", "random"])
def rand200(idtype=torch.int64, device="cpu"):
    gen = torch.Generator().manual_seed(7)
    u = torch.randint(0, 200, (1500,), generator=gen)
    v = torch.randint(0, 200, (1500,), generator=gen)
    return dgl.graph((u, v), num_nodes=200, idtype=idtype, device=device)


@register(["bipartite", "zero-degree"])
def bip(idtype=torch.int64, device="cpu"):
    return dgl.heterograph({("user", "plays", "game"): ([0, 1, 1, 2, 4], [0, 0, 1, 2, 2])},
                           {"user": 6, "game": 4}, idtype=idtype, device=device)


@register(["hetero"])
def hetero3(idtype=torch.int64, device="cpu"):
    return dgl.heterograph({
        ("user", "follows", "user"): ([0, 1, 2], [1, 2, 0]),
        ("user", "plays", "game"): ([0, 1, 2, 2], [0, 0, 1, 1]),
        ("developer", "develops", "game"): ([0, 1], [0, 1]),
    }, idtype=idtype, device=device)


@register(["block"])
def block_case(idtype=torch.int64, device="cpu"):
    g = rand200(idtype, device)
    frontier = dgl.in_subgraph(g, torch.arange(20))
    return dgl.to_block(frontier, torch.arange(20, dtype=idtype))
