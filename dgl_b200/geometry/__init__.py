"""Geometry ops (parity: /root/reference/python/dgl/geometry/{fps,edge_coarsening}.py, src/geometry/*):
farthest point sampling (batched, GPU friendly) and neighbour-matching edge coarsening."""
from __future__ import annotations

import torch

from ..base import DGLError

__all__ = ["farthest_point_sampler", "neighbor_matching"]


def farthest_point_sampler(pos, npoints, start_idx=None):
    """pos: [B, N, C] -> indices [B, npoints]; iterative FPS, each step one batched distance update."""
    B, N, C = pos.shape
    dev = pos.device
    if pos.is_cuda and C <= 16 and pos.dtype in (torch.float32, torch.float16, torch.bfloat16, torch.float64):
        from .._graph_kernels import fps
        start = (torch.randint(0, N, (B,), device=dev) if start_idx is None
                 else torch.full((B,), int(start_idx), device=dev, dtype=torch.long))
        return fps(pos, int(npoints), start)   # one CTA per cloud, whole loop in one launch
    if dev.type == "cpu" and pos.dtype in (torch.float32, torch.float64):
        # host clouds: native loop, threads over clouds (or over points of a single large cloud) — csrc/cpu/spatial.cc
        from .._native import loader
        from ..base import dtype_code
        start = (torch.randint(0, N, (B,)) if start_idx is None else torch.full((B,), int(start_idx), dtype=torch.long))
        out = torch.empty((B, npoints), dtype=torch.long)
        p = pos.contiguous()
        loader.cpu_lib().fn("dglb200_cpu_fps", "ipllllpp")(dtype_code(p.dtype), p.data_ptr(), B, N, C, int(npoints),
                                                           start.data_ptr(), out.data_ptr())
        return out
    dist = torch.full((B, N), float("inf"), device=dev, dtype=pos.dtype)
    if start_idx is None:
        cur = torch.randint(0, N, (B,), device=dev)
    else:
        cur = torch.full((B,), int(start_idx), device=dev, dtype=torch.long)
    out = torch.empty((B, npoints), dtype=torch.long, device=dev)
    ar = torch.arange(B, device=dev)
    for i in range(npoints):
        out[:, i] = cur
        d = ((pos - pos[ar, cur].unsqueeze(1)) ** 2).sum(-1)
        dist = torch.minimum(dist, d)
        cur = dist.argmax(1)
    return out


def neighbor_matching(graph, e_weights=None, relabel_idx=True):
    """Greedy maximal matching used for graph coarsening (Metis/Graclus style): returns a cluster id per node."""
    if not graph.is_homogeneous:
        raise DGLError("The graph should be homogeneous")
    n = graph.num_nodes()
    if graph.device.type == "cuda":
        # handshake matching rounds on the symmetrised adjacency (csrc/cuda/graph_kernels.cu match_*_kernel)
        from .. import random as dgl_random
        from .._graph_kernels import neighbor_matching as _nm
        from ..transforms.functional import to_bidirected_coo
        su, sv, sw = to_bidirected_coo(graph, e_weights)
        from ..graph_index import coo_to_csr
        indptr, indices, eids = coo_to_csr(su, sv, n, graph.idtype)
        res = _nm(indptr, indices, eids, sw, n, dgl_random.next_seed())
        if relabel_idx:
            res = torch.unique(res, return_inverse=True)[1].to(graph.idtype)
        return res
    # host graphs: one native greedy sweep over a random node order (csrc/cpu/spatial.cc)
    from .._native import loader
    from ..base import id_code
    from ..graph_index import coo_to_csr
    from ..transforms.functional import to_bidirected_coo
    su, sv, sw = to_bidirected_coo(graph, e_weights)
    indptr, indices, eids = coo_to_csr(su, sv, n, graph.idtype)
    if sw is not None:
        sw = (sw if eids is None else sw[eids.long()]).float().contiguous()
    order = torch.randperm(n)
    res = torch.full((n,), -1, dtype=graph.idtype)
    loader.cpu_lib().fn("dglb200_cpu_neighbor_matching", "ippplpp")(
        id_code(graph.idtype), indptr.data_ptr(), indices.data_ptr(), 0 if sw is None else sw.data_ptr(), n,
        order.data_ptr(), res.data_ptr())
    if relabel_idx:
        _, res = torch.unique(res, return_inverse=True)
        res = res.to(graph.idtype)
    return res
