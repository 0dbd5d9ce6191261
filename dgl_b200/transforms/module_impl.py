"""Module-style (composable) transforms.

Parity: /root/reference/python/dgl/transforms/module.py (BaseTransform, RowFeatNormalizer,
FeatMask, RandomWalkPE, LapPE, AddSelfLoop, RemoveSelfLoop, AddReverse, ToSimple, LineGraph,
KHopGraph, AddMetaPaths, Compose, GCNNorm, PPR, HeatKernel, GDC, NodeShuffle, DropNode,
DropEdge, AddEdge, SIGNDiffusion, ToLevi, SVDPE).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import function as fn
from ..base import EID, NID, DGLError
from ..convert import graph as make_graph
from ..convert import heterograph as make_heterograph
from . import functional

__all__ = ["BaseTransform", "RowFeatNormalizer", "FeatMask", "RandomWalkPE", "LapPE", "LaplacianPE", "AddSelfLoop",
           "RemoveSelfLoop", "AddReverse", "ToSimple", "LineGraph", "KHopGraph", "AddMetaPaths", "Compose", "GCNNorm",
           "PPR", "HeatKernel", "GDC", "NodeShuffle", "DropNode", "DropEdge", "AddEdge", "SIGNDiffusion", "ToLevi",
           "SVDPE"]


class BaseTransform:
    def __call__(self, g):
        raise NotImplementedError

    def __repr__(self):
        return self.__class__.__name__ + "()"


class RowFeatNormalizer(BaseTransform):
    def __init__(self, subtract_min=False, node_feat_names=None, edge_feat_names=None):
        self.subtract_min = subtract_min
        self.node_feat_names = [] if node_feat_names is None else node_feat_names
        self.edge_feat_names = [] if edge_feat_names is None else edge_feat_names

    def row_normalize(self, feat):
        if self.subtract_min:
            feat = feat - feat.min()
        feat = feat / feat.sum(dim=-1, keepdim=True).clamp_(min=1.0)
        return feat

    def __call__(self, g):
        for name in self.node_feat_names:
            data = g.ndata[name]
            if isinstance(data, dict):
                g.ndata[name] = {k: self.row_normalize(v) for k, v in data.items()}
            else:
                g.ndata[name] = self.row_normalize(data)
        for name in self.edge_feat_names:
            data = g.edata[name]
            if isinstance(data, dict):
                g.edata[name] = {k: self.row_normalize(v) for k, v in data.items()}
            else:
                g.edata[name] = self.row_normalize(data)
        return g


class FeatMask(BaseTransform):
    def __init__(self, p=0.5, node_feat_names=None, edge_feat_names=None):
        self.p = p
        self.node_feat_names = [] if node_feat_names is None else node_feat_names
        self.edge_feat_names = [] if edge_feat_names is None else edge_feat_names

    def _mask(self, feat):
        m = torch.bernoulli(torch.full((feat.shape[-1],), self.p, device=feat.device)).bool()
        feat = feat.clone()
        feat[..., m] = 0
        return feat

    def __call__(self, g):
        if self.p == 0:
            return g
        for name in self.node_feat_names:
            data = g.ndata[name]
            g.ndata[name] = {k: self._mask(v) for k, v in data.items()} if isinstance(data, dict) else self._mask(data)
        for name in self.edge_feat_names:
            data = g.edata[name]
            g.edata[name] = {k: self._mask(v) for k, v in data.items()} if isinstance(data, dict) else self._mask(data)
        return g


class RandomWalkPE(BaseTransform):
    def __init__(self, k, feat_name="PE", eweight_name=None):
        self.k, self.feat_name, self.eweight_name = k, feat_name, eweight_name

    def __call__(self, g):
        g.ndata[self.feat_name] = functional.random_walk_pe(g, self.k, self.eweight_name)
        return g


class LapPE(BaseTransform):
    def __init__(self, k, feat_name="PE", eigval_name=None, padding=False):
        self.k, self.feat_name, self.eigval_name, self.padding = k, feat_name, eigval_name, padding

    def __call__(self, g):
        if self.eigval_name:
            pe, ev = functional.lap_pe(g, self.k, self.padding, return_eigval=True)
            g.ndata[self.eigval_name] = ev.unsqueeze(0).repeat(g.num_nodes(), 1)
        else:
            pe = functional.lap_pe(g, self.k, self.padding)
        g.ndata[self.feat_name] = pe
        return g


LaplacianPE = LapPE


class SVDPE(BaseTransform):
    def __init__(self, k, feat_name="svd_pe", padding=False, random_flip=True):
        self.k, self.feat_name, self.padding, self.random_flip = k, feat_name, padding, random_flip

    def __call__(self, g):
        g.ndata[self.feat_name] = functional.svd_pe(g, self.k, self.padding, self.random_flip)
        return g


class AddSelfLoop(BaseTransform):
    def __init__(self, allow_duplicate=False, new_etypes=False, edge_feat_names=None, fill_data=1.0):
        self.allow_duplicate = allow_duplicate
        self.new_etypes = new_etypes
        self.edge_feat_names = edge_feat_names
        self.fill_data = fill_data

    def transform_etype(self, c_etype, g):
        utype, _, vtype = c_etype
        if utype != vtype:
            return g
        if not self.allow_duplicate:
            g = functional.remove_self_loop(g, etype=c_etype)
        return functional.add_self_loop(g, edge_feat_names=self.edge_feat_names, fill_data=self.fill_data, etype=c_etype)

    def __call__(self, g):
        for c in g.canonical_etypes:
            g = self.transform_etype(c, g)
        if self.new_etypes:
            device, idtype = g.device, g.idtype
            data = {c: g.all_edges(etype=c) for c in g.canonical_etypes}
            for nt in g.ntypes:
                nids = torch.arange(g.num_nodes(nt), device=device, dtype=idtype)
                data[(nt, "self", nt)] = (nids, nids)
            new_g = make_heterograph(data, {nt: g.num_nodes(nt) for nt in g.ntypes}, idtype=idtype, device=device)
            for nt in g.ntypes:
                for k, v in g.nodes[nt].data.items():
                    new_g.nodes[nt].data[k] = v
            for c in g.canonical_etypes:
                for k, v in g.edges[c].data.items():
                    new_g.edges[c].data[k] = v
            g = new_g
        return g


class RemoveSelfLoop(BaseTransform):
    def __call__(self, g):
        for c in g.canonical_etypes:
            if c[0] == c[2]:
                g = functional.remove_self_loop(g, etype=c)
        return g


class AddReverse(BaseTransform):
    def __init__(self, copy_edata=False, sym_new_etype=False):
        self.copy_edata = copy_edata
        self.sym_new_etype = sym_new_etype

    def __call__(self, g):
        device, idtype = g.device, g.idtype
        data, feats = {}, {}
        for c in g.canonical_etypes:
            u, v = g.all_edges(etype=c)
            s, e, d = c
            if s == d and not self.sym_new_etype:
                data[c] = (torch.cat([u, v]), torch.cat([v, u]))
                feats[c] = ("dup", c)
            else:
                data[c] = (u, v)
                feats[c] = ("same", c)
                rc = (d, "rev_" + e, s)
                data[rc] = (v, u)
                feats[rc] = ("same" if self.copy_edata else "none", c)
        new_g = make_heterograph(data, {nt: g.num_nodes(nt) for nt in g.ntypes}, idtype=idtype, device=device) \
            if len(data) > 1 or not g.is_homogeneous else make_graph(data[g.canonical_etypes[0]], num_nodes=g.num_nodes(),
                                                                      idtype=idtype, device=device)
        for nt in g.ntypes:
            for k, v in g.nodes[nt].data.items():
                new_g.nodes[nt].data[k] = v
        for c, (mode, src_c) in feats.items():
            for k, v in g.edges[src_c].data.items():
                tgt = new_g.edges[c].data if not new_g.is_homogeneous else new_g.edata
                if mode == "same":
                    tgt[k] = v
                elif mode == "dup":
                    tgt[k] = torch.cat([v, v], 0) if self.copy_edata else torch.cat([v, torch.zeros_like(v)], 0)
        return new_g


class ToSimple(BaseTransform):
    def __init__(self, return_counts="count", aggregator="arbitrary"):
        self.return_counts, self.aggregator = return_counts, aggregator

    def __call__(self, g):
        return functional.to_simple(g, return_counts=self.return_counts, copy_edata=True, aggregator=self.aggregator)


class LineGraph(BaseTransform):
    def __init__(self, backtracking=True):
        self.backtracking = backtracking

    def __call__(self, g):
        return functional.line_graph(g, backtracking=self.backtracking, shared=True)


class KHopGraph(BaseTransform):
    def __init__(self, k):
        self.k = k

    def __call__(self, g):
        return functional.khop_graph(g, self.k)


class AddMetaPaths(BaseTransform):
    def __init__(self, metapaths, keep_orig_edges=True):
        self.metapaths, self.keep_orig_edges = metapaths, keep_orig_edges

    def __call__(self, g):
        data = {}
        for name, mp in self.metapaths.items():
            mg = functional.metapath_reachable_graph(g, mp)
            u, v = mg.all_edges()
            s = g.to_canonical_etype(mp[0])[0]
            d = g.to_canonical_etype(mp[-1])[2]
            data[(s, name, d)] = (u, v)
        if self.keep_orig_edges:
            for c in g.canonical_etypes:
                data[c] = g.all_edges(etype=c)
        new_g = make_heterograph(data, {nt: g.num_nodes(nt) for nt in g.ntypes}, idtype=g.idtype, device=g.device)
        for nt in g.ntypes:
            for k, v in g.nodes[nt].data.items():
                new_g.nodes[nt].data[k] = v
        if self.keep_orig_edges:
            for c in g.canonical_etypes:
                for k, v in g.edges[c].data.items():
                    new_g.edges[c].data[k] = v
        return new_g


class Compose(BaseTransform):
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, g):
        for t in self.transforms:
            g = t(g)
        return g

    def __repr__(self):
        return self.__class__.__name__ + "([\n" + ",\n".join("  " + str(t) for t in self.transforms) + "\n])"


class GCNNorm(BaseTransform):
    def __init__(self, eweight_name="w"):
        self.eweight_name = eweight_name

    def calc_etype(self, c_etype, g):
        ntype = c_etype[0]
        with g.local_scope():
            if self.eweight_name in g.edges[c_etype].data:
                w = g.edges[c_etype].data[self.eweight_name]
            else:
                w = torch.ones(g.num_edges(c_etype), device=g.device)
            g.edges[c_etype].data["_w"] = w
            g.update_all(fn.copy_e("_w", "m"), fn.sum("m", "deg"), etype=c_etype)
            deg_inv_sqrt = 1.0 / torch.sqrt(g.nodes[ntype].data["deg"])
            deg_inv_sqrt[torch.isinf(deg_inv_sqrt)] = 0
            u, v = g.all_edges(etype=c_etype)
            return deg_inv_sqrt[u.long()] * w * deg_inv_sqrt[v.long()]

    def __call__(self, g):
        result = {}
        for c in g.canonical_etypes:
            if c[0] != c[2]:
                continue
            result[c] = self.calc_etype(c, g)
        for c, v in result.items():
            g.edges[c].data[self.eweight_name] = v
        return g


def _dense_adj(g, eweight_name):
    n = g.num_nodes()
    u, v = g.all_edges()
    w = g.edata[eweight_name].float() if eweight_name in g.edata else torch.ones(g.num_edges(), device=g.device)
    a = torch.zeros(n, n, device=g.device)
    a.index_put_((u.long(), v.long()), w, accumulate=True)
    return a


def _graph_from_dense(g, mat, eweight_name):
    u, v = torch.nonzero(mat, as_tuple=True)
    new_g = make_graph((u, v), num_nodes=g.num_nodes(), idtype=g.idtype, device=g.device)
    new_g.edata[eweight_name] = mat[u, v]
    for k, val in g.ndata.items():
        new_g.ndata[k] = val
    return new_g


def _sparsify(mat, eps, avg_degree):
    if avg_degree is not None:
        n = mat.shape[0]
        k = min(avg_degree * n, mat.numel())
        thr = torch.topk(mat.flatten(), k)[0][-1] if k > 0 else float("inf")
        return torch.where(mat >= thr, mat, torch.zeros_like(mat))
    if eps is not None:
        return torch.where(mat >= eps, mat, torch.zeros_like(mat))
    return mat


class _Diffusion(BaseTransform):
    """Shared sparsification rule of PPR / HeatKernel / GDC."""
    eps, avg_degree = None, 5

    def get_eps(self, num_nodes, mat):
        """Threshold below which diffused weights are dropped: the explicit ``eps``, or the value that keeps
        ``avg_degree * num_nodes`` entries of ``mat`` (midpoint between the last kept and the first dropped value, as
        in the reference module.py:get_eps)."""
        if self.eps is not None:
            return self.eps
        vals = torch.sort(torch.as_tensor(mat).flatten(), descending=True)[0]
        k = self.avg_degree * num_nodes
        if k >= vals.numel():
            return float("-inf")
        left = vals[k - 1] if k > 0 else vals[0]
        return float((left + vals[k]) / 2.0)


class PPR(_Diffusion):
    def __init__(self, alpha=0.15, eweight_name="w", eps=None, avg_degree=5):
        self.alpha, self.eweight_name, self.eps, self.avg_degree = alpha, eweight_name, eps, avg_degree

    def __call__(self, g):
        a = _dense_adj(g, self.eweight_name)
        n = a.shape[0]
        ppr = self.alpha * torch.inverse(torch.eye(n, device=a.device) - (1 - self.alpha) * a)
        ppr = _sparsify(ppr, self.eps, self.avg_degree if self.eps is None else None)
        return _graph_from_dense(g, ppr, self.eweight_name)


class HeatKernel(_Diffusion):
    def __init__(self, t=2.0, eweight_name="w", eps=None, avg_degree=5):
        self.t, self.eweight_name, self.eps, self.avg_degree = t, eweight_name, eps, avg_degree

    def __call__(self, g):
        a = _dense_adj(g, self.eweight_name)
        n = a.shape[0]
        hk = torch.matrix_exp(self.t * (a - torch.eye(n, device=a.device)))
        hk = _sparsify(hk, self.eps, self.avg_degree if self.eps is None else None)
        return _graph_from_dense(g, hk, self.eweight_name)


class GDC(_Diffusion):
    def __init__(self, coefs, eweight_name="w", eps=None, avg_degree=5):
        self.coefs, self.eweight_name, self.eps, self.avg_degree = coefs, eweight_name, eps, avg_degree

    def __call__(self, g):
        a = _dense_adj(g, self.eweight_name)
        n = a.shape[0]
        cur = torch.eye(n, device=a.device)
        acc = self.coefs[0] * cur
        for c in self.coefs[1:]:
            cur = cur @ a
            acc = acc + c * cur
        acc = _sparsify(acc, self.eps, self.avg_degree if self.eps is None else None)
        return _graph_from_dense(g, acc, self.eweight_name)


class NodeShuffle(BaseTransform):
    def __call__(self, g):
        g = g.clone()
        for nt in g.ntypes:
            perm = torch.randperm(g.num_nodes(nt), device=g.device)
            fr = g._node_frames[g.get_ntype_id(nt)]
            for k in list(fr.keys()):
                fr[k] = fr[k][perm]
        return g


class DropNode(BaseTransform):
    def __init__(self, p=0.5):
        self.p = p

    def __call__(self, g):
        g = g.clone()
        if self.p == 0:
            return g
        for nt in g.ntypes:
            mask = torch.bernoulli(torch.full((g.num_nodes(nt),), self.p, device=g.device)).bool()
            g.remove_nodes(torch.nonzero(mask).squeeze(1), ntype=nt)
        return g


class DropEdge(BaseTransform):
    def __init__(self, p=0.5):
        self.p = p

    def __call__(self, g):
        g = g.clone()
        if self.p == 0:
            return g
        for c in g.canonical_etypes:
            mask = torch.bernoulli(torch.full((g.num_edges(c),), self.p, device=g.device)).bool()
            g.remove_edges(torch.nonzero(mask).squeeze(1), etype=c)
        return g


class AddEdge(BaseTransform):
    def __init__(self, ratio=0.2):
        self.ratio = ratio

    def __call__(self, g):
        g = g.clone()
        if self.ratio == 0.0:
            return g
        for c in g.canonical_etypes:
            s, _, d = c
            n = int(g.num_edges(c) * self.ratio)
            u = torch.randint(0, g.num_nodes(s), (n,), device=g.device, dtype=g.idtype)
            v = torch.randint(0, g.num_nodes(d), (n,), device=g.device, dtype=g.idtype)
            g.add_edges(u, v, etype=c)
        return g


class SIGNDiffusion(BaseTransform):
    def __init__(self, k, in_feat_name="feat", out_feat_name="out_feat", eweight_name=None, diffuse_op="raw",
                 alpha=0.2):
        self.k, self.in_feat_name, self.out_feat_name = k, in_feat_name, out_feat_name
        self.eweight_name, self.diffuse_op, self.alpha = eweight_name, diffuse_op, alpha
        if diffuse_op not in ("raw", "rw", "gcn", "ppr"):
            raise DGLError(f"Expect diffuse_op to be from ['raw', 'rw', 'gcn', 'ppr'], got {diffuse_op}")

    def _weights(self, g):
        w = g.edata[self.eweight_name].float() if self.eweight_name else torch.ones(g.num_edges(), device=g.device)
        u, v = g.all_edges()
        if self.diffuse_op == "raw":
            return w
        indeg = torch.zeros(g.num_nodes(), device=g.device).index_add_(0, v.long(), w)
        if self.diffuse_op == "rw":
            return w / indeg[v.long()].clamp(min=1e-12)
        dis = indeg.pow(-0.5)
        dis[torch.isinf(dis)] = 0
        return dis[u.long()] * w * dis[v.long()]

    def __call__(self, g):
        from .. import ops
        w = self._weights(g)
        x = g.ndata[self.in_feat_name]
        cur = x
        for i in range(1, self.k + 1):
            nxt = ops.u_mul_e_sum(g, cur, w.to(cur.dtype).unsqueeze(-1) if cur.dim() > 1 else w.to(cur.dtype))
            if self.diffuse_op == "ppr":
                nxt = (1 - self.alpha) * nxt + self.alpha * x
            cur = nxt
            g.ndata[f"{self.out_feat_name}_{i}"] = cur
        return g


class ToLevi(BaseTransform):
    """Levi graph: every edge becomes a node connected to its endpoints."""

    def __call__(self, g):
        if not g.is_homogeneous:
            raise DGLError("ToLevi only supports homogeneous graphs")
        u, v = g.all_edges()
        e = torch.arange(g.num_edges(), device=g.device, dtype=g.idtype)
        lg = make_heterograph({("edge", "e2n", "node"): (e, v), ("node", "n2e", "edge"): (u, e)},
                              {"node": g.num_nodes(), "edge": g.num_edges()}, idtype=g.idtype, device=g.device)
        for k, val in g.ndata.items():
            lg.nodes["node"].data[k] = val
        for k, val in g.edata.items():
            lg.nodes["edge"].data[k] = val
        return lg
