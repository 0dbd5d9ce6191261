"""User-level functional sparse ops (``dgl.ops``).

Parity: /root/reference/python/dgl/ops/{spmm,sddmm,edge_softmax,segment,gather_mm}.py.
Generated names: ``u_{add,sub,mul,div}_e_{sum,max,min,mean}``, ``copy_{u,e}_{sum,...}``,
``{u,v,e}_{add,sub,mul,div,dot}_{u,v,e}``.
"""
from __future__ import annotations

import sys
from itertools import product

import torch

from ..base import DGLError
from . import sparse_autograd as SA

__all__ = ["gspmm", "gsddmm", "edge_softmax", "segment_reduce", "segment_softmax", "segment_mm", "gather_mm",
           "copy_u", "copy_v", "copy_e"]


def _gidx(g):
    if hasattr(g, "_graphs"):
        if len(g._graphs) != 1:
            raise DGLError("dgl.ops functions operate on graphs with a single edge type; slice with g[etype] first")
        return g._graphs[0]
    return g  # already a UnitGraph


def _unsq(x):
    return x.unsqueeze(-1) if x is not None and x.dim() == 1 else x


def gspmm(g, op, reduce_op, lhs_data, rhs_data):
    """Generalized SpMM: ``out[v] = reduce_{(u,e,v)} op(lhs[u], rhs[e])``
    (reference ops/spmm.py:39-116). ``mean`` = sum / clamp(in_degree, 1)."""
    gidx = _gidx(g)
    if op not in ("add", "sub", "mul", "div", "copy_lhs", "copy_rhs"):
        raise DGLError(f"unsupported binary op {op}")
    if reduce_op not in ("sum", "max", "min", "mean"):
        raise DGLError(f"unsupported reducer {reduce_op}")
    squeeze = False
    if op != "copy_rhs" and lhs_data is not None and lhs_data.dim() == 1:
        squeeze = True
    if op != "copy_lhs" and rhs_data is not None and rhs_data.dim() == 1 and (op == "copy_rhs" or squeeze or
                                                                              lhs_data is None or lhs_data.dim() == 1):
        squeeze = True
    lhs_data, rhs_data = _unsq(lhs_data), _unsq(rhs_data)
    red = "sum" if reduce_op == "mean" else reduce_op
    ret = SA.gspmm(gidx, op, red, lhs_data if op != "copy_rhs" else None, rhs_data if op != "copy_lhs" else None)
    if reduce_op == "mean":
        deg = gidx.in_degrees().clamp(min=1).to(ret.dtype)
        ret = ret / deg.view(-1, *([1] * (ret.dim() - 1)))
    if squeeze and ret.dim() > 1 and ret.shape[-1] == 1:
        ret = ret.squeeze(-1)
    return ret


def gsddmm(g, op, lhs_data, rhs_data, lhs_target="u", rhs_target="v"):
    """Generalized SDDMM: ``out[e] = op(lhs[lhs_target(e)], rhs[rhs_target(e)])``
    (reference ops/sddmm.py)."""
    gidx = _gidx(g)
    if op not in ("add", "sub", "mul", "div", "dot", "copy_lhs", "copy_rhs"):
        raise DGLError(f"unsupported binary op {op}")
    squeeze = False
    if lhs_data is not None and lhs_data.dim() == 1 and (rhs_data is None or rhs_data.dim() == 1):
        squeeze = op != "dot"
    lhs_data, rhs_data = _unsq(lhs_data), _unsq(rhs_data)
    out = SA.gsddmm(gidx, op, lhs_data if op != "copy_rhs" else None, rhs_data if op != "copy_lhs" else None,
                    lhs_target, rhs_target)
    if squeeze and out.shape[-1] == 1:
        out = out.squeeze(-1)
    return out


def edge_softmax(graph, logits, eids=None, norm_by="dst"):
    """Softmax of edge scores over each node's incoming (``norm_by='dst'``) or outgoing edges
    (reference ops/edge_softmax.py:12-153). One fused kernel forward, one backward."""
    from ..base import is_all
    if norm_by not in ("dst", "src"):
        raise DGLError("norm_by must be 'dst' or 'src'")
    if eids is not None and not is_all(eids):
        sub = graph.edge_subgraph(eids, relabel_nodes=False)
        return SA.edge_softmax(_gidx(sub), _unsq(logits), None, norm_by).view_as(logits)
    return SA.edge_softmax(_gidx(graph), _unsq(logits), None, norm_by).view_as(logits)


def segment_reduce(seglen, value, reducer="sum"):
    """Reduce contiguous segments of ``value`` (reference ops/segment.py:segment_reduce)."""
    if reducer not in ("sum", "max", "min", "mean"):
        raise DGLError(f"unsupported reducer {reducer}")
    offsets = torch.zeros(seglen.shape[0] + 1, dtype=torch.int64, device=value.device)
    torch.cumsum(seglen.to(value.device), 0, out=offsets[1:])
    v = _unsq(value)
    out = SA.segment_reduce(reducer, v, offsets)
    return out.squeeze(-1) if value.dim() == 1 else out


def segment_softmax(seglen, value):
    """Softmax inside each segment (reference ops/segment.py:segment_softmax)."""
    seglen = seglen.to(value.device)
    vmax = segment_reduce(seglen, value, "max")
    v = torch.exp(value - torch.repeat_interleave(vmax, seglen.long(), dim=0))
    vsum = segment_reduce(seglen, v, "sum")
    return v / torch.repeat_interleave(vsum, seglen.long(), dim=0)


def segment_mm(a, b, seglen_a):
    """``out[seg_r] = a[seg_r] @ b[r]`` (reference ops/gather_mm.py / kernel.cc:47-110).
    CUDA tensors use one grouped sm_100a kernel (no per-relation host loop, no D2H sync)."""
    from ..nn.functional_native import segment_mm as _segmm
    return _segmm(a, b, seglen_a)


def gather_mm(a, b, *, idx_b):
    """``out[i] = a[i] @ b[idx_b[i]]`` (reference ops/gather_mm.py)."""
    from ..nn.functional_native import gather_mm as _gmm
    return _gmm(a, b, idx_b)


def copy_u(g, x):
    return gsddmm(g, "copy_lhs", x, None, "u", "e")


def copy_v(g, x):
    return gsddmm(g, "copy_lhs", x, None, "v", "e")


def copy_e(g, x):
    return x


def _gen_spmm(op, red):
    name = f"u_{op}_e_{red}"

    def f(g, x, y):
        return gspmm(g, op, red, x, y)

    f.__name__ = name
    f.__doc__ = f"g-SpMM: message = u {op} e, aggregated by {red}."
    return name, f


def _gen_copy_reduce(target, red):
    name = f"copy_{target}_{red}"
    op = "copy_lhs" if target == "u" else "copy_rhs"

    def f(g, x):
        return gspmm(g, op, red, x if target == "u" else None, x if target == "e" else None)

    f.__name__ = name
    return name, f


def _gen_sddmm(lhs, op, rhs):
    name = f"{lhs}_{op}_{rhs}"

    def f(g, x, y):
        return gsddmm(g, op, x, y, lhs, rhs)

    f.__name__ = name
    f.__doc__ = f"g-SDDMM: out[e] = {lhs} {op} {rhs}."
    return name, f


_mod = sys.modules[__name__]
for _op, _red in product(("add", "sub", "mul", "div"), ("sum", "max", "min", "mean")):
    _n, _f = _gen_spmm(_op, _red)
    setattr(_mod, _n, _f)
    __all__.append(_n)
for _t, _red in product(("u", "e"), ("sum", "max", "min", "mean")):
    _n, _f = _gen_copy_reduce(_t, _red)
    setattr(_mod, _n, _f)
    __all__.append(_n)
for _l, _r in product("uve", "uve"):
    if _l == _r:
        continue
    for _op in ("add", "sub", "mul", "div", "dot"):
        _n, _f = _gen_sddmm(_l, _op, _r)
        setattr(_mod, _n, _f)
        __all__.append(_n)


def reshape_lhs_rhs(lhs_data, rhs_data):
    """Insert singleton feature dims after the row dim so both operands have the same rank, e.g. (N, 3, 1) and
    (E, 5, 3, 4) -> (N, 1, 3, 1) and (E, 5, 3, 4); broadcasting then lines up from the right
    (reference ops/spmm.py:13, ops/sddmm.py)."""
    ls, rs = tuple(lhs_data.shape), tuple(rhs_data.shape)
    if len(ls) != len(rs):
        nd = max(len(ls), len(rs))
        lhs_data = lhs_data.reshape((ls[0],) + (1,) * (nd - len(ls)) + ls[1:])
        rhs_data = rhs_data.reshape((rs[0],) + (1,) * (nd - len(rs)) + rs[1:])
    return lhs_data, rhs_data
