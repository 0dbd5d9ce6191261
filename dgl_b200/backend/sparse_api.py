"""Graph-level sparse algebra and typed message passing of the backend API.

The reference exposes these through ``dgl.backend`` (backend/backend.py:1600-1990, implemented in
backend/pytorch/sparse.py as autograd Functions over its C++ ``_CAPI_DGLCSRMM`` / ``CSRSum`` / ``CSRMask`` and the
``*_hetero`` kernels). Here a "graph index" is the graph object itself (a ``DGLGraph`` with one relation, or its
``UnitGraph``): weights are edge-aligned tensors, and the algebra runs on the package's own sparse-matrix kernels
(``dgl_b200.sparse``: native Gustavson SpGEMM on the host, hash SpGEMM on the GPU, both differentiable), so gradients
flow to the edge weights without a dedicated Function per operator.

Typed (heterogeneous) g-SpMM / g-SDDMM / edge-softmax take the reference's flat tuple conventions and dispatch one
relation at a time onto the single-relation kernels, accumulating per destination node type.
"""
from __future__ import annotations

import torch

from ..base import DGLError


# ----------------------------------------------------------------------------------------------------- adapters ----
def _unit(g):
    """Single-relation structure of ``g`` plus its (num_src, num_dst)."""
    if hasattr(g, "_graphs"):
        if len(g._graphs) != 1:
            raise DGLError("sparse algebra works on graphs with exactly one edge type")
        g = g._graphs[0]
    return g


def _as_matrix(g, weights):
    from ..sparse import SparseMatrix
    ug = _unit(g)
    # SparseMatrix keeps the relation column -> row (SpMM aggregates into rows), i.e. the reverse of the graph's
    # src -> dst; ``reverse`` preserves the edge order, so the weights stay aligned
    return SparseMatrix(ug.reverse(), weights, (ug.num_src, ug.num_dst))


def _as_graph(mat, like, num_vtypes):
    """Wrap a result matrix as a graph: one node type when the operands said so (``num_vtypes == 1``), else bipartite."""
    from ..convert import graph as make_graph
    from ..convert import heterograph as make_heterograph
    row, col = mat.coo()
    idtype = getattr(like, "idtype", torch.int64)
    if num_vtypes == 1:
        return make_graph((row, col), num_nodes=max(mat.shape), idtype=idtype, device=row.device)
    return make_heterograph({("_U", "_E", "_V"): (row, col)}, {"_U": mat.shape[0], "_V": mat.shape[1]}, idtype=idtype,
                            device=row.device)


# ------------------------------------------------------------------------------------------------ matrix algebra ----
def csrmm(gidxA, A_weights, gidxB, B_weights, num_vtypes):
    """``C = A @ B`` over edge-weighted graphs. Returns ``(graph of C, weights of C)``; differentiable in both
    weight tensors (backend.py:csrmm)."""
    from ..sparse import spspmm
    c = spspmm(_as_matrix(gidxA, A_weights), _as_matrix(gidxB, B_weights))
    return _as_graph(c, gidxA, num_vtypes), c.val


def csrsum(gidxs, weights):
    """Element-wise sum of edge-weighted graphs over the same node sets; the result has the union sparsity
    (backend.py:csrsum)."""
    if len(gidxs) != len(weights) or not gidxs:
        raise DGLError("csrsum expects one weight tensor per graph")
    acc = _as_matrix(gidxs[0], weights[0])
    for g, w in zip(gidxs[1:], weights[1:]):
        acc = acc + _as_matrix(g, w)
    acc = acc.coalesce()
    u = _unit(gidxs[0])
    same = getattr(gidxs[0], "ntypes", None)
    num_vtypes = 1 if (same is not None and len(same) == 1) or (same is None and u.num_src == u.num_dst) else 2
    return _as_graph(acc, gidxs[0], num_vtypes), acc.val


def csrmask(gidxA, A_weights, gidxB):
    """Weights of ``A`` read at the non-zero positions of ``B`` (zero where ``A`` has no entry), edge-aligned with
    ``B`` (backend.py:csrmask). Duplicate entries of ``A`` are summed. Differentiable in ``A_weights``."""
    a, b = _unit(gidxA), _unit(gidxB)
    ar, ac = a.coo()[:2]
    br, bc = b.coo()[:2]
    width = max(a.num_dst, b.num_dst, 1)
    akey = ar.long() * width + ac.long()
    bkey = br.long() * width + bc.long()
    uniq, inv = torch.unique(akey, return_inverse=True)
    summed = torch.zeros((len(uniq),) + tuple(A_weights.shape[1:]), dtype=A_weights.dtype, device=A_weights.device)
    summed = summed.index_add(0, inv, A_weights)
    pos = torch.searchsorted(uniq, bkey).clamp(max=max(len(uniq) - 1, 0))
    hit = (uniq[pos] == bkey) if len(uniq) else torch.zeros_like(bkey, dtype=torch.bool)
    out = summed[pos] if len(uniq) else summed.new_zeros((len(bkey),) + tuple(A_weights.shape[1:]))
    return out * hit.view(-1, *([1] * (out.dim() - 1))).to(out.dtype)


def scatter_add(x, idx, m):
    """``out[idx[i]] += x[i]`` into ``m`` rows (backend.py:scatter_add)."""
    out = torch.zeros((m,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    return out.index_add(0, idx.long(), x)


class _Named:
    """The reference names one autograd Function per operator; the functions above are already differentiable, so
    these keep the ``X.apply(...)`` spelling for code written against backend/pytorch/sparse.py."""
    fn = None

    @classmethod
    def apply(cls, *args):
        return cls.fn(*args)


class CSRMM(_Named):
    fn = staticmethod(csrmm)


class CSRSum(_Named):
    fn = staticmethod(csrsum)


class CSRMask(_Named):
    fn = staticmethod(csrmask)


class ScatterAdd(_Named):
    fn = staticmethod(scatter_add)


# ------------------------------------------------------------------------------------------- typed message passing ----
def _relations(g):
    """(relation index, src ntype id, dst ntype id, UnitGraph) for every canonical edge type of ``g``."""
    nid = {nt: i for i, nt in enumerate(g.ntypes)}
    return [(i, nid[s], nid[d], g._graphs[i]) for i, (s, _, d) in enumerate(g.canonical_etypes)]


def gspmm_hetero(g, op, reduce_op, lhs_len, *lhs_and_rhs_tuple):
    """g-SpMM over every relation of a heterograph (backend.py:gspmm_hetero).

    ``lhs_and_rhs_tuple`` is ``lhs_len`` source feature tensors indexed by *source node type id* followed by the edge
    feature tensors indexed by *relation id* (entries may be None). Returns one tensor per node type id: relations
    that end in the same destination type are combined with the reducer (sum / max / min; ``mean`` divides the sum by
    the total in-degree over those relations); types that receive nothing get None."""
    from ..ops import sparse_autograd as SA
    lhs, rhs = lhs_and_rhs_tuple[:lhs_len], lhs_and_rhs_tuple[lhs_len:]
    red = "sum" if reduce_op == "mean" else reduce_op
    out = [None] * len(g.ntypes)
    deg = [None] * len(g.ntypes)
    reached = [None] * len(g.ntypes)
    for rel, s, d, ug in _relations(g):
        u = lhs[s] if (op != "copy_rhs" and s < len(lhs)) else None
        e = rhs[rel] if (op != "copy_lhs" and rel < len(rhs)) else None
        if (op != "copy_rhs" and u is None) or (op != "copy_lhs" and e is None) or ug.num_edges() == 0:
            continue
        part = SA.gspmm(ug, op, red, u.unsqueeze(-1) if u is not None and u.dim() == 1 else u,
                        e.unsqueeze(-1) if e is not None and e.dim() == 1 else e)
        if reduce_op == "mean":
            dd = ug.in_degrees().to(part.dtype)
            deg[d] = dd if deg[d] is None else deg[d] + dd
        has_new = (ug.in_degrees() > 0).view(-1, *([1] * (part.dim() - 1)))
        if out[d] is None:
            out[d], reached[d] = part, has_new
        elif red == "sum":
            out[d] = out[d] + part
        else:
            # rows a relation does not reach hold 0 in the single-relation kernels: compare only where both reached
            pick = torch.maximum if red == "max" else torch.minimum
            both = has_new & reached[d]
            out[d] = torch.where(both, pick(out[d], part), torch.where(has_new, part, out[d]))
            reached[d] = reached[d] | has_new
    if reduce_op == "mean":
        for d, o in enumerate(out):
            if o is not None:
                out[d] = o / deg[d].clamp(min=1).view(-1, *([1] * (o.dim() - 1)))
    return tuple(out)


def gsddmm_hetero(g, op, lhs_len, lhs_target="u", rhs_target="v", *lhs_and_rhs_tuple):
    """g-SDDMM over every relation (backend.py:gsddmm_hetero). Operands are indexed by node type id when their target
    is ``u`` / ``v`` and by relation id when it is ``e``; returns one edge tensor per relation (None when an operand is
    missing)."""
    from ..ops import sparse_autograd as SA
    lhs, rhs = lhs_and_rhs_tuple[:lhs_len], lhs_and_rhs_tuple[lhs_len:]

    def pick(seq, target, rel, s, d):
        i = {"u": s, "v": d, "e": rel}[target]
        return seq[i] if i < len(seq) else None

    out = []
    for rel, s, d, ug in _relations(g):
        a, b = pick(lhs, lhs_target, rel, s, d), pick(rhs, rhs_target, rel, s, d)
        if (op != "copy_rhs" and a is None) or (op != "copy_lhs" and b is None):
            out.append(None)
            continue
        a = a.unsqueeze(-1) if a is not None and a.dim() == 1 else a
        b = b.unsqueeze(-1) if b is not None and b.dim() == 1 else b
        out.append(SA.gsddmm(ug, op, a, b, lhs_target, rhs_target))
    return tuple(out)


def edge_softmax_hetero(g, eids=None, norm_by="dst", *logits):
    """Edge softmax normalised over ALL relations that share the normalising node type (backend.py:
    edge_softmax_hetero): one score tensor per relation in, one per relation out."""
    if eids is not None and not (isinstance(eids, str) and eids == "__ALL__"):
        raise DGLError("edge_softmax_hetero normalises over all edges; edge subsets are not supported")
    rels = _relations(g)
    n_types = len(g.ntypes)
    key = (lambda s, d: d) if norm_by == "dst" else (lambda s, d: s)
    ends = []
    for rel, s, d, ug in rels:
        row, col = ug.coo()[:2]
        ends.append((col if norm_by == "dst" else row).long())
    outs = [None] * len(rels)
    for t in range(n_types):
        members = [i for i, (rel, s, d, ug) in enumerate(rels) if key(s, d) == t and logits[i] is not None]
        if not members:
            continue
        n = g.num_nodes(g.ntypes[t])
        shape_tail = tuple(logits[members[0]].shape[1:])
        top = torch.full((n,) + shape_tail, float("-inf"), dtype=logits[members[0]].dtype, device=logits[members[0]].device)
        for i in members:
            idx = ends[i].view(-1, *([1] * len(shape_tail))).expand_as(logits[i])
            top = top.scatter_reduce(0, idx, logits[i].detach(), reduce="amax", include_self=True)
        total = torch.zeros_like(top)
        exps = {}
        for i in members:
            exps[i] = torch.exp(logits[i] - top[ends[i]])
            total = total.index_add(0, ends[i], exps[i])
        for i in members:
            outs[i] = exps[i] / total[ends[i]]
    return tuple(outs)


# ------------------------------------------------------------------------------------- legacy kernel entry points ----
_TARGET = {0: "u", 1: "v", 2: "e", "src": "u", "dst": "v", "edge": "e", "u": "u", "v": "v", "e": "e"}


def binary_reduce(reducer, binary_op, graph, lhs, rhs, lhs_data, rhs_data, out_size, lhs_map=(None, None),
                  rhs_map=(None, None), out_map=(None, None)):
    """Pre-0.5 kernel interface (backend.py:binary_reduce): ``out[target] = reducer(binary_op(lhs_data[lhs],
    rhs_data[rhs]))`` with operands addressed by ``src`` / ``dst`` / ``edge`` targets. ``reducer='none'`` writes one
    row per edge. Expressed through g-SDDMM (+ a segment reduction onto the destination)."""
    from ..ops import sparse_autograd as SA
    ug = _unit(graph)
    msg = SA.gsddmm(ug, binary_op, lhs_data, rhs_data, _TARGET[lhs], _TARGET[rhs])
    if reducer == "none":
        return msg
    red = {"sum": "sum", "max": "max", "min": "min", "mean": "sum", "prod": None}[reducer]
    if red is None:
        raise DGLError("binary_reduce: the 'prod' reducer is not supported")
    out = SA.gspmm(ug, "copy_rhs", red, None, msg)
    if reducer == "mean":
        out = out / ug.in_degrees().clamp(min=1).to(out.dtype).view(-1, *([1] * (out.dim() - 1)))
    return out[:out_size] if out_size is not None and out_size < out.shape[0] else out


def copy_reduce(reducer, graph, target, in_data, out_size, in_map=(None, None), out_map=(None, None)):
    """Pre-0.5 kernel interface (backend.py:copy_reduce): reduce source (``target`` = src) or edge data onto the
    destination nodes."""
    from ..ops import sparse_autograd as SA
    ug = _unit(graph)
    tgt = _TARGET[target]
    red = "sum" if reducer in ("sum", "mean") else reducer
    if reducer == "none":
        return SA.gsddmm(ug, "copy_lhs", in_data, None, tgt, "v")
    out = (SA.gspmm(ug, "copy_lhs", red, in_data, None) if tgt == "u" else SA.gspmm(ug, "copy_rhs", red, None, in_data))
    if reducer == "mean":
        out = out / ug.in_degrees().clamp(min=1).to(out.dtype).view(-1, *([1] * (out.dim() - 1)))
    return out[:out_size] if out_size is not None and out_size < out.shape[0] else out


# ------------------------------------------------------------------------- names of backend/pytorch/sparse.py ----
class GSpMM_hetero(_Named):  # noqa: N801 - reference spelling
    fn = staticmethod(gspmm_hetero)


class GSDDMM_hetero(_Named):  # noqa: N801
    fn = staticmethod(gsddmm_hetero)


class EdgeSoftmax_hetero(_Named):  # noqa: N801
    fn = staticmethod(edge_softmax_hetero)


class GATHERMM(_Named):
    @staticmethod
    def fn(a, b, idx_a=None, idx_b=None):
        from ..nn.functional_native import gather_mm
        if idx_a is not None:
            a = a[idx_a.long()]
        return gather_mm(a, b, idx_b=idx_b)


class SEGMENTMM(_Named):
    @staticmethod
    def fn(a, b, seglen_a):
        from ..nn.functional_native import segment_mm
        return segment_mm(a, b, seglen_a)


class empty_context:  # noqa: N801
    """Stand-in for an autocast scope when mixed precision is off (backend/pytorch/sparse.py:empty_context)."""

    def __init__(self, *args, **kwargs):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


# What a backward pass has to keep from the forward operands. out = reduce(op(X[u], Y[e])) for g-SpMM,
# out[e] = op(X[.], Y[.]) for g-SDDMM; d/dX needs Y exactly when op is multiplicative, and division also needs X for
# d/dY. Arg-indices are only consulted by max / min reducers.
_MULTIPLICATIVE = ("mul", "div", "dot")


def spmm_cache_X(binary_op, reduce_op, req_grad_X, req_grad_Y):  # noqa: N802, ARG001
    return bool(req_grad_Y and binary_op in _MULTIPLICATIVE)


def spmm_cache_Y(binary_op, reduce_op, req_grad_X, req_grad_Y):  # noqa: N802, ARG001
    return bool((req_grad_X and binary_op in _MULTIPLICATIVE) or (req_grad_Y and binary_op == "div"))


def spmm_cache_argX(binary_op, reduce_op, req_grad_X, req_grad_Y):  # noqa: N802
    return bool((req_grad_X or req_grad_Y) and reduce_op in ("max", "min") and binary_op != "copy_rhs")


def spmm_cache_argY(binary_op, reduce_op, req_grad_X, req_grad_Y):  # noqa: N802
    return bool((req_grad_X or req_grad_Y) and reduce_op in ("max", "min") and binary_op != "copy_lhs")


def sddmm_cache_X(op, req_grad_X, req_grad_Y):  # noqa: N802, ARG001
    return bool(req_grad_Y and op in _MULTIPLICATIVE)


def sddmm_cache_Y(op, req_grad_X, req_grad_Y):  # noqa: N802
    return bool((req_grad_X and op in _MULTIPLICATIVE) or (req_grad_Y and op == "div"))
