"""Tensor-level message-passing vocabulary of ``dgl_b200.nn``.

Every layer of this package is written as a short composition of the functions below, which call the framework's own
sm_100a operators (``ops.gspmm`` / ``gsddmm`` / ``edge_softmax`` / fused kernels) on plain tensors: no node/edge frames
are touched, nothing is stored on the graph, no ``local_scope`` is needed, and - unlike the frame-based formulation of
the reference's layers (python/dgl/nn/pytorch/conv/*.py) - no step synchronises with the host.

Conventions: ``g`` is a homogeneous graph, a bipartite graph or a block (single edge type); ``x_src`` lives on source
nodes, results on destination nodes; edge tensors follow the graph's edge order.
"""
from __future__ import annotations

import torch

from .. import ops
from ..base import DGLError

__all__ = ["src_dst", "gather_sum", "gather_mean", "gather_max", "gather_min", "gather", "edge_reduce", "pair_op",
           "attend", "degrees", "sym_norm_propagate", "require_in_edges", "hop", "edge_softmax_agg", "norm_coefficients",
           "structure_cache"]


def src_dst(g, feat):
    """Split a layer input into (source features, destination features).

    A pair is returned as is; on a block the destination nodes are the leading rows of the source nodes."""
    if isinstance(feat, (tuple, list)):
        return feat[0], feat[1]
    if getattr(g, "is_block", False):
        return feat, feat[: g.number_of_dst_nodes()]
    return feat, feat


def _flat(x):
    """[N, a, b, ...] -> ([N, a*b*...], tail) so the operators see 2-D operands."""
    if x.dim() <= 2:
        return x, None
    return x.reshape(x.shape[0], -1), tuple(x.shape[1:])


def _unflat(y, tail):
    return y if tail is None else y.reshape(y.shape[0], *tail)


def gather(g, x_src, how="sum", w=None):
    """out[v] = reduce_{u -> v} (w[e] * x_src[u]); ``how`` in sum | mean | max | min; ``w`` is [E] or broadcastable
    to the message shape (None = unweighted copy)."""
    if w is None:
        return ops.gspmm(g, "copy_lhs", how, x_src, None)
    if w.dim() < x_src.dim():
        w = w.reshape(w.shape + (1,) * (x_src.dim() - w.dim()))
    return ops.gspmm(g, "mul", how, x_src, w)


def gather_sum(g, x_src, w=None):
    return gather(g, x_src, "sum", w)


def gather_mean(g, x_src, w=None):
    return gather(g, x_src, "mean", w)


def gather_max(g, x_src):
    return gather(g, x_src, "max")


def gather_min(g, x_src):
    return gather(g, x_src, "min")


def edge_reduce(g, e, how="sum"):
    """out[v] = reduce over the in-edges of v of the edge tensor ``e``."""
    return ops.gspmm(g, "copy_rhs", how, None, e)


def pair_op(g, op, x_src, x_dst):
    """Per-edge binary op between the source and the destination endpoint: add | sub | mul | div | dot."""
    return ops.gsddmm(g, op, x_src, x_dst, "u", "v")


def degrees(g, side="in", dtype=torch.float32, floor=1.0):
    """Clamped node degrees as a float column-free vector (``side`` = in | out)."""
    d = g.in_degrees() if side == "in" else g.out_degrees()
    d = d.to(dtype)
    return d.clamp(min=floor) if floor is not None else d


def structure_cache(g):
    """Per-STRUCTURE scratch dict: it lives on the immutable adjacency object (mutating a graph creates a new one), so
    anything derived from the topology alone - degree vectors, normalisation coefficients, the isolated-node verdict -
    is computed once per graph, not once per forward."""
    units = getattr(g, "_graphs", None)
    if units is None or len(units) != 1:
        return {}
    return units[0]._degs


def require_in_edges(g, allow, layer):
    """Raise if a destination node has no in-edge (its output would silently be zero / the softmax undefined).
    The verdict is cached per graph structure, so a training loop pays the reduction + host read ONCE instead of
    every forward (the reference layers re-check, and synchronise, on every call)."""
    if allow:
        return
    cache = structure_cache(g)
    ok = cache.get("no_isolated_dst")
    if ok is None:
        ok = not bool((g.in_degrees() == 0).any()) if g.number_of_dst_nodes() else True
        cache["no_isolated_dst"] = ok
    if not ok:
        raise DGLError(
            f"There are 0-in-degree nodes in the graph, output for those nodes will be invalid. Add self-loops with "
            f"`g = dgl.add_self_loop(g)` or construct {layer} with `allow_zero_in_degree=True` to suppress this check.")


def sym_norm_propagate(g, x_src, mode="both", w=None):
    """Degree-normalised aggregation of GCN-style layers as ONE weighted g-SpMM.

    both : D_in^-1/2 A D_out^-1/2 x      right : D_in^-1 A x      left : A D_out^-1 x      none : A x
    With edge weights ``w`` the structural degrees are still used (weights are expected to be pre-normalised, e.g. by
    ``EdgeWeightNorm``)."""
    return gather_sum(g, x_src, norm_coefficients(g, mode, x_src.dtype, w))


def norm_coefficients(g, mode, dtype, w=None):
    """Per-edge coefficient that folds the degree normalisation of ``sym_norm_propagate`` into ONE weighted g-SpMM:
    both: d_out(u)^-1/2 d_in(v)^-1/2, right: 1/d_in(v), left: 1/d_out(u), in_both: d_in(u)^-1/2 d_in(v)^-1/2 (the
    convention of the spectral filters, which read one degree vector for both ends) - times ``w`` when given. The structural part
    is cached per graph, so full-graph training computes it once; the two N x F row scalings of the textbook
    formulation disappear from every forward and backward."""
    if mode == "none":
        return w
    cache = structure_cache(g)
    key = ("norm_coeff", mode, dtype)
    c = cache.get(key)
    if c is None:
        if mode == "in_both":
            one_src = one_dst = degrees(g, "in", dtype).pow(-0.5)
        else:
            one_src = degrees(g, "out", dtype).pow(-0.5 if mode == "both" else -1.0) if mode in ("both", "left") else None
            one_dst = degrees(g, "in", dtype).pow(-0.5 if mode == "both" else -1.0) if mode in ("both", "right") else None
        if one_src is not None and one_dst is not None:
            c = pair_op(g, "mul", one_src.unsqueeze(-1), one_dst.unsqueeze(-1)).reshape(-1)
        elif one_src is not None:
            c = ops.gsddmm(g, "copy_lhs", one_src.unsqueeze(-1), None, "u", "v").reshape(-1)
        else:
            c = ops.gsddmm(g, "copy_rhs", None, one_dst.unsqueeze(-1), "u", "v").reshape(-1)
        c = c.detach()
        cache[key] = c
    return c if w is None else c * w


def hop(g, x, norm=None, w=None):
    """One propagation step on a homogeneous graph with an optional symmetric pre/post scaling vector ``norm``
    (already shaped [N, 1]): norm * A (norm * x)."""
    if norm is not None:
        x = x * norm
    y = gather_sum(g, x, w)
    return y * norm if norm is not None else y


def attend(g, scores):
    """Softmax of per-edge scores over the in-edges of every destination node."""
    return ops.edge_softmax(g, scores)


def edge_softmax_agg(g, scores, values_src):
    """softmax(scores) over in-edges, then the attention-weighted sum of source values: the GAT aggregation."""
    return gather_sum(g, values_src, attend(g, scores))


# ---------------------------------------------------------------------------------------------- capacity-shaped blocks
def is_capacity_block(graph):
    """True for the fixed-capacity blocks of the whole-step CUDA-graph trainer (dataloading/graphed.py): an ELL neighbour
    table plus a device-side live-row count instead of variable-shaped CSR arrays."""
    return getattr(graph, "_caps", None) is not None and getattr(graph, "_ell_cache", None) is not None


def capacity_aggregate(graph, x_src, reduce="mean"):
    """(rows of the destination nodes, their neighbourhood reduction) on a capacity block: one ELL kernel forward, one
    scatter kernel backward, no size read back (``ops.fused.ell_mean_concat``). ``reduce``: 'mean' or 'sum'."""
    from ..ops.fused import ell_mean_concat
    nbr, cnt, _ = graph._ell_cache
    width = x_src.shape[1]
    both = ell_mean_concat(x_src, nbr, cnt, graph._n_dst_dev)
    own, pooled = both[:, :width], both[:, width:]
    if reduce == "sum":
        pooled = pooled * cnt.clamp(min=0).unsqueeze(1).to(pooled.dtype)
    elif reduce != "mean":
        raise DGLError(f"capacity blocks support 'mean' and 'sum' aggregation, got {reduce!r}")
    return own, pooled


def capacity_degrees(graph, n_src):
    """(in-degree of every destination row, out-degree of every source row INSIDE the block) as float tensors, computed
    on the device from the ELL table (padding rows / slots excluded)."""
    nbr, cnt, fanout = graph._ell_cache
    n_cap = nbr.shape[0]
    live_row = torch.arange(n_cap, device=nbr.device).unsqueeze(1) < graph._n_dst_dev
    live = (torch.arange(fanout, device=nbr.device).unsqueeze(0) < cnt.unsqueeze(1)) & live_row
    indeg = (cnt * live_row.squeeze(1)).to(torch.float32)
    outdeg = torch.zeros(n_src, dtype=torch.float32, device=nbr.device)
    outdeg.index_add_(0, nbr.clamp(min=0).reshape(-1).long(), live.reshape(-1).to(torch.float32))
    return indeg, outdeg


class CapacityRelation:
    """The two sparse layouts the fused GAT kernels need, built ON THE DEVICE from a capacity block's ELL table with
    capacity-sized buffers (nothing is read back, so construction is CUDA-graph capturable):

    * ``csc()``  - destination-major: ``indptr`` = exclusive scan of the live per-row counts, ``indices`` compacted by one
      ``ell_to_csc`` launch (the tail of the buffer beyond the live edge count is never addressed);
    * ``csr()``  - source-major, for the backward: the ELL slots as a COO list in which padding slots are sent to a
      sentinel source row ``n_src`` and sorted away by the radix-sort CSR builder (``structure_build.cu``)."""

    def __init__(self, graph, n_src):
        from .._native import loader
        from ..base import id_code
        from ..ops.cuda_struct import cuda_coo_to_csr
        nbr, cnt, fanout = graph._ell_cache
        dev = nbr.device
        n_cap = nbr.shape[0]
        live_row = torch.arange(n_cap, device=dev) < graph._n_dst_dev
        cnt_live = (cnt * live_row).to(torch.int32)
        ptr64 = torch.zeros(n_cap + 1, dtype=torch.int64, device=dev)
        torch.cumsum(cnt_live, 0, out=ptr64[1:])
        idx = torch.zeros(n_cap * fanout, dtype=torch.int32, device=dev)
        loader.cuda_lib().fn("dglb200_ell_to_csc", "ipppplipppp")(
            id_code(torch.int32), nbr.data_ptr(), 0, cnt_live.data_ptr(), ptr64.data_ptr(), n_cap, fanout, idx.data_ptr(), 0, 0,
            loader.stream_ptr())
        self._csc = (ptr64.to(torch.int32), idx, None)
        live = torch.arange(fanout, device=dev).unsqueeze(0) < cnt_live.unsqueeze(1)
        src = torch.where(live, nbr, torch.full_like(nbr, n_src)).reshape(-1)
        dst = torch.arange(n_cap, dtype=torch.int32, device=dev).repeat_interleave(fanout)
        rp, ri, _ = cuda_coo_to_csr(src, dst, n_src + 1)           # row n_src collects the padding slots
        self._csr = (rp[: n_src + 1].contiguous(), ri, None)
        self.num_src, self.num_dst = n_src, n_cap

    def csc(self):
        return self._csc

    def csr(self):
        return self._csr
