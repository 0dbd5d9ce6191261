"""Relational / heterogeneous convolutions: RelGraphConv, HeteroGraphConv, HGTConv.

API parity: /root/reference/python/dgl/nn/pytorch/conv/relgraphconv.py, hetero.py, hgtconv.py (constructor
signatures and parameter names are kept so checkpoints interchange). Execution is planned for the grouped tcgen05 GEMMs
(csrc/cuda/grouped_mm_tc.cu) and the g-SpMM of this framework instead of per-edge Python UDFs:

* ``RelGraphConv``  - basis regulariser: aggregate FIRST (B weighted g-SpMMs over the whole edge set + ONE dense GEMM), no
  [E, out] message tensor; otherwise typed messages from ONE grouped GEMM whose A rows are gathered inside the kernel
  (source features are never expanded to [E, in] when the edges are presorted by type) + one copy_e-sum g-SpMM;
* ``HeteroGraphConv`` - generic container; when every relation carries a same-shape ``GraphConv`` the R weight GEMMs run
  as ONE grouped launch over the concatenated per-relation aggregates;
* ``HGTConv``       - all heads and relation types of the attention / message transforms go through ONE grouped GEMM
  each (combined (head, etype) index) instead of 2 x heads typed-linear calls inside an edge UDF.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import ops
from ...base import DGLError
from ..._graph_kernels import keyed_gather
from .. import _mp
from ..linear import TypedLinear


# =====================================================================================================================
#  RelGraphConv
# =====================================================================================================================
class RelGraphConv(nn.Module):
    r"""R-GCN layer ``h_i' = act( sum_r sum_{j in N_r(i)} e_ji W_r h_j + W_0 h_i + b )``."""

    def __init__(self, in_feat, out_feat, num_rels, regularizer=None, num_bases=None, bias=True, activation=None,
                 self_loop=True, dropout=0.0, layer_norm=False):
        super().__init__()
        if regularizer is not None and num_bases is None:
            num_bases = num_rels
        self.linear_r = TypedLinear(in_feat, out_feat, num_rels, regularizer, num_bases)
        self.bias, self.activation, self.self_loop, self.layer_norm = bias, activation, self_loop, layer_norm
        if bias:
            self.h_bias = nn.Parameter(torch.zeros(out_feat))
        if layer_norm:
            self.layer_norm_weight = nn.LayerNorm(out_feat, elementwise_affine=True)
        if self_loop:
            self.loop_weight = nn.Parameter(torch.empty(in_feat, out_feat))
        self.dropout = nn.Dropout(dropout)
        self._init_own()

    def _init_own(self):
        if self.bias:
            nn.init.zeros_(self.h_bias)
        if self.self_loop:
            nn.init.xavier_uniform_(self.loop_weight, gain=nn.init.calculate_gain("relu"))

    def reset_parameters(self):
        self.linear_r.reset_parameters()
        self._init_own()
        if self.layer_norm:
            self.layer_norm_weight.reset_parameters()

    # ---- plan A: aggregate first (basis regulariser)
    def _via_bases(self, g, feat, etypes, norm):
        """W_r = sum_b c_rb V_b  =>  sum_e W_{r(e)} h_src(e) = sum_b ( sum_e c_{r(e),b} h_src(e) ) V_b :
        one weighted g-SpMM per basis (edge weight = the coefficient of the edge's relation, times ``norm``) and ONE
        GEMM [N, B*in] @ [B*in, out]. The coefficient gradient is a shared-memory keyed sum, not an index_add."""
        lin = self.linear_r
        coeff_e = keyed_gather(lin.coeff, etypes)                                  # [E, B]
        if norm is not None:
            coeff_e = coeff_e * norm.reshape(-1, 1)
        per_basis = [_mp.gather_sum(g, feat, coeff_e[:, b].contiguous()) for b in range(lin.num_bases)]
        return torch.cat(per_basis, dim=1) @ lin.W.reshape(lin.num_bases * lin.in_size, lin.out_size)

    # ---- plan B: typed messages
    def _via_messages(self, g, feat, etypes, norm, presorted):
        lin = self.linear_r
        src_of_edge = g.edges()[0].long()
        if presorted and lin.regularizer is None and _grouped_gather_ok(feat, lin.W):
            # the grouped GEMM reads h[src(e)] itself: no [E, in] expansion of the source features
            from ..functional_native import _SegmentMMTC
            msg = _SegmentMMTC.apply(feat, lin.W, torch.bincount(etypes.long(), minlength=lin.num_types), src_of_edge)
        else:
            msg = lin(feat[src_of_edge], etypes, presorted)
        if norm is not None:
            msg = msg * norm.reshape(-1, 1)
        return _mp.edge_reduce(g, msg, "sum")

    def _aggregate_first_pays(self, g, feat):
        lin = self.linear_r
        if lin.regularizer != "basis" or feat.dim() != 2 or not feat.is_floating_point():
            return False
        # B SpMMs over [*, in] + a [N, B*in] GEMM  versus  an [E, out] message tensor
        return lin.num_bases * lin.in_size * g.num_dst_nodes() <= 4 * lin.out_size * max(g.num_edges(), 1)

    def forward(self, g, feat, etypes, norm=None, *, presorted=False):
        if self._aggregate_first_pays(g, feat):
            h = self._via_bases(g, feat, etypes, norm)
        else:
            h = self._via_messages(g, feat, etypes, norm, presorted)
        if self.layer_norm:
            h = self.layer_norm_weight(h)
        if self.bias:
            h = h + self.h_bias
        if self.self_loop:
            h = h + feat[: g.num_dst_nodes()] @ self.loop_weight
        if self.activation:
            h = self.activation(h)
        return self.dropout(h)


def _grouped_gather_ok(x, w):
    from ..functional_native import _tc_ok
    return x.dim() == 2 and _tc_ok(x, w) and not torch.is_autocast_enabled()


# =====================================================================================================================
#  HeteroGraphConv
# =====================================================================================================================
def _stack_reduce(op):
    def run(tensors, dsttype):  # noqa: ARG001 - signature of the reference's cross-type reducers
        if not tensors:
            return None
        return op(torch.stack(tensors, dim=0))
    return run


_CROSS_REDUCERS = {
    "sum": _stack_reduce(lambda t: t.sum(0)),
    "mean": _stack_reduce(lambda t: t.mean(0)),
    "max": _stack_reduce(lambda t: t.amax(0)),
    "min": _stack_reduce(lambda t: t.amin(0)),
    "stack": lambda tensors, dsttype: torch.stack(tensors, dim=1) if tensors else None,
}


def get_aggregate_fn(agg):
    """Cross-relation reducer ``f(list_of_tensors, dsttype)`` by name: sum | max | min | mean | stack."""
    try:
        return _CROSS_REDUCERS[agg]
    except KeyError:
        raise DGLError('Invalid cross type aggregator. Must be one of "sum", "max", "min", "mean" or "stack". '
                       f"But got {agg!r}") from None


class HeteroGraphConv(nn.Module):
    r"""One sub-module per relation; the per-relation results that land on the same destination type are reduced by
    ``aggregate`` (name or callable ``f(tensors, dsttype)``)."""

    def __init__(self, mods, aggregate="sum"):
        super().__init__()
        self.mod_dict = mods
        self.mods = nn.ModuleDict({str(k): m for k, m in mods.items()})
        for m in self.mods.values():
            allow = getattr(m, "set_allow_zero_in_degree", None)
            if callable(allow):
                allow(True)                       # a relation slice routinely has destination nodes without in-edges
        self.agg_fn = get_aggregate_fn(aggregate) if isinstance(aggregate, str) else aggregate
        self._agg_name = aggregate if isinstance(aggregate, str) else None

    def _get_module(self, etype):
        if etype in self.mod_dict:
            return self.mod_dict[etype]
        if isinstance(etype, tuple) and etype[1] in self.mod_dict:
            return self.mod_dict[etype[1]]
        raise KeyError("Cannot find module with edge type %s" % (etype,))

    @staticmethod
    def _split_inputs(g, inputs):
        """(source feature dict, destination feature dict) for a pair, a block or a plain heterograph."""
        if isinstance(inputs, tuple):
            return inputs
        if g.is_block:
            return inputs, {nt: x[: g.number_of_dst_nodes(nt)] for nt, x in inputs.items() if nt in g.dsttypes}
        return inputs, inputs

    def _grouped_graphconv(self, g, rels, src_in, dst_in):
        """All relations carry a GraphConv of one shape that aggregates first: run the R propagations, then ONE grouped
        GEMM (tcgen05 segment_mm over the concatenated aggregates) instead of R separate weight GEMMs."""
        from .graphconv import GraphConv
        convs = [self._get_module(r) for r in rels]
        c0 = convs[0]
        same = all(isinstance(c, GraphConv) and c.weight is not None and c._in_feats == c0._in_feats
                   and c._out_feats == c0._out_feats and c._in_feats <= c._out_feats for c in convs)
        if not same or len(rels) < 2 or self._agg_name not in ("sum", "mean", "max", "min", "stack"):
            return None
        if any(src_in[r[0]].dim() != 2 for r in rels):
            return None
        aggs = []
        for rel, conv in zip(rels, convs):
            sub = g[rel]
            coeff = _mp.norm_coefficients(sub, conv._norm, src_in[rel[0]].dtype)
            aggs.append(_mp.gather_sum(sub, src_in[rel[0]], coeff))
        rows = torch.tensor([a.shape[0] for a in aggs])
        y = ops.segment_mm(torch.cat(aggs, 0), torch.stack([c.weight for c in convs]), rows)
        outs, off = {}, 0
        for rel, conv, n in zip(rels, convs, rows.tolist()):
            h = y[off:off + n]
            off += n
            if conv.bias is not None:
                h = h + conv.bias
            if conv._activation is not None:
                h = conv._activation(h)
            outs.setdefault(rel[2], []).append(h)
        return outs

    def forward(self, g, inputs, mod_args=None, mod_kwargs=None):
        mod_args, mod_kwargs = mod_args or {}, mod_kwargs or {}
        src_in, dst_in = self._split_inputs(g, inputs)
        rels = [r for r in g.canonical_etypes if r[0] in src_in and r[2] in dst_in]
        per_dst = None
        if not mod_args and not mod_kwargs and rels:
            per_dst = self._grouped_graphconv(g, rels, src_in, dst_in)
        if per_dst is None:
            per_dst = {}
            for rel in rels:
                stype, etype, dtype = rel
                out = self._get_module(rel)(g[rel], (src_in[stype], dst_in[dtype]), *mod_args.get(etype, ()),
                                            **mod_kwargs.get(etype, {}))
                per_dst.setdefault(dtype, []).append(out)
        return {nt: self.agg_fn(outs, nt) for nt, outs in per_dst.items() if outs}


# =====================================================================================================================
#  HGTConv
# =====================================================================================================================
class HGTConv(nn.Module):
    r"""Heterogeneous graph transformer layer on a homogenised graph (node / edge type id vectors)."""

    def __init__(self, in_size, head_size, num_heads, num_ntypes, num_etypes, dropout=0.2, use_norm=False):
        super().__init__()
        self.in_size, self.head_size, self.num_heads = in_size, head_size, num_heads
        self.sqrt_d = math.sqrt(head_size)
        self.use_norm = use_norm
        width = head_size * num_heads
        self.linear_k = TypedLinear(in_size, width, num_ntypes)
        self.linear_q = TypedLinear(in_size, width, num_ntypes)
        self.linear_v = TypedLinear(in_size, width, num_ntypes)
        self.linear_a = TypedLinear(width, width, num_ntypes)
        self.relation_pri = nn.ParameterList([nn.Parameter(torch.ones(num_etypes)) for _ in range(num_heads)])
        self.relation_att = nn.ModuleList([TypedLinear(head_size, head_size, num_etypes) for _ in range(num_heads)])
        self.relation_msg = nn.ModuleList([TypedLinear(head_size, head_size, num_etypes) for _ in range(num_heads)])
        self.skip = nn.Parameter(torch.ones(num_ntypes))
        self.drop = nn.Dropout(dropout)
        if use_norm:
            self.norm = nn.LayerNorm(width)
        if in_size != width:
            self.residual_w = nn.Parameter(nn.init.xavier_uniform_(torch.empty(in_size, width)))

    def _per_edge_transform(self, per_head_mods, x_edge, etype):
        """x_edge [E, H, d] -> [E, H, d] with the matrix of (head h, relation etype[e]): ONE grouped GEMM over the
        combined index h * T + t (the reference loops over heads inside an edge UDF)."""
        H, d = self.num_heads, self.head_size
        W = torch.stack([m.get_weight() for m in per_head_mods], dim=0)            # [H, T, d, d]
        T = W.shape[1]
        combo = (torch.arange(H, device=etype.device).unsqueeze(0) * T + etype.long().unsqueeze(1)).reshape(-1)
        y = ops.gather_mm(x_edge.reshape(-1, d), W.reshape(H * T, d, d), idx_b=combo)
        return y.view(-1, H, d)

    def forward(self, g, x, ntype, etype, *, presorted=False):
        H, d = self.num_heads, self.head_size
        x_src, x_dst = _mp.src_dst(g, x)
        nt_src, nt_dst = (ntype, ntype[: g.num_dst_nodes()]) if g.is_block else (ntype, ntype)
        k = self.linear_k(x_src, nt_src, presorted)
        v = self.linear_v(x_src, nt_src, presorted)
        q = self.linear_q(x_dst, nt_dst, presorted).view(-1, H, d)
        # per-edge keys / values of the source endpoint, transformed by the (head, relation) matrices
        k_e = self._per_edge_transform(self.relation_att, ops.copy_u(g, k).view(-1, H, d), etype)
        m_e = self._per_edge_transform(self.relation_msg, ops.copy_u(g, v).view(-1, H, d), etype)
        prior = torch.stack(list(self.relation_pri), dim=1)[etype.long()]           # [E, H]
        score = ops.gsddmm(g, "dot", k_e, q, "e", "v").reshape(-1, H) * prior / self.sqrt_d
        attn = _mp.attend(g, score)
        h = _mp.edge_reduce(g, m_e * attn.unsqueeze(-1), "sum").reshape(-1, H * d)
        h = self.drop(self.linear_a(h, nt_dst, presorted))
        gate = torch.sigmoid(self.skip[nt_dst.long()]).unsqueeze(-1)
        skip_in = x_dst if x_dst.shape == h.shape else x_dst @ self.residual_w
        h = h * gate + skip_in * (1 - gate)
        return self.norm(h) if self.use_norm else h
