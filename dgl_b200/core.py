"""Message-passing dispatcher.

Parity: /root/reference/python/dgl/core.py:372-425 (``message_passing``),
:311-369 (``invoke_gspmm``), :273-309 (``invoke_gsddmm``), :99-174 (degree
bucketing for UDF reducers), and /root/reference/python/dgl/udf.py
(EdgeBatch/NodeBatch). Builtin (message, reduce) pairs lower to ONE fused g-SpMM
kernel; builtin messages alone lower to g-SDDMM; anything else takes the UDF
path (edge UDF -> materialised messages -> builtin or degree-bucketed reduce).
"""
from __future__ import annotations

import torch

from . import function as fn
from .base import DGLError
from .ops import sparse_autograd as SA


class EdgeBatch:
    """Argument of an edge UDF (reference: python/dgl/udf.py:EdgeBatch)."""

    def __init__(self, graph, eid, etype, src_data, edge_data, dst_data):
        self._graph = graph
        self._eid = eid
        self._etype = etype
        self._src_data = src_data
        self._edge_data = edge_data
        self._dst_data = dst_data

    @property
    def src(self):
        return self._src_data

    @property
    def dst(self):
        return self._dst_data

    @property
    def data(self):
        return self._edge_data

    def edges(self):
        u, v = self._graph.find_edges(self._eid, etype=self._etype)
        return u, v, self._eid

    def batch_size(self):
        return len(self._eid)

    def __len__(self):
        return self.batch_size()

    @property
    def canonical_etype(self):
        return self._etype


class NodeBatch:
    """Argument of a node UDF (reference: python/dgl/udf.py:NodeBatch)."""

    def __init__(self, graph, nodes, ntype, data, msgs=None):
        self._graph = graph
        self._nodes = nodes
        self._ntype = ntype
        self._data = data
        self._msgs = msgs

    @property
    def data(self):
        return self._data

    @property
    def mailbox(self):
        return self._msgs

    def nodes(self):
        return self._nodes

    def batch_size(self):
        return len(self._nodes)

    def __len__(self):
        return self.batch_size()

    @property
    def ntype(self):
        return self._ntype


def data_dict_to_list(graph, data_dict, func, target):
    """Per-type list of the operand a builtin message function reads (reference core.py:207).

    ``target`` is 'u', 'v' or 'e'. For node targets the list is indexed by node-type id, for edges by edge-type id;
    types that do not take part stay ``None``. With a single source type ``data_dict`` may be a bare tensor."""
    if isinstance(func, fn.BinaryMessageFunction):
        if target in ("u", "v"):
            out = [None] * len(graph.ntypes)
            if not isinstance(data_dict, dict):
                s, _, d = graph.canonical_etypes[0]
                out[graph.get_ntype_id(s if target == "u" else d)] = data_dict
            else:
                for s, _, d in graph.canonical_etypes:
                    t = s if target == "u" else d
                    out[graph.get_ntype_id(t)] = data_dict[t]
            return out
        out = [None] * len(graph.canonical_etypes)
        for rel in graph.canonical_etypes:
            out[graph.get_etype_id(rel)] = data_dict[rel]
        return out
    if target == "u":  # copy_u
        out = [None] * len(graph.ntypes)
        if not isinstance(data_dict, dict):
            out[graph.get_ntype_id(graph.canonical_etypes[0][0])] = data_dict
        else:
            for s, _, _ in graph.canonical_etypes:
                out[graph.get_ntype_id(s)] = data_dict[s]
        return out
    out = [None] * len(graph.canonical_etypes)  # copy_e
    for rel in graph.canonical_etypes:
        out[graph.get_etype_id(rel)] = data_dict[rel]
    return out


def is_builtin(func):
    return isinstance(func, fn.BuiltinFunction)


class _RelView:
    """Everything message passing needs about one relation."""

    def __init__(self, graph, etid):
        self.graph = graph
        self.etid = etid
        self.gidx = graph._graphs[etid]
        self.canonical_etype = graph.canonical_etypes[etid]
        self.src_frame = graph._node_frames[graph._src_ntid(etid)]
        self.dst_frame = graph._node_frames[graph._dst_ntid(etid)]
        self.edge_frame = graph._edge_frames[etid]


def _get_field(view, target, field):
    frame = {"u": view.src_frame, "v": view.dst_frame, "e": view.edge_frame}[target]
    if field not in frame:
        raise DGLError(f"feature '{field}' not found on target '{target}' of relation {view.canonical_etype}")
    return frame[field]


def invoke_gsddmm(view, func):
    """builtin message function -> g-SDDMM (dict out_field -> tensor)."""
    if isinstance(func, fn.BinaryMessageFunction):
        x = _get_field(view, func.lhs, func.lhs_field)
        y = _get_field(view, func.rhs, func.rhs_field)
        z = SA.gsddmm(view.gidx, func.binary_op, x, y, func.lhs, func.rhs)
    else:
        x = _get_field(view, func.target, func.in_field)
        if func.target == "e":
            z = x
        else:
            z = SA.gsddmm(view.gidx, "copy_lhs", x, None, func.target, "e")
    return {func.out_field: z}


def _zero_deg_mean(view, z):
    deg = view.gidx.in_degrees().clamp(min=1).to(z.dtype)
    return z / deg.view(-1, *([1] * (z.dim() - 1)))


def invoke_gspmm(view, mfunc, rfunc, *, edata=None):
    """builtin message + builtin reduce -> one g-SpMM (reference core.py:311-369)."""
    red = rfunc.name
    kred = "sum" if red == "mean" else red
    if isinstance(mfunc, fn.BinaryMessageFunction):
        if mfunc.lhs == "v" or mfunc.rhs == "v" or mfunc.binary_op == "dot" or (mfunc.lhs == "e" and mfunc.rhs == "e"):
            # not expressible as SpMM: materialise messages then reduce
            msg = invoke_gsddmm(view, mfunc)[mfunc.out_field]
            z = SA.gspmm(view.gidx, "copy_rhs", kred, None, msg)
        else:
            x = _get_field(view, mfunc.lhs, mfunc.lhs_field)
            y = _get_field(view, mfunc.rhs, mfunc.rhs_field)
            op = mfunc.binary_op
            if mfunc.lhs == "e":  # e op u  ->  u op' e
                if op in ("add", "mul"):
                    x, y = y, x
                elif op == "sub":  # e - u = (-u) + e
                    x, y, op = -y, x, "add"
                else:  # e / u = (1/u) * e
                    x, y, op = 1.0 / y, x, "mul"
            z = SA.gspmm(view.gidx, op, kred, x, y)
    else:
        x = _get_field(view, mfunc.target, mfunc.in_field) if edata is None else edata
        if mfunc.target == "u":
            z = SA.gspmm(view.gidx, "copy_lhs", kred, x, None)
        else:
            z = SA.gspmm(view.gidx, "copy_rhs", kred, None, x)
    if red == "mean":
        z = _zero_deg_mean(view, z)
    return {rfunc.out_field: z}


def invoke_edge_udf(view, eid, func, orig_eid=None):
    graph = view.graph
    etype = view.canonical_etype
    if isinstance(eid, slice) or eid is None:
        u, v = view.gidx.coo()
        eid_t = torch.arange(u.shape[0], device=u.device, dtype=view.gidx.idtype)
        edata = view.edge_frame
        all_edges = True
    else:
        u, v = graph.find_edges(eid, etype=etype)
        eid_t = eid
        edata = view.edge_frame.subframe(eid)
        all_edges = False
    srcdata = view.src_frame.subframe(u)
    dstdata = view.dst_frame.subframe(v)
    ebatch = EdgeBatch(graph, eid_t if orig_eid is None else orig_eid, etype, srcdata, edata, dstdata)
    out = func(ebatch)
    if out is None:
        return {}
    if not isinstance(out, dict):
        raise DGLError("edge UDF must return a dict of tensors")
    return out


def _bucketing(degs):
    """unique in-degrees and node buckets (reference core.py:99-174)."""
    unique_degs, inverse = torch.unique(degs, return_inverse=True)
    order = torch.argsort(inverse, stable=True)
    counts = torch.bincount(inverse, minlength=unique_degs.shape[0])
    splits = torch.split(order, counts.tolist())
    return unique_degs.tolist(), splits


def invoke_udf_reduce(view, func, msgdata, *, orig_nid=None):
    """Degree-bucketed reduce: dst nodes with equal in-degree are batched so the UDF sees a
    dense mailbox [n, deg, ...]."""
    graph = view.graph
    gidx = view.gidx
    indptr, _, eids = gidx.csc()
    degs = (indptr[1:] - indptr[:-1]).long()
    nodes = torch.arange(gidx.num_dst, device=indptr.device, dtype=gidx.idtype)
    unique_degs, buckets = _bucketing(degs)
    bkt_rsts = []
    bkt_nodes = []
    dst_ntype = view.canonical_etype[2]
    for deg, node_bkt in zip(unique_degs, buckets):
        if deg == 0:
            continue
        bkt_nodes.append(node_bkt)
        ndata_bkt = view.dst_frame.subframe(node_bkt)
        starts = indptr[:-1][node_bkt].long()
        pos = (starts.view(-1, 1) + torch.arange(deg, device=starts.device).view(1, -1)).reshape(-1)
        eid_bkt = eids[pos].long() if eids is not None else pos
        maildata = {}
        for k, v in msgdata.items():
            m = v[eid_bkt]
            maildata[k] = m.view(len(node_bkt), deg, *m.shape[1:])
        nbatch = NodeBatch(graph, nodes[node_bkt] if orig_nid is None else orig_nid[node_bkt], dst_ntype, ndata_bkt,
                           msgs=maildata)
        bkt_rsts.append(func(nbatch))
    retf = {}
    if not bkt_rsts:
        return retf
    merged_nodes = torch.cat(bkt_nodes, 0)
    for k in bkt_rsts[0].keys():
        vals = torch.cat([r[k] for r in bkt_rsts], 0)
        full = torch.zeros((gidx.num_dst,) + tuple(vals.shape[1:]), dtype=vals.dtype, device=vals.device)
        full = full.index_copy(0, merged_nodes, vals)
        retf[k] = full
    return retf


def invoke_node_udf(graph, nid, ntype, func, *, ndata=None, orig_nid=None):
    ntid = graph.get_ntype_id(ntype)
    if ndata is None:
        if isinstance(nid, slice) or nid is None:
            ndata = graph._node_frames[ntid]
            nid = torch.arange(graph.num_nodes(ntype), device=graph.device, dtype=graph.idtype)
        else:
            ndata = graph._node_frames[ntid].subframe(nid)
    nbatch = NodeBatch(graph, nid if orig_nid is None else orig_nid, ntype, ndata)
    return func(nbatch)


def message_passing(view, mfunc, rfunc, afunc):
    """One round of message passing on a relation; returns dict of new dst features."""
    if is_builtin(mfunc) and is_builtin(rfunc) and (
            isinstance(mfunc, fn.CopyMessageFunction) or mfunc.out_field == rfunc.msg_field):
        if isinstance(mfunc, fn.CopyMessageFunction) and mfunc.out_field != rfunc.msg_field:
            raise DGLError(f"reduce function reads message field '{rfunc.msg_field}' but message function "
                           f"writes '{mfunc.out_field}'")
        ndata = invoke_gspmm(view, mfunc, rfunc)
    else:
        if is_builtin(mfunc):
            msgdata = invoke_gsddmm(view, mfunc)
        else:
            msgdata = invoke_edge_udf(view, None, mfunc)
        if is_builtin(rfunc):
            if rfunc.msg_field not in msgdata:
                raise DGLError(f"message field '{rfunc.msg_field}' was not produced by the message function")
            msg = msgdata[rfunc.msg_field]
            ndata = invoke_gspmm(view, fn.copy_e(rfunc.msg_field, rfunc.msg_field), rfunc, edata=msg)
        else:
            ndata = invoke_udf_reduce(view, rfunc, msgdata)
    if afunc is not None:
        for k, v in view.dst_frame.items():
            if k not in ndata:
                ndata[k] = v
        dst_ntype = view.canonical_etype[2]
        from .frame import Frame
        out = invoke_node_udf(view.graph, None, dst_ntype, afunc, ndata=Frame(ndata, num_rows=view.gidx.num_dst))
        ndata.update(out or {})
    return ndata
