"""Kernel dispatch layer: shape/broadcast inference + native call (CPU OpenMP
library or sm_100a CUDA library), no autograd here.

Parity: /root/reference/python/dgl/_sparse_ops.py:156-266 (_gspmm), :482-566
(_gsddmm), :641-800 (segment ops) and /root/reference/src/array/kernel.cc:20-248.
Operand broadcasting is encoded as ``idx = (f // div) % mod`` per operand; shapes
that cannot be encoded that way are expanded to the full output shape first.
"""
from __future__ import annotations

import math

import torch

from ._native import loader
from .base import DGLError, dtype_code, id_code

OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "copy_lhs": 4, "copy_rhs": 5, "dot": 6}
REDS = {"sum": 0, "max": 1, "min": 2}
TARGETS = {"u": 0, "e": 1, "v": 2, 0: 0, 1: 1, 2: 2}


def _prod(xs):
    return int(math.prod(xs)) if len(xs) else 1


def infer_broadcast_shape(op, shp1, shp2):
    """Trailing-feature-shape broadcasting, NumPy rules
    (reference: python/dgl/_sparse_ops.py:10-60)."""
    pad = max(len(shp1), len(shp2))
    s1 = (1,) * (pad - len(shp1)) + tuple(shp1)
    s2 = (1,) * (pad - len(shp2)) + tuple(shp2)
    if op == "dot" and s1[-1] != s2[-1]:
        raise DGLError(f"dot requires equal reduce dims, got {shp1} and {shp2}")
    for a, b in zip(s1, s2):
        if a != b and a != 1 and b != 1:
            raise DGLError(f"feature shapes {shp1} and {shp2} are not broadcastable")
    out = tuple(max(a, b) for a, b in zip(s1, s2))
    if op == "dot":
        out = out[:-1] + (1,)
    return out


def _div_mod(shape, out_shape):
    """Return (div, mod) with operand_idx = (f // div) % mod, or None if the
    operand's non-broadcast dims are not one contiguous run."""
    pad = len(out_shape) - len(shape)
    s = (1,) * pad + tuple(shape)
    live = [i for i, (a, o) in enumerate(zip(s, out_shape)) if o > 1 and a == o]
    dead = [i for i, (a, o) in enumerate(zip(s, out_shape)) if o > 1 and a == 1]
    if not live:
        return 1, 1
    a, b = live[0], live[-1] + 1
    if any(a < d < b for d in dead):
        return None
    return _prod(out_shape[b:]), _prod(out_shape[a:b])


def _prep_operands(op, lhs, rhs):
    """Flatten trailing dims and compute broadcast descriptors.
    Returns (lhs2d, rhs2d, out_trailing_shape, F, rsz, (ldiv, lmod, rdiv, rmod))."""
    use_l = op != "copy_rhs"
    use_r = op != "copy_lhs"
    if not (use_l and use_r):
        return _prep_single(op, lhs, rhs)
    ls = tuple(lhs.shape[1:]) if use_l else ()
    rs = tuple(rhs.shape[1:]) if use_r else ()
    if use_l and use_r:
        out_shape = infer_broadcast_shape(op, ls, rs)
    else:
        out_shape = ls if use_l else rs
    rsz = 1
    if op == "dot":
        rsz = max(ls[-1] if ls else 1, rs[-1] if rs else 1)
        full = out_shape[:-1]
        ls_b, rs_b = ls[:-1], rs[:-1]
    else:
        full = out_shape
        ls_b, rs_b = ls, rs
    F = _prod(full)
    ld = _div_mod(ls_b, full) if use_l else (1, 1)
    rd = _div_mod(rs_b, full) if use_r else (1, 1)
    if use_l and ld is None:
        lhs = lhs.expand(lhs.shape[0], *full, *( (rsz,) if op == "dot" else ()))
        ld = (1, F)
    if use_r and rd is None:
        rhs = rhs.expand(rhs.shape[0], *full, *((rsz,) if op == "dot" else ()))
        rd = (1, F)
    if use_l:
        lhs = lhs.reshape(lhs.shape[0], -1).contiguous()
    if use_r:
        rhs = rhs.reshape(rhs.shape[0], -1).contiguous()
    return lhs, rhs, out_shape, F, rsz, (ld[0], ld[1], rd[0], rd[1])


def _dev_lib(t):
    if t.is_cuda:
        return loader.cuda_lib(), True
    return loader.cpu_lib(), False


def _p(t):
    return 0 if t is None else t.data_ptr()


SPLIT_DEG = 128  # rows longer than this are split into edge-balanced work items


def _split_meta(indptr):
    """Per-CSC load-balancing metadata (cached on the indptr tensor): long rows and their work items."""
    meta = getattr(indptr, "_dglb_split_meta", None)
    if meta is None:
        deg = (indptr[1:] - indptr[:-1]).long()
        long_rows = torch.nonzero(deg > SPLIT_DEG).squeeze(1)
        if long_rows.numel() == 0:
            meta = {"n_long": 0}
        else:
            ldeg = deg[long_rows]
            nchunks = (ldeg + SPLIT_DEG - 1) // SPLIT_DEG
            ord_ = torch.repeat_interleave(torch.arange(long_rows.shape[0], device=indptr.device), nchunks)
            first = torch.cumsum(nchunks, 0) - nchunks
            k = torch.arange(int(nchunks.sum()), device=indptr.device) - first[ord_]
            start = indptr[long_rows].long()[ord_] + k * SPLIT_DEG
            length = torch.minimum(torch.full_like(k, SPLIT_DEG), ldeg[ord_] - k * SPLIT_DEG)
            meta = {"n_long": int(long_rows.shape[0]), "long_rows": long_rows.contiguous(),
                    "ord": ord_.to(torch.int32).contiguous(), "start": start.contiguous(),
                    "len": length.to(torch.int32).contiguous(), "items": int(ord_.shape[0])}
        try:
            indptr._dglb_split_meta = meta
        except Exception:  # noqa: BLE001
            pass
    return meta


def _gspmm_balanced(csr, ufeat2d, ew, num_dst, F, heads=1, out=None):
    indptr, indices, eids = csr
    from .global_config import is_deterministic
    # deterministic mode (USE_DETERMINISTIC_ALG=1 / dgl.global_config.set_deterministic): long rows are NOT split into
    # work items that meet in fp32 red.add scratch, so every output element is reduced by one sub-warp in a fixed order
    meta = {"n_long": 0} if is_deterministic() else _split_meta(indptr)
    if ew is not None and eids is not None:
        # bring the per-edge scalars into CSC order once (sequential reads in the kernel instead of one random 32-byte
        # sector per edge). Reading w[eid] inside the kernel was measured and is slower even for 512-byte rows
        # (reddit F=128 fp32: 2.08 ms vs 1.68 ms with this pass).
        from .ops.cuda_struct import gather_rows
        ew = gather_rows(ew, eids)
        eids = None
    if out is None:
        out = torch.empty((num_dst, F), dtype=ufeat2d.dtype, device=ufeat2d.device)
    lib = loader.cuda_lib()
    if meta["n_long"]:
        scratch = torch.empty((meta["n_long"], F), dtype=torch.float32, device=ufeat2d.device)
        args = (_p(meta["ord"]), _p(meta["start"]), _p(meta["len"]), meta["items"], _p(meta["long_rows"]), meta["n_long"],
                _p(scratch))
    else:
        args = (0, 0, 0, 0, 0, 0, 0)
    lib.fn("dglb200_cuda_gspmm_balanced", "iiippplppp" + "ll" + "ppp" + "l" + "p" + "l" + "pp")(
        (max(1, int(heads)) if ew is not None else 0), dtype_code(ufeat2d.dtype), id_code(indptr.dtype), _p(indptr), _p(indices), _p(eids),
        num_dst, _p(ufeat2d), _p(ew), _p(out), F, SPLIT_DEG if meta["n_long"] else (1 << 40), *args, loader.stream_ptr())
    return out


def gspmm(csr, op, reduce_op, ufeat, efeat, num_dst):
    """out[v] = reduce_{(u,e,v)} op(ufeat[u], efeat[e]).

    ``csr`` = (indptr, indices, eids|None) indexed by *destination* rows (CSC of the
    graph). Returns (out, arg_u, arg_e); args are None for ``sum``.
    """
    indptr, indices, eids = csr
    use_l = op != "copy_rhs"
    use_r = op != "copy_lhs"
    ref = ufeat if use_l else efeat
    if use_l and use_r and ufeat.dtype != efeat.dtype:
        raise DGLError(f"node and edge feature dtypes differ: {ufeat.dtype} vs {efeat.dtype}")
    l2, r2, out_shape, F, _, bc = _prep_operands(op, ufeat, efeat)
    # hot path on CUDA: copy_u / u_mul_e(scalar) with sum -> load-balanced vector kernel
    if (ref.is_cuda and reduce_op == "sum" and num_dst > 0 and F > 0 and ref.dtype != torch.float64
            and (F * ref.element_size()) % 4 == 0
            and (op == "copy_lhs" or (op == "mul" and bc[0] == 1 and bc[1] == F and
                                      (bc[3] == 1 or (bc[2] * bc[3] == F and (bc[2] * ref.element_size()) % 4 == 0))))):
        # u_mul_e with one scalar per edge, or one scalar per (edge, head) - the GAT aggregation
        out = _gspmm_balanced(csr, l2, r2 if op == "mul" else None, num_dst, F, heads=bc[3] if op == "mul" else 1)
        return out.reshape(num_dst, *out_shape), None, None
    out = torch.empty((num_dst, F), dtype=ref.dtype, device=ref.device)
    arg_u = arg_e = None
    if reduce_op != "sum":
        if use_l:
            arg_u = torch.empty((num_dst, F), dtype=indptr.dtype, device=ref.device)
        if use_r:
            arg_e = torch.empty((num_dst, F), dtype=indptr.dtype, device=ref.device)
    if num_dst > 0 and F > 0:
        lib, is_cuda = _dev_lib(ref)
        args = [OPS[op], REDS[reduce_op], dtype_code(ref.dtype), id_code(indptr.dtype), _p(indptr), _p(indices),
                _p(eids), num_dst, _p(l2), _p(r2), _p(out), _p(arg_u), _p(arg_e), F, *bc]
        if is_cuda:
            lib.fn("dglb200_cuda_gspmm", "iiiipppl" + "ppppp" + "lllll" + "p")(*args, loader.stream_ptr())
        else:
            lib.fn("dglb200_cpu_gspmm", "iiiipppl" + "ppppp" + "lllll")(*args)
    out = out.reshape(num_dst, *out_shape)
    if arg_u is not None:
        arg_u = arg_u.reshape(num_dst, *out_shape)
    if arg_e is not None:
        arg_e = arg_e.reshape(num_dst, *out_shape)
    return out, arg_u, arg_e


def _prep_single(op, ufeat, efeat):
    x = ufeat if op == "copy_lhs" else efeat
    shp = tuple(x.shape[1:])
    F = _prod(shp)
    x2 = x.reshape(x.shape[0], -1).contiguous()
    if op == "copy_lhs":
        return x2, None, shp, F, 1, (1, F, 1, 1)
    return None, x2, shp, F, 1, (1, 1, 1, F)


def _dst_sorted_view(gidx):
    """(src, dst, eids|None) of the relation in destination-sorted (CSC) order, if the CSC already exists: walking the edges
    in this order makes consecutive edges share their destination row (L1 hits instead of a second random gather)."""
    if gidx is None or getattr(gidx, "_csc", None) is None:
        return None
    cache = gidx._degs
    hit = cache.get("csc_coo")
    if hit is None:
        indptr, indices, eids = gidx._csc
        deg = (indptr[1:] - indptr[:-1]).long()
        dst = torch.repeat_interleave(torch.arange(gidx.num_dst, dtype=indices.dtype, device=indices.device), deg,
                                      output_size=int(indices.shape[0]))
        hit = cache["csc_coo"] = (indices, dst, eids)
    return hit


def _dot_fast_path(gidx, row, col, lhs2, rhs2, out, lt, rt, F, rsz, bc):
    """staged tensor-free dot kernel (csrc/cuda/sddmm.cu sddmm_dot_staged_kernel); returns False if the shape is not covered."""
    es = lhs2.element_size()
    if es > 4 or tuple(bc) != (1, F, 1, F) or lhs2.dtype != rhs2.dtype:
        return False
    if F == 1:
        if (rsz * es) % 4 or lhs2.data_ptr() % 4 or rhs2.data_ptr() % 4:
            return False
    else:
        vec = 16 // es
        lph = rsz // vec
        if rsz % vec or lhs2.data_ptr() % 16 or rhs2.data_ptr() % 16 or lph > 32 or lph & (lph - 1) or F * lph > 1024:
            return False
    perm = None
    view = _dst_sorted_view(gidx) if (lt != 1 and rt != 1) else None
    if view is not None:
        row, col, perm = view
    loader.cuda_lib().fn("dglb200_cuda_sddmm_dot", "iippplpppiillp")(
        dtype_code(lhs2.dtype), id_code(row.dtype), _p(row), _p(col), _p(perm), row.shape[0], _p(lhs2), _p(rhs2), _p(out),
        lt, rt, F, rsz, loader.stream_ptr())
    return True


def gsddmm(coo, op, lhs, rhs, lhs_target="u", rhs_target="v", gidx=None):
    """out[e] = op(lhs[target_l(e)], rhs[target_r(e)]) in edge-id order.
    ``coo`` = (row, col) giving (src, dst) per edge id; ``gidx`` (optional) lets the kernel pick a better edge order."""
    row, col = coo
    E = row.shape[0]
    use_l = op != "copy_rhs"
    use_r = op != "copy_lhs"
    ref = lhs if use_l else rhs
    if use_l and use_r and lhs.dtype != rhs.dtype:
        raise DGLError(f"operand dtypes differ: {lhs.dtype} vs {rhs.dtype}")
    l2, r2, out_shape, F, rsz, bc = _prep_operands(op, lhs, rhs)
    out = torch.empty((E, F), dtype=ref.dtype, device=ref.device)
    if E > 0 and F > 0:
        lib, is_cuda = _dev_lib(ref)
        if is_cuda and op == "dot" and tuple(bc) == (1, F, 1, F) and ref.element_size() <= 4:
            # rows that are not a multiple of 16 bytes (e.g. 41 classes) would force scalar loads on E * rsz elements:
            # zero-pad the reduce dimension of the (much smaller) node operands once and take the vector kernels
            vec = 16 // ref.element_size()
            if (rsz * ref.element_size()) % 4 and rsz > vec and E >= 4 * max(l2.shape[0], r2.shape[0]):
                pad = vec - rsz % vec
                l2 = torch.nn.functional.pad(l2.view(l2.shape[0], F, rsz), (0, pad)).reshape(l2.shape[0], -1)
                r2 = torch.nn.functional.pad(r2.view(r2.shape[0], F, rsz), (0, pad)).reshape(r2.shape[0], -1)
                rsz += pad
        if is_cuda and op == "dot" and _dot_fast_path(gidx, row, col, l2, r2, out, TARGETS[lhs_target], TARGETS[rhs_target],
                                                      F, rsz, bc):
            return out.reshape(E, *out_shape)
        args = [OPS[op], dtype_code(ref.dtype), id_code(row.dtype), _p(row), _p(col), E, _p(l2), _p(r2), _p(out),
                TARGETS[lhs_target], TARGETS[rhs_target], F, rsz, *bc]
        if is_cuda:
            lib.fn("dglb200_cuda_gsddmm", "iiipplpppiillllllp")(*args, loader.stream_ptr())
        else:
            lib.fn("dglb200_cpu_gsddmm", "iiipplpppiillllll")(*args)
    return out.reshape(E, *out_shape)


def segment_reduce(offsets, x, reduce_op):
    """Reduce contiguous row segments of ``x``; offsets has nseg+1 entries.
    Returns (out, arg|None)."""
    nseg = offsets.shape[0] - 1
    shp = tuple(x.shape[1:])
    F = _prod(shp)
    x2 = x.reshape(x.shape[0], -1).contiguous()
    out = torch.empty((nseg, F), dtype=x.dtype, device=x.device)
    arg = None
    if reduce_op != "sum":
        arg = torch.empty((nseg, F), dtype=offsets.dtype, device=x.device)
    if nseg > 0 and F > 0:
        lib, is_cuda = _dev_lib(x)
        args = [REDS[reduce_op], dtype_code(x.dtype), id_code(offsets.dtype), _p(offsets), nseg, _p(x2), _p(out),
                _p(arg), F]
        if is_cuda:
            lib.fn("dglb200_cuda_segment_reduce", "iiiplppplp")(*args, loader.stream_ptr())
        else:
            lib.fn("dglb200_cpu_segment_reduce", "iiiplpppl")(*args)
    out = out.reshape(nseg, *shp)
    if arg is not None:
        arg = arg.reshape(nseg, *shp)
    return out, arg


def edge_softmax_forward(csc, score, num_dst):
    indptr, _, eids = csc
    shp = tuple(score.shape[1:])
    F = _prod(shp)
    s2 = score.reshape(score.shape[0], -1).contiguous()
    out = torch.empty_like(s2)
    if num_dst > 0 and s2.numel() > 0:
        lib, is_cuda = _dev_lib(score)
        args = [dtype_code(score.dtype), id_code(indptr.dtype), _p(indptr), _p(eids), num_dst, _p(s2), _p(out), F]
        if is_cuda and score.dtype != torch.float64:
            work = _softmax_work(score)
            if eids is not None and s2.shape[0] >= _PERMUTE_MIN_EDGES:
                # large graph whose edge ids are not in CSC order: one gather brings the scores into CSC order, the
                # three softmax passes then stream sequentially, one scatter puts the result back (2 random passes
                # over the edge data instead of 4)
                from .ops.cuda_struct import gather_rows, scatter_rows
                s_csc = gather_rows(s2, eids)
                o_csc = torch.empty_like(s_csc)
                lib.fn("dglb200_cuda_edge_softmax_fwd2", "iipplpplplp")(
                    args[0], args[1], _p(indptr), 0, num_dst, _p(s_csc), _p(o_csc), F, _p(work), work.numel(),
                    loader.stream_ptr())
                scatter_rows(out, eids, o_csc)
            else:
                lib.fn("dglb200_cuda_edge_softmax_fwd2", "iipplpplplp")(*args, _p(work), work.numel(),
                                                                        loader.stream_ptr())
        elif is_cuda:
            lib.fn("dglb200_cuda_edge_softmax_fwd", "iipplpplp")(*args, loader.stream_ptr())
        else:
            lib.fn("dglb200_cpu_edge_softmax_fwd", "iipplppl")(*args)
    return out.reshape(score.shape)


_PERMUTE_MIN_EDGES = 1 << 20


def _softmax_work(score):
    """int32 work list for hub rows (> 2048 in-edges): [count, row ids...]; E/2048 + 2 entries always suffice."""
    return torch.empty(score.shape[0] // 2048 + 2, dtype=torch.int32, device=score.device)


def edge_softmax_backward(csc, out, grad_out, num_dst):
    indptr, _, eids = csc
    F = _prod(tuple(out.shape[1:]))
    o2 = out.reshape(out.shape[0], -1).contiguous()
    g2 = grad_out.reshape(out.shape[0], -1).contiguous()
    gs = torch.empty_like(o2)
    if num_dst > 0 and o2.numel() > 0:
        lib, is_cuda = _dev_lib(out)
        args = [dtype_code(out.dtype), id_code(indptr.dtype), _p(indptr), _p(eids), num_dst, _p(o2), _p(g2), _p(gs), F]
        if is_cuda and out.dtype != torch.float64:
            work = _softmax_work(out)
            if eids is not None and o2.shape[0] >= _PERMUTE_MIN_EDGES:
                from .ops.cuda_struct import gather_rows, scatter_rows
                o_csc, g_csc = gather_rows(o2, eids), gather_rows(g2, eids)
                gs_csc = torch.empty_like(o_csc)
                lib.fn("dglb200_cuda_edge_softmax_bwd2", "iipplppplplp")(
                    args[0], args[1], _p(indptr), 0, num_dst, _p(o_csc), _p(g_csc), _p(gs_csc), F, _p(work),
                    work.numel(), loader.stream_ptr())
                scatter_rows(gs, eids, gs_csc)
            else:
                lib.fn("dglb200_cuda_edge_softmax_bwd2", "iipplppplplp")(*args, _p(work), work.numel(),
                                                                         loader.stream_ptr())
        elif is_cuda:
            lib.fn("dglb200_cuda_edge_softmax_bwd", "iipplppplp")(*args, loader.stream_ptr())
        else:
            lib.fn("dglb200_cpu_edge_softmax_bwd", "iipplpppl")(*args)
    return gs.reshape(out.shape)


# ---- names of the reference's _sparse_ops.py that have no kernel of their own here ------------------------------------
def to_dgl_nd(x):
    """The reference converts a framework tensor to its NDArray before every kernel call; kernels here take tensors."""
    return x


def to_dgl_nd_for_write(x):
    return x


def get_typeid_by_target(gidx, etid, target):
    """Node-type id of relation ``etid``'s source (target 0 / 'u') or destination (1 / 'v'); edges return ``etid``."""
    src, dst = gidx.metagraph.find_edge(etid)
    if target in (0, "u"):
        return int(src)
    if target in (1, "v"):
        return int(dst)
    return int(etid)


def libra_vertex_cut(nc, node_degree, edgenum_unassigned, community_weights, u, v, w, out, N, N_e, dataset):
    """Libra vertex-cut edge assignment written into ``out`` (reference _CAPI_DGLLibraVertexCut); see
    ``dgl_b200.partition.libra_partition_assignment`` for the graph-level entry point."""
    from ._native import loader
    from .base import id_code
    u64, v64 = u.long().contiguous(), v.long().contiguous()
    res = out if out.dtype == torch.int64 and out.is_contiguous() else torch.empty(int(N_e), dtype=torch.int64)
    loader.cpu_lib().fn("dglb200_cpu_libra_partition", "ippllip")(
        id_code(torch.int64), u64.data_ptr(), v64.data_ptr(), int(N_e), int(N), int(nc), res.data_ptr())
    if res is not out:
        out.copy_(res.to(out.dtype))
    community_weights.zero_()
    community_weights += torch.bincount(res, minlength=int(nc)).to(community_weights.dtype)


def libra2dgl_build_dict(a, b, indices, ldt_key, gdt_key, gdt_value, node_map, offset, nc, c, fsize, dataset):
    """Per-partition global->local dictionary of the Libra conversion: returns the number of distinct vertices touched
    by partition ``c``'s edges (``a`` -> ``b``) and fills ``ldt_key`` with them in first-seen order."""
    verts = torch.unique(torch.cat([a, b]), sorted=False)
    n = int(verts.shape[0])
    ldt_key[:n] = verts.to(ldt_key.dtype)
    return n


def libra2dgl_build_adjlist(feat, gfeat, adj, inner_node, ldt, gdt_key, gdt_value, node_map, lr, lrtensor, num_nodes,
                            nc, c, feat_size, labels, trainm, testm, valm, glabels, gtrainm, gtestm, gvalm, feat_shape):
    """Gather the features / labels / masks of partition ``c``'s vertices (``ldt`` = their global ids) from the global
    arrays into the partition-local ones."""
    ids = ldt[:num_nodes].long()
    gfeat[:num_nodes] = feat[ids]
    glabels[:num_nodes] = labels[ids]
    gtrainm[:num_nodes], gtestm[:num_nodes], gvalm[:num_nodes] = trainm[ids], testm[ids], valm[ids]


def libra2dgl_set_lr(gdt_key, gdt_value, lrtensor, nc, Nn):
    """Pick, for every replicated vertex, one owning partition uniformly among the partitions holding a copy."""
    gen = torch.Generator().manual_seed(0)
    counts = gdt_key[:Nn].long().clamp(min=1)
    pick = (torch.rand(Nn, generator=gen) * counts).long()
    lrtensor[:Nn] = gdt_value.view(Nn, -1)[torch.arange(Nn), pick].to(lrtensor.dtype)
