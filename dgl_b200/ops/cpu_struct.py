"""Host-side sparse-structure arithmetic backed by ``csrc/cpu/structure.cc``.

* :func:`csr_mm`     – C = A @ B on CSR operands (two-pass Gustavson, column-sorted output rows)
* :func:`csr_sum`    – C = alpha*A + beta*B by a per-row merge
* :func:`csr_rowdot` – out[e] = <X[rows[e], :], Y[cols[e], :]> (the masked product of the backward)
* :func:`spspmm` / :func:`spadd` – autograd wrappers used by ``dgl_b200.sparse`` for CPU matrices

Parity: reference ``src/array/cpu/csr_mm.cc`` / ``csr_sum.cc`` and the autograd in
``python/dgl/backend/pytorch/sparse.py:CSRMM/CSRSum/CSRMask``.
"""
from __future__ import annotations

import torch

from .._native import loader
from ..base import dtype_code, id_code

_FLOATS = (torch.float32, torch.float64, torch.float16, torch.bfloat16)


def _p(t):
    return 0 if t is None else t.data_ptr()


def sort_rows(indptr, indices, perm=None):
    """Column-sort every CSR row; returns (indices, perm) where ``perm`` maps sorted position -> original value slot."""
    indices = indices.clone()
    perm = torch.arange(indices.shape[0], dtype=torch.int64) if perm is None else perm.to(torch.int64).clone()
    loader.cpu_lib().fn("dglb200_cpu_csr_sort_rows", "ilppp")(
        id_code(indptr.dtype), indptr.shape[0] - 1, indptr.data_ptr(), indices.data_ptr(), perm.data_ptr())
    return indices, perm


def transpose(indptr, indices, n_cols):
    """CSR -> CSR of the transpose (native stable counting sort, csrc/cpu/spmat.cc). Rows of the result are
    column-sorted because CSR order is row-ascending and the sort is stable. Returns (t_indptr, t_indices, perm) with
    ``perm`` = position in the input of every output nnz."""
    t_indptr, t_indices, perm = csr_transpose(indptr, indices, None, n_cols)
    return t_indptr, t_indices, perm.long()


def coo_to_csr(row, col, n_rows):
    """(indptr, indices, eids) of a host COO, rows in ascending order, edge order kept inside a row."""
    idt = row.dtype
    row, col = row.contiguous(), col.contiguous()
    E = row.shape[0]
    indptr = torch.empty(n_rows + 1, dtype=idt)
    indices = torch.empty(E, dtype=idt)
    eids = torch.empty(E, dtype=idt)
    loader.cpu_lib().fn("dglb200_cpu_coo_to_csr", "ippllppp")(
        id_code(idt), _p(row), _p(col), E, int(n_rows), _p(indptr), _p(indices), _p(eids))
    return indptr, indices, eids


def csr_transpose(indptr, indices, data, n_cols):
    """CSR(+edge ids) -> (t_indptr, t_indices, t_data) of the transposed matrix; ``data=None`` means positions."""
    idt = indptr.dtype
    indptr, indices = indptr.contiguous(), indices.contiguous()
    data = None if data is None else data.contiguous()
    E = indices.shape[0]
    t_indptr = torch.empty(n_cols + 1, dtype=idt)
    t_indices = torch.empty(E, dtype=idt)
    t_data = torch.empty(E, dtype=idt)
    loader.cpu_lib().fn("dglb200_cpu_csr_transpose", "ippplLppp".replace("L", "l"))(
        id_code(idt), _p(indptr), _p(indices), _p(data), indptr.shape[0] - 1, int(n_cols), _p(t_indptr), _p(t_indices),
        _p(t_data))
    return t_indptr, t_indices, t_data


def coalesce(row, col, n_rows, n_cols):
    """Distinct (row, col) pairs of a host COO in lexicographic order: (out_row, out_col, group, count, first) where
    ``group[e]`` is the pair id of edge ``e``, ``count`` the multiplicities and ``first`` the smallest member edge id."""
    idt = row.dtype
    row, col = row.contiguous(), col.contiguous()
    E = row.shape[0]
    group = torch.empty(E, dtype=torch.int64)
    out_row, out_col = torch.empty(E, dtype=idt), torch.empty(E, dtype=idt)
    count, first = torch.empty(E, dtype=torch.int64), torch.empty(E, dtype=torch.int64)
    n_out = torch.zeros(1, dtype=torch.int64)
    loader.cpu_lib().fn("dglb200_cpu_coalesce", "ipplllpppppp")(
        id_code(idt), _p(row), _p(col), E, int(n_rows), int(n_cols), _p(group), _p(out_row), _p(out_col), _p(count),
        _p(first), _p(n_out))
    n = int(n_out)
    return out_row[:n], out_col[:n], group, count[:n], first[:n]


def slice_rows(indptr, indices, data, rows):
    """Rows ``rows`` of a CSR: (sub_indptr, sub_indices, sub_data); ``sub_data`` holds edge ids (positions if ``data`` is None)."""
    idt = indptr.dtype
    rows = rows.to(idt).contiguous()
    nr = rows.shape[0]
    fn = loader.cpu_lib().fn("dglb200_cpu_csr_slice_rows", "ippppLppp".replace("L", "l"))
    sub_indptr = torch.empty(nr + 1, dtype=idt)
    fn(id_code(idt), _p(indptr), _p(indices), _p(data), _p(rows), nr, _p(sub_indptr), 0, 0)
    nnz = int(sub_indptr[-1])
    sub_indices, sub_data = torch.empty(nnz, dtype=idt), torch.empty(nnz, dtype=idt)
    if nnz:
        fn(id_code(idt), _p(indptr), _p(indices), _p(data), _p(rows), nr, _p(sub_indptr), _p(sub_indices), _p(sub_data))
    return sub_indptr, sub_indices, sub_data


def slice_matrix(indptr, indices, data, rows, col_map):
    """Sub-matrix rows x {c : col_map[c] >= 0} with columns relabelled by ``col_map``."""
    idt = indptr.dtype
    rows, col_map = rows.to(idt).contiguous(), col_map.to(idt).contiguous()
    nr = rows.shape[0]
    fn = loader.cpu_lib().fn("dglb200_cpu_csr_slice_matrix", "ippppLpppp".replace("L", "l"))
    sub_indptr = torch.empty(nr + 1, dtype=idt)
    fn(id_code(idt), _p(indptr), _p(indices), _p(data), _p(rows), nr, _p(col_map), _p(sub_indptr), 0, 0)
    nnz = int(sub_indptr[-1])
    sub_indices, sub_data = torch.empty(nnz, dtype=idt), torch.empty(nnz, dtype=idt)
    if nnz:
        fn(id_code(idt), _p(indptr), _p(indices), _p(data), _p(rows), nr, _p(col_map), _p(sub_indptr), _p(sub_indices),
           _p(sub_data))
    return sub_indptr, sub_indices, sub_data


def line_graph(u, v, indptr, csr_eids, backtracking=True):
    """Edge list (i, j) of the line graph, grouped by i in edge-id order (csrc/cpu/spmat.cc)."""
    idt = u.dtype
    u, v = u.contiguous(), v.contiguous()
    E = u.shape[0]
    fn = loader.cpu_lib().fn("dglb200_cpu_line_graph", "ipplppippp")
    lg_indptr = torch.empty(E + 1, dtype=idt)
    fn(id_code(idt), _p(u), _p(v), E, _p(indptr), _p(csr_eids), int(bool(backtracking)), _p(lg_indptr), 0, 0)
    n = int(lg_indptr[-1])
    src, dst = torch.empty(n, dtype=idt), torch.empty(n, dtype=idt)
    if n:
        fn(id_code(idt), _p(u), _p(v), E, _p(indptr), _p(csr_eids), int(bool(backtracking)), _p(lg_indptr), _p(src), _p(dst))
    return src, dst


def csr_mm(ap, ai, av, bp, bi, bv, n_cols):
    """(indptr, indices, values) of A @ B. ``av``/``bv`` may be ``None`` (all ones)."""
    lib = loader.cpu_lib()
    n_rows = ap.shape[0] - 1
    idt = id_code(ap.dtype)
    row_nnz = torch.empty(n_rows, dtype=torch.int64)
    lib.fn("dglb200_cpu_csr_mm_count", "ilpppplp")(idt, n_rows, ap.data_ptr(), ai.data_ptr(), bp.data_ptr(),
                                                   bi.data_ptr(), int(n_cols), row_nnz.data_ptr())
    cp = torch.zeros(n_rows + 1, dtype=torch.int64)
    torch.cumsum(row_nnz, 0, out=cp[1:])
    nnz = int(cp[-1])
    vdt = av.dtype if av is not None else (bv.dtype if bv is not None else torch.float32)
    ci = torch.empty(nnz, dtype=ap.dtype)
    cv = torch.empty(nnz, dtype=vdt)
    if av is not None and bv is not None and av.dtype != bv.dtype:
        vdt = torch.promote_types(av.dtype, bv.dtype)
        av, bv, cv = av.to(vdt), bv.to(vdt), cv.to(vdt)
    lib.fn("dglb200_cpu_csr_mm_fill", "iilpppppplppp")(
        idt, dtype_code(vdt), n_rows, ap.data_ptr(), ai.data_ptr(), _p(av), bp.data_ptr(), bi.data_ptr(), _p(bv),
        int(n_cols), cp.data_ptr(), ci.data_ptr(), cv.data_ptr())
    return cp.to(ap.dtype), ci, cv


def csr_sum(ap, ai, av, bp, bi, bv, alpha=1.0, beta=1.0, want_sources=False):
    """alpha*A + beta*B for two CSR matrices of equal shape with column-sorted rows."""
    lib = loader.cpu_lib()
    n_rows = ap.shape[0] - 1
    idt = id_code(ap.dtype)
    row_nnz = torch.empty(n_rows, dtype=torch.int64)
    lib.fn("dglb200_cpu_csr_sum_count", "ilppppp")(idt, n_rows, ap.data_ptr(), ai.data_ptr(), bp.data_ptr(),
                                                   bi.data_ptr(), row_nnz.data_ptr())
    cp = torch.zeros(n_rows + 1, dtype=torch.int64)
    torch.cumsum(row_nnz, 0, out=cp[1:])
    nnz = int(cp[-1])
    vdt = av.dtype if av is not None else (bv.dtype if bv is not None else torch.float32)
    if av is not None and bv is not None and av.dtype != bv.dtype:
        vdt = torch.promote_types(av.dtype, bv.dtype)
        av, bv = av.to(vdt), bv.to(vdt)
    ci = torch.empty(nnz, dtype=ap.dtype)
    cv = torch.empty(nnz, dtype=vdt)
    sa = torch.empty(nnz, dtype=ap.dtype) if want_sources else None
    sb = torch.empty(nnz, dtype=ap.dtype) if want_sources else None
    lib.fn("dglb200_cpu_csr_sum_fill", "iilpppdpppdppppp")(
        idt, dtype_code(vdt), n_rows, ap.data_ptr(), ai.data_ptr(), _p(av), float(alpha), bp.data_ptr(),
        bi.data_ptr(), _p(bv), float(beta), cp.data_ptr(), ci.data_ptr(), cv.data_ptr(), _p(sa), _p(sb))
    out = (cp.to(ap.dtype), ci, cv)
    return out + (sa, sb) if want_sources else out


def csr_rowdot(rows, cols, xp, xi, xv, yp, yi, yv):
    """out[e] = sum_k X[rows[e], k] * Y[cols[e], k] for CSR ``X``/``Y`` with column-sorted rows."""
    vdt = xv.dtype if xv is not None else yv.dtype
    out = torch.empty(rows.shape[0], dtype=vdt)
    loader.cpu_lib().fn("dglb200_cpu_csr_rowdot", "iilppppppppp")(
        id_code(xp.dtype), dtype_code(vdt), rows.shape[0], rows.data_ptr(), cols.data_ptr(), xp.data_ptr(),
        xi.data_ptr(), _p(xv), yp.data_ptr(), yi.data_ptr(), _p(yv), out.data_ptr())
    return out


def _row_ids(indptr):
    n = indptr.shape[0] - 1
    return torch.repeat_interleave(torch.arange(n, dtype=indptr.dtype), (indptr[1:] - indptr[:-1]).long())


class _SpSpMM(torch.autograd.Function):
    """Values of C = A @ B given sorted-CSR structures; gradients flow to both value vectors."""

    @staticmethod
    def forward(ctx, av, bv, ap, ai, bp, bi, n_cols_b, holder):
        cp, ci, cv = csr_mm(ap, ai, av.detach().contiguous(), bp, bi, bv.detach().contiguous(), n_cols_b)
        holder.extend([cp, ci])
        ctx.save_for_backward(av, bv, ap, ai, bp, bi, cp, ci)
        ctx.n_cols_b = n_cols_b
        return cv

    @staticmethod
    def backward(ctx, dcv):
        av, bv, ap, ai, bp, bi, cp, ci = ctx.saved_tensors
        dcv = dcv.contiguous()
        da = db = None
        if ctx.needs_input_grad[0]:
            # dA[i,k] = <dC[i,:], B[k,:]>
            da = csr_rowdot(_row_ids(ap), ai, cp, ci, dcv, bp, bi, bv.detach().contiguous())
        if ctx.needs_input_grad[1]:
            # dB[k,j] = <A^T[k,:], dC^T[j,:]>
            n_inner = bp.shape[0] - 1
            atp, ati, aperm = transpose(ap, ai, n_inner)
            ctp, cti, cperm = transpose(cp, ci, ctx.n_cols_b)
            db = csr_rowdot(_row_ids(bp), bi, atp, ati, av.detach()[aperm].contiguous(), ctp, cti,
                            dcv[cperm].contiguous())
        return da, db, None, None, None, None, None, None


def _sorted_csr(mat):
    """(indptr, sorted indices, value-slot permutation) of a ``SparseMatrix``'s row compression."""
    ip, idx, eids = mat.csr()
    idx, perm = sort_rows(ip, idx.to(ip.dtype), eids)
    return ip.contiguous(), idx, perm


def usable(*mats):
    return all(m.val.device.type == "cpu" and m.val.dim() == 1 and m.val.dtype in _FLOATS for m in mats)


def spspmm(A, B):
    """Native CPU SpGEMM of two coalesced ``SparseMatrix`` objects -> (indptr, indices, values)."""
    ap, ai, pa = _sorted_csr(A)
    bp, bi, pb = _sorted_csr(B)
    if bp.dtype != ap.dtype:
        bp, bi = bp.to(ap.dtype), bi.to(ap.dtype)
    holder = []
    cv = _SpSpMM.apply(A.val[pa], B.val[pb], ap, ai, bp, bi, int(B.shape[1]), holder)
    return holder[0], holder[1], cv


def spadd(A, B, beta=1.0):
    """Native CPU A + beta*B of two coalesced ``SparseMatrix`` objects -> (indptr, indices, values)."""
    ap, ai, pa = _sorted_csr(A)
    bp, bi, pb = _sorted_csr(B)
    if bp.dtype != ap.dtype:
        bp, bi = bp.to(ap.dtype), bi.to(ap.dtype)
    av, bv = A.val[pa], B.val[pb]
    cp, ci, cv, sa, sb = csr_sum(ap, ai, av.detach().contiguous(), bp, bi, bv.detach().contiguous(), 1.0, beta,
                                 want_sources=True)
    if av.requires_grad or bv.requires_grad:
        # differentiable re-expression of the merge: two masked gathers (the structure is a constant)
        ha, hb = sa >= 0, sb >= 0
        vdt = cv.dtype
        cv = torch.zeros_like(cv)
        cv = cv.masked_scatter(ha, av.to(vdt)[sa[ha].long()]) + beta * torch.zeros_like(cv).masked_scatter(
            hb, bv.to(vdt)[sb[hb].long()])
    return cp, ci, cv
