"""``dgl.sparse``: sparse-matrix API over the same sm_100a g-SpMM / g-SDDMM / edge-softmax kernels.

Parity: /root/reference/python/dgl/sparse/*.py (SparseMatrix, spmatrix/from_coo/from_csr/from_csc/diag/identity/
val_like, spmm/bspmm/spspmm/sddmm/bsddmm, softmax, reductions, elementwise, broadcast) and the C++ library
/root/reference/dgl_sparse/src/*.cc whose compute it delegates to libdgl (matmul.cc:20-140). Here the matrix is a
thin Python object over torch tensors with lazily materialised COO/CSR/CSC; differentiability w.r.t. the non-zero
values comes from the autograd Functions in ops/sparse_autograd.py.
"""
from __future__ import annotations

import operator

import torch

from ..base import DGLError
from ..graph_index import UnitGraph
from ..ops import sparse_autograd as SA
from ..ops import cpu_struct as _cpu_struct

__all__ = ["SparseMatrix", "spmatrix", "from_coo", "from_csr", "from_csc", "diag", "identity", "val_like", "spmm",
           "bspmm", "spspmm", "sddmm", "bsddmm", "softmax", "sum", "smax", "smin", "smean", "sprod", "add", "sub",
           "mul", "div", "power", "neg", "inv", "matmul", "sp_broadcast_v", "sp_add_v", "sp_sub_v", "sp_mul_v",
           "sp_div_v", "to_torch_sparse_coo", "to_torch_sparse_csr", "to_torch_sparse_csc", "from_torch_sparse", "reduce", "sp_add", "sp_sub", "sp_mul", "sp_div", "sp_power",
           "spsp_add", "spsp_mul", "spsp_div", "is_scalar", "load_dgl_sparse"]


class SparseMatrix:
    """2-D sparse matrix with scalar or vector non-zero values (``val``: [nnz] or [nnz, D])."""

    def __init__(self, ug: UnitGraph, val, shape):
        self._g = ug  # relation with src = column index, dst = row index (so SpMM aggregates into rows)
        self.val = val
        self._shape = tuple(shape)

    # ---- structure ---------------------------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    @property
    def nnz(self):
        return self._g.num_edges()

    @property
    def dtype(self):
        return self.val.dtype

    @property
    def device(self):
        return self.val.device

    @property
    def row(self):
        return self._g.coo()[1]

    @property
    def col(self):
        return self._g.coo()[0]

    def coo(self):
        c, r = self._g.coo()
        return r, c

    def indices(self):
        r, c = self.coo()
        return torch.stack([r, c])

    def csr(self):
        """(indptr, indices, value_indices) compressed over rows."""
        ip, idx, e = self._g.csc()  # rows of the matrix are dst nodes
        return ip, idx, e

    def csc(self):
        ip, idx, e = self._g.csr()
        return ip, idx, e

    def to_dense(self):
        r, c = self.coo()
        shape = self._shape + tuple(self.val.shape[1:])
        out = torch.zeros(shape, dtype=self.val.dtype, device=self.val.device)
        out.index_put_((r.long(), c.long()), self.val, accumulate=True)
        return out

    def t(self):
        return SparseMatrix(self._g.reverse(), self.val, (self._shape[1], self._shape[0]))

    transpose = t

    @property
    def T(self):
        return self.t()

    def to(self, device=None, dtype=None):
        g, v = self._g, self.val
        if device is not None:
            g, v = g.to(device), v.to(device)
        if dtype is not None:
            v = v.to(dtype)
        return SparseMatrix(g, v, self._shape)

    def cuda(self):
        return self.to(device="cuda")

    def cpu(self):
        return self.to(device="cpu")

    def float(self):
        return self.to(dtype=torch.float32)

    def double(self):
        return self.to(dtype=torch.float64)

    def int(self):
        return self.to(dtype=torch.int32)

    def long(self):
        return self.to(dtype=torch.int64)

    def is_diag(self):
        r, c = self.coo()
        return bool((r == c).all())

    def compact(self, dim, leading_indices=None):
        """Drop empty rows (``dim=0``) / columns (``dim=1``) and relabel that dimension; ``leading_indices`` are placed
        first (like ``include_dst_in_src`` of ``to_block``). Returns (matrix, new index -> original index)
        (reference sparse_matrix.py:683)."""
        r, c = self.coo()
        ids = (r if dim == 0 else c).long()
        if leading_indices is not None:
            lead = torch.as_tensor(leading_indices, device=ids.device).long()
            present = torch.unique(ids)
            rest = present[~torch.isin(present, lead)]
            original = torch.cat([lead, rest])
        else:
            original = torch.unique(ids)
        lookup = torch.full((int(max(self.shape[dim], int(original.max()) + 1 if original.numel() else 0)),), -1,
                            dtype=torch.long, device=ids.device)
        lookup[original] = torch.arange(original.shape[0], device=ids.device)
        new = lookup[ids]
        shape = (original.shape[0], self.shape[1]) if dim == 0 else (self.shape[0], original.shape[0])
        rows, cols = (new, c.long()) if dim == 0 else (r.long(), new)
        return from_coo(rows, cols, self.val, shape), original

    def has_duplicate(self):
        r, c = self.coo()
        key = r.long() * max(self._shape[1], 1) + c.long()
        return torch.unique(key).numel() != key.numel()

    def coalesce(self):
        r, c = self.coo()
        key = r.long() * max(self._shape[1], 1) + c.long()
        uniq, inv = torch.unique(key, return_inverse=True)
        val = torch.zeros((uniq.shape[0],) + tuple(self.val.shape[1:]), dtype=self.val.dtype, device=self.val.device)
        val = val.index_add(0, inv, self.val)
        nr = (uniq // max(self._shape[1], 1)).to(r.dtype)
        nc = (uniq % max(self._shape[1], 1)).to(c.dtype)
        return from_coo(nr, nc, val, self._shape)

    def index_select(self, dim, index):
        r, c = self.coo()
        tgt = r if dim == 0 else c
        n = self._shape[dim]
        pos = torch.full((n,), -1, dtype=torch.long, device=tgt.device)
        # allow repeated indices: expand matches
        outs_r, outs_c, outs_v = [], [], []
        for new_i, old_i in enumerate(index.tolist()):
            m = tgt == old_i
            k = int(m.sum())
            if k == 0:
                continue
            outs_r.append(torch.full((k,), new_i, dtype=r.dtype, device=r.device) if dim == 0 else r[m])
            outs_c.append(c[m] if dim == 0 else torch.full((k,), new_i, dtype=c.dtype, device=c.device))
            outs_v.append(self.val[m])
        shape = (len(index), self._shape[1]) if dim == 0 else (self._shape[0], len(index))
        if not outs_r:
            e = torch.empty(0, dtype=r.dtype, device=r.device)
            return from_coo(e, e.clone(), self.val[:0], shape)
        return from_coo(torch.cat(outs_r), torch.cat(outs_c), torch.cat(outs_v), shape)

    def range_select(self, dim, rng):
        lo, hi = rng.start or 0, rng.stop
        r, c = self.coo()
        tgt = r if dim == 0 else c
        m = (tgt >= lo) & (tgt < hi)
        nr = r[m] - (lo if dim == 0 else 0)
        nc = c[m] - (lo if dim == 1 else 0)
        shape = (hi - lo, self._shape[1]) if dim == 0 else (self._shape[0], hi - lo)
        return from_coo(nr, nc, self.val[m], shape)

    def sample(self, dim, fanout, ids=None, replace=False, bias=False):
        """Sample up to ``fanout`` non-zeros per row (dim=0) / column (dim=1)."""
        from ..sampling.neighbor import sample_rel
        g = self._g if dim == 0 else self._g.reverse()
        n = self._shape[dim]
        seeds = torch.arange(n, device=self.device, dtype=g.idtype) if ids is None else ids.to(g.idtype)
        prob = self.val.float() if bias else None
        src, dst, eid = sample_rel(g, seeds, fanout, "in", prob, replace)
        r, c = (dst, src) if dim == 0 else (src, dst)
        return from_coo(r, c, self.val[eid.long()], self._shape)

    # ---- reductions ----------------------------------------------------------------------------
    def _reduce(self, dim, op):
        return _reduce(self, dim, op)

    def sum(self, dim=None):
        return _reduce(self, dim, "sum")

    def smax(self, dim=None):
        return _reduce(self, dim, "max")

    def smin(self, dim=None):
        return _reduce(self, dim, "min")

    def smean(self, dim=None):
        return _reduce(self, dim, "mean")

    def sprod(self, dim=None):
        return _reduce(self, dim, "prod")

    def softmax(self, dim=1):
        return softmax(self, dim)

    # ---- operators -------------------------------------------------------------------------------
    def __add__(self, other):
        return add(self, other)

    def __radd__(self, other):
        return add(self, other)

    def __sub__(self, other):
        return sub(self, other)

    def __mul__(self, other):
        return mul(self, other)

    def __rmul__(self, other):
        return mul(self, other)

    def __truediv__(self, other):
        return div(self, other)

    def __pow__(self, s):
        return power(self, s)

    def __neg__(self):
        return neg(self)

    def neg(self):
        return neg(self)

    def inv(self):
        return inv(self)

    def __matmul__(self, other):
        return matmul(self, other)

    def __rmatmul__(self, other):
        if torch.is_tensor(other):
            return matmul(self.t(), other.transpose(0, 1) if other.dim() == 2 else other).transpose(0, 1)
        return NotImplemented

    def __repr__(self):
        r, c = self.coo()
        return (f"SparseMatrix(indices={torch.stack([r, c])},\n             values={self.val},\n"
                f"             shape={self._shape}, nnz={self.nnz})")


# --------------------------------------------------------------------------------------- construction
def from_coo(row, col, val=None, shape=None):
    row, col = torch.as_tensor(row), torch.as_tensor(col)
    if shape is None:
        shape = (int(row.max()) + 1 if row.numel() else 0, int(col.max()) + 1 if col.numel() else 0)
    if val is None:
        val = torch.ones(row.shape[0], device=row.device)
    if col.dtype != row.dtype:
        col = col.to(row.dtype)
    ug = UnitGraph.from_coo(shape[1], shape[0], col, row)
    return SparseMatrix(ug, val, shape)


def spmatrix(indices, val=None, shape=None):
    indices = torch.as_tensor(indices)
    return from_coo(indices[0], indices[1], val, shape)


def from_csr(indptr, indices, val=None, shape=None):
    indptr, indices = torch.as_tensor(indptr), torch.as_tensor(indices)
    nrow = indptr.shape[0] - 1
    if shape is None:
        shape = (nrow, int(indices.max()) + 1 if indices.numel() else 0)
    if val is None:
        val = torch.ones(indices.shape[0], device=indices.device)
    ug = UnitGraph.from_csc(shape[1], shape[0], indptr, indices.to(indptr.dtype))
    return SparseMatrix(ug, val, shape)


def from_csc(indptr, indices, val=None, shape=None):
    indptr, indices = torch.as_tensor(indptr), torch.as_tensor(indices)
    ncol = indptr.shape[0] - 1
    if shape is None:
        shape = (int(indices.max()) + 1 if indices.numel() else 0, ncol)
    if val is None:
        val = torch.ones(indices.shape[0], device=indices.device)
    ug = UnitGraph.from_csr(shape[1], shape[0], indptr, indices.to(indptr.dtype))
    return SparseMatrix(ug, val, shape)


def diag(val, shape=None):
    n = val.shape[0]
    idx = torch.arange(n, device=val.device)
    if shape is None:
        shape = (n, n)
    return from_coo(idx, idx, val, shape)


def identity(shape, d=None, dtype=None, device=None):
    n = min(shape)
    val = torch.ones(n if d is None else (n, d), dtype=dtype or torch.float32, device=device)
    idx = torch.arange(n, device=device)
    return from_coo(idx, idx, val, shape)


def val_like(mat, val):
    return SparseMatrix(mat._g, val, mat.shape)


def from_torch_sparse(t):
    t = t.coalesce() if t.layout == torch.sparse_coo else t.to_sparse_coo().coalesce()
    idx = t.indices()
    return from_coo(idx[0], idx[1], t.values(), tuple(t.shape[:2]))


def to_torch_sparse_csr(mat):
    """torch sparse CSR tensor of ``mat`` (the order of the non-zeros may change; reference sparse_matrix.py:1373)."""
    indptr, indices, eids = mat.csr()
    val = mat.val if eids is None else mat.val[eids.long()]
    return torch.sparse_csr_tensor(indptr.long(), indices.long(), val, size=tuple(mat.shape) + tuple(mat.val.shape[1:]))


def to_torch_sparse_csc(mat):
    """torch sparse CSC tensor of ``mat`` (reference sparse_matrix.py:1411)."""
    indptr, indices, eids = mat.csc()
    val = mat.val if eids is None else mat.val[eids.long()]
    return torch.sparse_csc_tensor(indptr.long(), indices.long(), val, size=tuple(mat.shape) + tuple(mat.val.shape[1:]))


def to_torch_sparse_coo(mat):
    r, c = mat.coo()
    return torch.sparse_coo_tensor(torch.stack([r.long(), c.long()]), mat.val, mat.shape + tuple(mat.val.shape[1:]))


# --------------------------------------------------------------------------------------------- matmul
def spmm(A, X):
    """A @ X with dense X [N] or [N, F] (one g-SpMM: u_mul_e_sum)."""
    if not isinstance(A, SparseMatrix):
        raise DGLError("spmm expects a SparseMatrix as the first operand")
    squeeze = X.dim() == 1
    Xv = X.unsqueeze(-1) if squeeze else X
    if A.val.dim() != 1:
        raise DGLError("spmm needs scalar non-zero values; use bspmm for batched values")
    out = SA.gspmm(A._g, "mul", "sum", Xv, A.val.to(Xv.dtype).unsqueeze(-1))
    return out.squeeze(-1) if squeeze else out


def bspmm(A, X):
    """Batched: A.val [nnz, B], X [N, F, B] -> [M, F, B]."""
    val = A.val.unsqueeze(1)  # [nnz, 1, B]
    return SA.gspmm(A._g, "mul", "sum", X, val.to(X.dtype))


def spspmm(A, B):
    """Sparse @ sparse (SpGEMM via torch.sparse / cuSPARSE); differentiable w.r.t. both value vectors."""
    if A.shape[1] != B.shape[0]:
        raise DGLError(f"shape mismatch for spspmm: {A.shape} @ {B.shape}")
    A = A.coalesce() if A.has_duplicate() else A
    B = B.coalesce() if B.has_duplicate() else B
    if _cpu_struct.usable(A, B):
        # host matrices: native two-pass Gustavson SpGEMM with a masked row-dot backward (csrc/cpu/structure.cc)
        cp, ci, cv = _cpu_struct.spspmm(A, B)
        return from_csr(cp, ci, cv, (A.shape[0], B.shape[1]))
    if A.val.is_cuda and A.val.dim() == 1 and B.val.dim() == 1 and not (A.val.requires_grad or B.val.requires_grad):
        # device matrices: hash-table SpGEMM of this framework (csrc/cuda/structure_build.cu), no cuSPARSE
        from ..ops.cuda_struct import cuda_spgemm
        ap, ai, ae = A.csr()
        bp, bi, be = B.csr()
        av = A.val if ae is None else A.val[ae.long()]
        bv = B.val if be is None else B.val[be.long()]
        r, c, v = cuda_spgemm(ap, ai, av, bp, bi, bv, A.shape[0], B.shape[1])
        return from_coo(r, c, v.to(A.val.dtype), (A.shape[0], B.shape[1]))
    ta, tb = to_torch_sparse_coo(A), to_torch_sparse_coo(B)
    tc = torch.sparse.mm(ta, tb).coalesce()
    idx = tc.indices()
    return from_coo(idx[0].to(A.row.dtype), idx[1].to(A.row.dtype), tc.values(), (A.shape[0], B.shape[1]))


def sddmm(A, X1, X2):
    """out.val = A.val * (X1 @ X2)[row, col] computed only at A's non-zeros (g-SDDMM u_dot_v)."""
    if X1.dim() == 1:
        X1 = X1.view(-1, 1)
    if X2.dim() == 1:
        X2 = X2.view(1, -1)
    # graph src = col, dst = row:  lhs on v (rows) = X1, rhs on u (cols) = X2^T
    dots = SA.gsddmm(A._g, "dot", X1, X2.t().contiguous(), "v", "u").squeeze(-1)
    return val_like(A, A.val * dots)


def bsddmm(A, X1, X2):
    """Batched sddmm: X1 [M, K, B], X2 [K, N, B]; A.val is [nnz] (shared by the batch) or [nnz, B]; result val [nnz, B]."""
    x1 = X1.permute(0, 2, 1).contiguous()  # [M, B, K]
    x2 = X2.permute(1, 2, 0).contiguous()  # [N, B, K]
    dots = SA.gsddmm(A._g, "dot", x1, x2, "v", "u").squeeze(-1)  # [nnz, B]
    scale = A.val if A.val.dim() > 1 else A.val.unsqueeze(-1)
    return val_like(A, scale * dots)


def matmul(A, B):
    if isinstance(A, SparseMatrix) and isinstance(B, SparseMatrix):
        return spspmm(A, B)
    if isinstance(A, SparseMatrix) and torch.is_tensor(B):
        return spmm(A, B)
    if torch.is_tensor(A) and isinstance(B, SparseMatrix):
        return spmm(B.t(), A.t()).t()
    raise DGLError("unsupported operand types for matmul")


# --------------------------------------------------------------------------------------------- softmax / reduce
def softmax(A, dim=1):
    """Softmax over the non-zeros of each row (dim=1) or column (dim=0); fused edge-softmax kernel."""
    v = A.val if A.val.dim() > 1 else A.val.unsqueeze(-1)
    out = SA.edge_softmax(A._g, v, None, "dst" if dim == 1 else "src")
    return val_like(A, out.view_as(A.val))


def _reduce(A, dim, op):
    v = A.val
    if dim is None:
        if op == "sum":
            return v.sum(0)
        if op == "max":
            return v.max(0)[0]
        if op == "min":
            return v.min(0)[0]
        if op == "mean":
            return v.mean(0)
        if op == "prod":
            return v.prod(0)
    r, c = A.coo()
    key = (r if dim == 1 else c).long()
    n = A.shape[0] if dim == 1 else A.shape[1]
    shape = (n,) + tuple(v.shape[1:])
    idx = key.view(-1, *([1] * (v.dim() - 1))).expand_as(v)
    if op in ("sum", "mean"):
        out = torch.zeros(shape, dtype=v.dtype, device=v.device).index_add(0, key, v)
        if op == "mean":
            cnt = torch.bincount(key, minlength=n).clamp(min=1).to(v.dtype)
            out = out / cnt.view(-1, *([1] * (v.dim() - 1)))
        return out
    if op == "prod":
        return torch.ones(shape, dtype=v.dtype, device=v.device).scatter_reduce(0, idx, v, reduce="prod")
    red = "amax" if op == "max" else "amin"
    out = torch.zeros(shape, dtype=v.dtype, device=v.device).scatter_reduce(0, idx, v, reduce=red, include_self=False)
    return out


def reduce(A, dim=None, rtype="sum"):  # noqa: A001
    """Reduce the non-zeros along ``dim`` with ``rtype`` in sum / smin / smax / smean / sprod (reference
    sparse/reduction.py:reduce)."""
    op = {"sum": "sum", "smin": "min", "smax": "max", "smean": "mean", "sprod": "prod",
          "min": "min", "max": "max", "mean": "mean", "prod": "prod"}.get(rtype)
    if op is None:
        raise ValueError(f"unknown reduce type {rtype}")
    return _reduce(A, dim, op)


def sum(A, dim=None):  # noqa: A001
    return _reduce(A, dim, "sum")


def smax(A, dim=None):
    return _reduce(A, dim, "max")


def smin(A, dim=None):
    return _reduce(A, dim, "min")


def smean(A, dim=None):
    return _reduce(A, dim, "mean")


def sprod(A, dim=None):
    return _reduce(A, dim, "prod")


# --------------------------------------------------------------------------------------------- elementwise
def _binary_sp(A, B, fn):
    if A.shape != B.shape:
        raise DGLError(f"shape mismatch: {A.shape} vs {B.shape}")
    if _cpu_struct.usable(A, B) and not A.has_duplicate() and not B.has_duplicate():
        # host matrices: per-row sorted merge in native code instead of concat + unique
        cp, ci, cv = _cpu_struct.spadd(A, B, beta=float(fn(1.0)))
        return from_csr(cp, ci, cv, A.shape)
    ra, ca = A.coo()
    rb, cb = B.coo()
    r, c = torch.cat([ra, rb.to(ra.dtype)]), torch.cat([ca, cb.to(ca.dtype)])
    v = torch.cat([A.val, fn(B.val)])
    return from_coo(r, c, v, A.shape).coalesce()


def add(A, B):
    if isinstance(A, SparseMatrix) and isinstance(B, SparseMatrix):
        return _binary_sp(A, B, lambda x: x)
    raise DGLError("SparseMatrix can only be added to another SparseMatrix")


def sub(A, B):
    if isinstance(A, SparseMatrix) and isinstance(B, SparseMatrix):
        return _binary_sp(A, B, operator.neg)
    raise DGLError("SparseMatrix can only be subtracted by another SparseMatrix")


def is_scalar(x):
    """Python number or 0-dim tensor (reference sparse/utils.py)."""
    import numbers
    return isinstance(x, numbers.Number) or (torch.is_tensor(x) and x.dim() == 0)


def _match_nonzeros(A, B):
    """Positions (ia, ib) of the coordinates that A and B share (both coalesced); order follows A's non-zeros."""
    ra, ca = A.coo()
    rb, cb = B.coo()
    ncol = max(A.shape[1], 1)
    ka = ra.long() * ncol + ca.long()
    kb = rb.long() * ncol + cb.long()
    if kb.numel() == 0 or ka.numel() == 0:
        e = torch.empty(0, dtype=torch.long, device=ka.device)
        return e, e
    skb, order = torch.sort(kb)
    pos = torch.searchsorted(skb, ka).clamp(max=skb.numel() - 1)
    hit = skb[pos] == ka
    ia = torch.nonzero(hit).view(-1)
    return ia, order[pos[hit]]


def spsp_add(A, B):
    return _binary_sp(A, B, lambda x: x)


def spsp_mul(A, B):
    """Elementwise product of two sparse matrices: non-zero only where both are (reference elementwise_op_sp.py:17)."""
    if A.shape != B.shape:
        raise DGLError(f"shape mismatch: {A.shape} vs {B.shape}")
    A = A.coalesce() if A.has_duplicate() else A
    B = B.coalesce() if B.has_duplicate() else B
    ia, ib = _match_nonzeros(A, B)
    r, c = A.coo()
    return from_coo(r[ia], c[ia], A.val[ia] * B.val[ib], A.shape)


def spsp_div(A, B):
    """Elementwise quotient; both operands must have the same sparsity and the result keeps A's non-zero order
    (reference elementwise_op_sp.py:24)."""
    if A.shape != B.shape:
        raise DGLError(f"shape mismatch: {A.shape} vs {B.shape}")
    if A.has_duplicate() or B.has_duplicate():
        raise DGLError("sp_div requires both matrices to be coalesced (no duplicate entries)")
    ia, ib = _match_nonzeros(A, B)
    if not (ia.numel() == A.nnz == B.nnz):
        raise DGLError("Cannot divide two COO matrices with different sparsities.")
    return val_like(A, A.val / B.val[ib])


def mul(A, B):
    if isinstance(A, SparseMatrix) and isinstance(B, SparseMatrix):
        if A.is_diag() and B.is_diag() and A.nnz == B.nnz and A.shape == B.shape and not A.has_duplicate():
            ra, _ = A.coo()
            rb, _ = B.coo()
            if torch.equal(ra, rb):
                return val_like(A, A.val * B.val)
        return spsp_mul(A, B)
    if isinstance(B, SparseMatrix):
        A, B = B, A
    return val_like(A, A.val * B)


def div(A, B):
    if isinstance(B, SparseMatrix):
        if not isinstance(A, SparseMatrix):
            raise DGLError("Only a sparse matrix can be divided by a sparse matrix.")
        return spsp_div(A, B)
    return val_like(A, A.val / B)


def sp_add(A, B):
    """``A + B`` for two sparse matrices (operator form; other operand types are rejected)."""
    return spsp_add(A, B) if isinstance(B, SparseMatrix) else NotImplemented


def sp_sub(A, B):
    return sub(A, B) if isinstance(B, SparseMatrix) else NotImplemented


def sp_mul(A, B):
    """``A * B``: scalar scaling, or the intersection product of two sparse matrices."""
    if is_scalar(B):
        return val_like(A, A.val * B)
    return spsp_mul(A, B) if isinstance(B, SparseMatrix) else NotImplemented


def sp_div(A, B):
    """``A / B``: scalar, or a sparse matrix with identical sparsity."""
    if is_scalar(B):
        return val_like(A, A.val / B)
    return spsp_div(A, B) if isinstance(B, SparseMatrix) else NotImplemented


def sp_power(A, scalar):
    return val_like(A, A.val ** scalar) if is_scalar(scalar) else NotImplemented


def load_dgl_sparse():
    """The reference dlopens libdgl_sparse here; the sparse kernels of this package live in the two native libraries
    loaded by ``dgl_b200._native.loader`` — make sure the host one is present."""
    from .._native import loader
    loader.cpu_lib()


def power(A, scalar):
    return val_like(A, A.val ** scalar)


def neg(A):
    return val_like(A, -A.val)


def inv(A):
    if not A.is_diag() or A.shape[0] != A.shape[1]:
        raise DGLError("inv is only defined for square diagonal matrices")
    return val_like(A, 1.0 / A.val)


def sp_broadcast_v(A, v, op):
    """Broadcast a vector along rows (v: [1, N] or [N]) or columns (v: [M, 1]) onto the non-zeros."""
    r, c = A.coo()
    if v.dim() == 2 and v.shape[1] == 1:
        vv = v.view(-1)[r.long()]
    else:
        vv = v.view(-1)[c.long()]
    f = {"add": operator.add, "sub": operator.sub, "mul": operator.mul, "truediv": operator.truediv,
         "div": operator.truediv}[op]
    return val_like(A, f(A.val, vv if A.val.dim() == 1 else vv.unsqueeze(-1)))


def sp_add_v(A, v):
    return sp_broadcast_v(A, v, "add")


def sp_sub_v(A, v):
    return sp_broadcast_v(A, v, "sub")


def sp_mul_v(A, v):
    return sp_broadcast_v(A, v, "mul")


def sp_div_v(A, v):
    return sp_broadcast_v(A, v, "div")
