"""``dgl.backend`` (parity: /root/reference/python/dgl/backend/__init__.py, backend.py, pytorch/tensor.py).

The reference abstracts over PyTorch / MXNet / TensorFlow and selects one with ``DGLBACKEND``. This framework is
PyTorch-only by design (the other backends are out of scope), so the module exposes the same tensor
helper names implemented directly on torch - enough for user code written against ``dgl.backend as F``."""
from __future__ import annotations

import os

import numpy as np
import torch

backend_name = "pytorch"
if os.environ.get("DGLBACKEND", "pytorch").lower() not in ("pytorch", ""):
    raise ImportError(f"dgl_b200 supports the pytorch backend only (DGLBACKEND={os.environ['DGLBACKEND']!r})")

float16, bfloat16, float32, float64 = torch.float16, torch.bfloat16, torch.float32, torch.float64
uint8, int8, int16, int32, int64, bool = torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64, torch.bool  # noqa: A001
data_type_dict = {"float16": float16, "bfloat16": bfloat16, "float32": float32, "float64": float64, "uint8": uint8,
                  "int8": int8, "int16": int16, "int32": int32, "int64": int64, "bool": bool}


def get_preferred_backend():
    return "pytorch"


def set_default_backend(default_dir, backend_name="pytorch"):  # noqa: ARG001 - kept for API parity
    if backend_name != "pytorch":
        raise ValueError("dgl_b200 supports the pytorch backend only")


def cpu():
    return torch.device("cpu")


def cuda():
    return torch.device("cuda", torch.cuda.current_device())


def is_tensor(obj):
    return torch.is_tensor(obj)


def tensor(data, dtype=None):
    return torch.as_tensor(data, dtype=dtype)


def as_scalar(data):
    return data.item()


def shape(x):
    return x.shape


def dtype(x):
    return x.dtype


def ndim(x):
    return x.dim()


def context(x):
    return x.device


def device_type(ctx):
    return torch.device(ctx).type


def device_id(ctx):
    d = torch.device(ctx)
    return 0 if d.index is None else d.index


def to_backend_ctx(ctx):
    return torch.device(ctx)


def astype(x, ty):
    return x.to(ty)


def asnumpy(x):
    return x.detach().cpu().numpy()


def zerocopy_to_numpy(x):
    return x.detach().cpu().numpy()


def zerocopy_from_numpy(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def copy_to(x, ctx, **kwargs):
    return x.to(ctx, **kwargs)


def is_pinned(x):
    return x.is_pinned()


def sum(x, dim, keepdims=False):  # noqa: A001
    return x.sum(dim, keepdim=keepdims)


def reduce_sum(x):
    return x.sum()


def mean(x, dim):
    return x.mean(dim)


def reduce_mean(x):
    return x.mean()


def max(x, dim):  # noqa: A001
    return x.max(dim)[0]


def reduce_max(x):
    return x.max()


def min(x, dim):  # noqa: A001
    return x.min(dim)[0]


def reduce_min(x):
    return x.min()


def argsort(x, dim, descending):
    return torch.argsort(x, dim=dim, descending=descending)


def topk(x, k, dim, descending=True):
    return torch.topk(x, k, dim, largest=descending)[0]


def argtopk(x, k, dim, descending=True):
    return torch.topk(x, k, dim, largest=descending)[1]


def exp(x):
    return torch.exp(x)


def sqrt(x):
    return torch.sqrt(x)


def softmax(x, dim):
    return torch.softmax(x, dim)


def cat(seq, dim):
    return torch.cat(seq, dim)


def stack(seq, dim):
    return torch.stack(seq, dim)


def split(x, sizes_or_sections, dim):
    return torch.split(x, sizes_or_sections, dim)


def repeat(x, repeats, dim):
    return torch.repeat_interleave(x, repeats, dim)


def gather_row(data, row_index):
    return torch.index_select(data, 0, row_index.long())


def slice_axis(data, axis, begin, end):
    return torch.narrow(data, axis, begin, end - begin)


def take(data, indices, dim):
    return torch.index_select(data, dim, indices.reshape(-1).long()).view(data.shape[:dim] + indices.shape + data.shape[dim + 1:])


def narrow_row(x, start, stop):
    return x[start:stop]


def index_add_inplace(data, row_idx, value):
    data.index_add_(0, row_idx.long(), value)


def scatter_row(data, row_index, value):
    return data.index_copy(0, row_index.long(), value)


def scatter_row_inplace(data, row_index, value):
    data[row_index.long()] = value


def squeeze(x, dim):
    return torch.squeeze(x, dim)


def unsqueeze(x, dim):
    return torch.unsqueeze(x, dim)


def reshape(x, shape):  # noqa: A002
    return torch.reshape(x, shape)


def swapaxes(x, a, b):
    return torch.transpose(x, a, b)


def zeros(shape, dtype, ctx):  # noqa: A002
    return torch.zeros(shape, dtype=dtype, device=ctx)


def zeros_like(x):
    return torch.zeros_like(x)


def ones(shape, dtype, ctx):  # noqa: A002
    return torch.ones(shape, dtype=dtype, device=ctx)


def uniform(shape, dtype, ctx, low, high):  # noqa: A002
    return torch.empty(shape, dtype=dtype, device=ctx).uniform_(low, high)


def randint(shape, dtype, ctx, low, high):  # noqa: A002
    return torch.randint(low, high, shape, dtype=dtype, device=ctx)


def pad_packed_tensor(x, lengths, value, l_min=None):
    lengths = torch.as_tensor(lengths)
    max_len = int(lengths.max()) if l_min is None else builtins_max(int(lengths.max()), l_min)
    out = x.new_full((len(lengths), max_len) + tuple(x.shape[1:]), value)
    off = 0
    for i, n in enumerate(lengths.tolist()):
        out[i, :n] = x[off: off + n]
        off += n
    return out


def pack_padded_tensor(x, lengths):
    lengths = torch.as_tensor(lengths).tolist()
    return torch.cat([x[i, :n] for i, n in enumerate(lengths)], 0)


def boolean_mask(x, mask):
    return x[mask.bool()]


def equal(x, y):
    return x == y


def allclose(x, y, rtol=1e-4, atol=1e-4):
    return torch.allclose(x, y, rtol=rtol, atol=atol)


def logical_not(x):
    return ~x.bool()


def logical_and(x, y):
    return x & y


def clone(x):
    return x.clone()


def clamp(x, min_val, max_val):
    return torch.clamp(x, min_val, max_val)


def replace_inf_with_zero(x):
    return torch.masked_fill(x, torch.isinf(x), 0)


def count_nonzero(x):
    return int((x != 0).sum())


def unique(x, return_inverse=False, return_counts=False):
    return torch.unique(x, return_inverse=return_inverse, return_counts=return_counts)


def full_1d(length, fill_value, dtype, ctx):  # noqa: A002
    return torch.full((length,), fill_value, dtype=dtype, device=ctx)


def nonzero_1d(x):
    return torch.nonzero(x, as_tuple=False).squeeze(1)


def sort_1d(x):
    return torch.sort(x)


def arange(start, stop, dtype=torch.int64, ctx=None):  # noqa: A002
    return torch.arange(start, stop, dtype=dtype, device=ctx)


def rand_shuffle(arr):
    return arr[torch.randperm(len(arr))]


def zerocopy_to_dlpack(x):
    return torch.utils.dlpack.to_dlpack(x.contiguous())


def zerocopy_from_dlpack(dl):
    return torch.utils.dlpack.from_dlpack(dl)


import builtins as _builtins  # noqa: E402

builtins_max = _builtins.max

# sparse autograd ops of the backend (reference backend/pytorch/sparse.py): same entry points, our kernels
from ..ops.sparse_autograd import edge_softmax, gsddmm, gspmm, segment_reduce  # noqa: E402,F401
from ..nn.functional_native import gather_mm, segment_mm  # noqa: E402,F401


# ---- names of the reference's backend loader (backend/__init__.py:51-118); PyTorch is the only backend here -----------
from .._sparse_ops import to_dgl_nd  # noqa: E402,F401


def from_dgl_nd(x):
    """NDArray -> tensor; tensors are the array type of this framework, so this is the identity."""
    return x


def load_backend(mod_name="pytorch"):
    if mod_name != "pytorch":
        raise NotImplementedError(f"backend {mod_name!r}: this framework is PyTorch-only (DGLBACKEND is ignored otherwise)")


def is_enabled(api):
    """Whether the backend provides ``api`` (reference backend/__init__.py:is_enabled)."""
    return api in globals() and callable(globals()[api])


# ---- autograd bookkeeping of the backend API (reference backend/backend.py:1930-2050). PyTorch records eagerly, so the
# ---- recording scope is a no-op context and "attaching" a gradient is requires_grad_ ---------------------------------
def attach_grad(x):
    """Make ``x`` a leaf that accumulates a gradient, clearing any old one."""
    if x.grad is not None:
        x.grad.zero_()
        return x
    return x.requires_grad_()


def backward(x, head_gradient=None):
    """Back-propagate from ``x`` (summed when no head gradient is given and ``x`` is not a scalar)."""
    if head_gradient is not None and head_gradient.shape[0] == 1 and head_gradient.dim() == 1 and x.dim() == 0:
        head_gradient = head_gradient[0]
    x.backward(head_gradient)


def grad(x):
    return x.grad


def is_no_grad(x):
    """True when ``x`` holds no gradient yet or the gradient is identically zero."""
    return x.grad is None or _builtins.bool((x.grad == 0).all())


def is_recording():
    return torch.is_grad_enabled()


class record_grad:
    """Scope in which operations are recorded for differentiation - always the case in eager PyTorch."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


no_grad = torch.no_grad


def sync():
    """Wait for queued device work (reference ``F.sync``); nothing to do on a CPU-only process."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()


# ---- remaining tensor helpers ------------------------------------------------------------------------------------------
def empty(shape, dtype, ctx):  # noqa: A002
    return torch.empty(shape, dtype=dtype, device=ctx)


def cumsum(x, dim):
    return torch.cumsum(x, dim)


def floor_div(x, y):
    return torch.div(x, y, rounding_mode="floor")


def inverse(x):
    return torch.inverse(x)


def get_preferred_sparse_format():
    """Layout ``sparse_matrix`` produces without conversion."""
    return "coo"


def sparse_matrix(data, index, shape, force_format=False):  # noqa: A002, ARG001
    """Backend sparse tensor from ``('coo', [2, nnz] ids)`` or ``('csr', (indices, indptr, edge order))``.
    Returns ``(matrix, permutation-or-None)`` like the reference (backend.py:sparse_matrix); CSR input is expanded to
    COO through the row pointer, so no reordering of ``data`` is ever needed."""
    fmt, payload = index
    if fmt == "coo":
        ids = payload
    elif fmt == "csr":
        indices, indptr = payload[0], payload[1]
        rows = torch.repeat_interleave(torch.arange(len(indptr) - 1, device=indices.device), indptr[1:] - indptr[:-1])
        ids = torch.stack([rows, indices.long()])
    else:
        raise TypeError(f"unsupported sparse index format {fmt!r}")
    return torch.sparse_coo_tensor(ids, data, shape), None


def sparse_matrix_indices(spmat):
    return "coo", spmat.coalesce().indices() if not spmat.is_coalesced() else spmat.indices()


def zerocopy_to_dgl_ndarray(x):
    """Tensors are the array type of the native layer (pointer + stream hand-off), so this is the identity."""
    return x


zerocopy_to_dgl_ndarray_for_write = zerocopy_to_dgl_ndarray


def zerocopy_from_dgl_ndarray(x):
    return x


# graph-level sparse algebra and the typed (heterogeneous) message-passing entry points live in sparse_api
from .sparse_api import (CSRMM, CSRSum, CSRMask, ScatterAdd, binary_reduce, copy_reduce, csrmask, csrmm,  # noqa: E402,F401
                         csrsum, edge_softmax_hetero, gsddmm_hetero, gspmm_hetero, scatter_add)
from ..ops.sparse_autograd import GSDDMM, EdgeSoftmax, GSpMM, SegmentReduce  # noqa: E402,F401
from ..nn.sparse_emb import NodeEmbedding  # noqa: E402,F401
