"""Typed / heterogeneous linear layers.

``TypedLinear`` maps row i with its own type's matrix, ``y_i = x_i W[t_i]``. Execution plans (all device-side, no
``.cpu()`` on the type counts as in /root/reference/python/dgl/nn/pytorch/linear.py:205-207):

* rows sorted by type  -> ONE grouped tcgen05 GEMM over the segments (``ops.segment_mm`` -> grouped_mm_tc.cu);
* unsorted rows        -> gathered-A grouped GEMM (``ops.gather_mm``);
* ``basis`` regulariser -> the type matrices are never materialised when there are fewer bases than types:
  ``y = sum_b coeff[t, b] * (x V_b)`` is one dense GEMM against the stacked bases + a per-row contraction;
* ``bdd`` regulariser   -> block-diagonal matrices act as a batched small GEMM on the reshaped rows.

State-dict layout (``W``, ``coeff``) matches the reference so checkpoints interchange.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..ops import gather_mm, segment_mm

_KINDS = {None: "full", "basis": "basis", "bdd": "bdd"}


class TypedLinear(nn.Module):
    r"""Linear map selected per row by an integer type id, with optional basis / block-diagonal sharing."""

    def __init__(self, in_size, out_size, num_types, regularizer=None, num_bases=None):
        super().__init__()
        if regularizer not in _KINDS:
            raise ValueError(f'Supported regularizer options: "basis", "bdd", but got {regularizer}')
        if regularizer is not None and num_bases is None:
            raise ValueError(f'Missing "num_bases" for {regularizer} regularization.')
        self.in_size, self.out_size, self.num_types = in_size, out_size, num_types
        self.regularizer = regularizer
        kind = _KINDS[regularizer]
        if kind == "full":
            w_shape = (num_types, in_size, out_size)
        elif kind == "basis":
            w_shape = (num_bases, in_size, out_size)
            self.coeff = nn.Parameter(torch.empty(num_types, num_bases))
            self.num_bases = num_bases
        else:
            if in_size % num_bases or out_size % num_bases:
                raise ValueError("Input and output sizes must be divisible by num_bases.")
            self.submat_in, self.submat_out = in_size // num_bases, out_size // num_bases
            w_shape = (num_types, num_bases * self.submat_in * self.submat_out)
            self.num_bases = num_bases
        self.W = nn.Parameter(torch.empty(*w_shape))
        self.reset_parameters()

    def reset_parameters(self):
        fan = self.submat_in if self.regularizer == "bdd" else self.in_size
        bound = 1.0 / math.sqrt(fan)
        with torch.no_grad():
            self.W.uniform_(-bound, bound)
            if self.regularizer == "basis":
                nn.init.xavier_uniform_(self.coeff, gain=nn.init.calculate_gain("relu"))

    def get_weight(self):
        """Per-type matrices [num_types, in, out] (for ``bdd``: the packed block parameters)."""
        if self.regularizer == "basis":
            return torch.einsum("tb,bio->tio", self.coeff, self.W)
        return self.W

    # ---- execution plans
    def _apply_blocks(self, x, x_type):
        blocks = self.W[x_type.long()].view(-1, self.submat_in, self.submat_out)     # [n * B, in/B, out/B]
        return torch.bmm(x.reshape(-1, 1, self.submat_in), blocks).view(-1, self.out_size)

    def _apply_bases(self, x, x_type):
        # x [n, in] @ [in, B * out] -> [n, B, out]; contract with the row's coefficients
        stacked = self.W.permute(1, 0, 2).reshape(self.in_size, self.num_bases * self.out_size)
        per_base = (x @ stacked).view(-1, self.num_bases, self.out_size)
        return torch.einsum("nb,nbo->no", self.coeff[x_type.long()].to(per_base.dtype), per_base)

    def forward(self, x, x_type, sorted_by_type=False):
        if self.regularizer == "bdd":
            return self._apply_blocks(x, x_type)
        if self.regularizer == "basis" and self.num_bases < self.num_types:
            return self._apply_bases(x, x_type)
        w = self.get_weight()
        if sorted_by_type:
            return segment_mm(x, w, seglen_a=torch.bincount(x_type.long(), minlength=self.num_types))
        return gather_mm(x, w, idx_b=x_type)

    def extra_repr(self):
        s = f"in_size={self.in_size}, out_size={self.out_size}, num_types={self.num_types}"
        if self.regularizer is not None:
            s += f", regularizer={self.regularizer}, num_bases={self.num_bases}"
        return s


class _PerKey(nn.Module):
    """A dict of sub-modules addressed by arbitrary hashable keys (ModuleDict wants strings)."""

    def __init__(self, factory, spec):
        super().__init__()
        self._names = {}
        store = nn.ModuleDict()
        for key, arg in spec.items():
            name = str(key)
            store[name] = factory(arg)
            self._names[name] = key
        self._store_attr(store)

    def _store_attr(self, store):  # pragma: no cover - overridden to pick the reference's attribute name
        raise NotImplementedError

    def _run(self, store, inputs):
        return {key: store[str(key)](val) for key, val in inputs.items()}


class HeteroLinear(_PerKey):
    """One ``nn.Linear(in_size[key], out_size)`` per key of the input dict."""

    def __init__(self, in_size, out_size, bias=True):
        super().__init__(lambda width: nn.Linear(width, out_size, bias=bias), in_size)

    def _store_attr(self, store):
        self.linears = store

    def forward(self, feat):
        return self._run(self.linears, feat)


class HeteroEmbedding(_PerKey):
    """One ``nn.Embedding(num_embeddings[key], embedding_dim)`` per key; ``weight`` returns them keyed like the input."""

    def __init__(self, num_embeddings, embedding_dim):
        super().__init__(lambda rows: nn.Embedding(rows, embedding_dim), num_embeddings)
        self.raw_keys = dict(self._names)

    def _store_attr(self, store):
        self.embeds = store

    @property
    def weight(self):
        return {self._names[name]: table.weight for name, table in self.embeds.items()}

    def reset_parameters(self):
        for table in self.embeds.values():
            nn.init.xavier_uniform_(table.weight)

    def forward(self, input_ids):
        return self._run(self.embeds, input_ids)
