"""Columnar feature storage.

Parity: /root/reference/python/dgl/frame.py (Column/Frame/LazyFeature). A
``Frame`` maps names to tensors whose first dimension equals ``num_rows``.
Sub-frames created by indexing are *lazy*: a column stores ``(base, index)`` and
materialises on first access, so slicing a graph does not copy every feature.
"""
from __future__ import annotations

from collections.abc import MutableMapping

import torch

from .base import DGLError


class Scheme(tuple):
    """(shape, dtype) of one column: ``shape`` excludes the row dimension (reference frame.py:125). A plain tuple
    subclass so it compares equal to the ``(shape, dtype)`` pairs of :attr:`Frame.schemes` and pickles by value."""
    __slots__ = ()

    def __new__(cls, shape, dtype):
        return super().__new__(cls, (tuple(shape), dtype))

    @property
    def shape(self):
        return self[0]

    @property
    def dtype(self):
        return self[1]

    def __reduce__(self):
        return Scheme, (self[0], self[1])

    def __repr__(self):
        return f"Scheme(shape={self[0]}, dtype={self[1]})"


def infer_scheme(tensor):
    """Scheme of the column a tensor would make."""
    return Scheme(tuple(tensor.shape[1:]), tensor.dtype)


class LazyFeature:
    """Placeholder for a feature fetched later by the DataLoader
    (reference: python/dgl/frame.py:class LazyFeature)."""

    __slots__ = ("name", "id_")

    def __init__(self, name=None, id_=None):
        self.name = name
        self.id_ = id_

    def to(self, *args, **kwargs):
        return self

    @property
    def data(self):
        return self

    def pin_memory_(self):
        pass

    def unpin_memory_(self):
        pass

    def record_stream(self, stream):
        pass


class Column:
    """A tensor column with optional lazy row index and lazy device move."""

    __slots__ = ("storage", "index", "device", "_pinned")

    def __init__(self, storage, index=None, device=None):
        self.storage = storage
        self.index = index
        self.device = device
        self._pinned = False

    def __len__(self):
        if self.index is None:
            return self.storage.shape[0]
        return self.index.shape[0]

    @property
    def shape(self):
        return (len(self),) + tuple(self.storage.shape[1:])

    @property
    def dtype(self):
        return self.storage.dtype

    @property
    def data(self):
        if isinstance(self.storage, LazyFeature):
            return self.storage
        if self.index is not None:
            idx = self.index
            if idx.device != self.storage.device:
                if self.storage.is_cuda or not (self.storage.is_pinned() and idx.is_cuda):
                    idx = idx.to(self.storage.device)
            if idx.is_cuda and not self.storage.is_cuda:
                # UVA gather from pinned host memory
                from .utils.pin_memory import gather_pinned_tensor_rows
                self.storage = gather_pinned_tensor_rows(self.storage, idx)
            else:
                self.storage = self.storage[idx.long()]
            self.index = None
        if self.device is not None:
            self.storage = self.storage.to(self.device[0], **self.device[1])
            self.device = None
        return self.storage

    def to(self, device, **kwargs):
        c = Column(self.storage, self.index, (torch.device(device), kwargs))
        return c

    def subcolumn(self, rowids):
        if self.index is None:
            return Column(self.storage, rowids, self.device)
        idx = self.index
        if idx.device != rowids.device:
            rowids = rowids.to(idx.device)
        return Column(self.storage, idx[rowids.long()], self.device)

    def clone(self):
        return Column(self.storage, self.index, self.device)

    # ---- reference Column surface (python/dgl/frame.py:306-560) --------------------------------
    def astype(self, new_dtype):
        """Column of another dtype; the lazy index / device move are kept lazy."""
        if new_dtype == self.dtype:
            return self
        return Column(self.storage.to(new_dtype), self.index, self.device)

    def __getitem__(self, rowids):
        return self.data[rowids.long() if torch.is_tensor(rowids) and rowids.dtype != torch.bool else rowids]

    def __setitem__(self, rowids, feats):
        self.update(rowids, feats)

    def update(self, rowids, feats):
        """Write rows; the column is materialised and copied first so views sharing the storage are unaffected."""
        data = self.data.clone()
        data[rowids.long() if torch.is_tensor(rowids) else rowids] = feats.to(data.dtype).to(data.device)
        self.storage = data

    def extend(self, feats, feat_scheme=None):
        """Append rows."""
        feats = torch.as_tensor(feats)
        if tuple(feats.shape[1:]) != tuple(self.storage.shape[1:]):
            raise DGLError(f"Cannot append rows of shape {tuple(feats.shape[1:])} to a column of shape "
                           f"{tuple(self.storage.shape[1:])}.")
        data = self.data
        self.storage = torch.cat([data, feats.to(data.dtype).to(data.device)], 0)

    def deepclone(self):
        return Column(self.data.clone())

    @staticmethod
    def create(data):
        """Wrap a tensor (or pass a Column through)."""
        return data.clone() if isinstance(data, Column) else Column(torch.as_tensor(data))

    def fetch(self, indices, device, pin_memory=False, **kwargs):
        """Rows ``indices`` on ``device`` (the FeatureStorage protocol the dataloader uses); pinned host storage with a
        CUDA target goes through the UVA gather."""
        sub = self.subcolumn(torch.as_tensor(indices))
        out = sub.data
        if pin_memory and not out.is_cuda and not out.is_pinned():
            out = out.pin_memory()
        return out.to(device, **kwargs)

    def pin_memory_(self):
        """Page-lock the materialised storage in place (remembered so :meth:`unpin_memory_` only undoes our own pin)."""
        from .utils.pin_memory import pin_memory_inplace
        data = self.data
        if not data.is_cuda and not data.is_pinned():
            pin_memory_inplace(data)
            self._pinned = True

    def unpin_memory_(self):
        from .utils.pin_memory import unpin_memory_inplace
        if self._pinned and self.index is None and self.device is None:
            unpin_memory_inplace(self.storage)
            self._pinned = False

    def record_stream(self, stream):
        if torch.is_tensor(self.storage) and self.storage.is_cuda:
            self.storage.record_stream(stream)
        if self.index is not None and self.index.is_cuda:
            self.index.record_stream(stream)

    def __repr__(self):
        return repr(self.data)

    def __getstate__(self):
        # materialise: a pickled column never drags the full parent storage along
        return self.data

    def __setstate__(self, state):
        self.storage, self.index, self.device, self._pinned = state, None, None, False

    def __copy__(self):
        return self.clone()


class Frame(MutableMapping):
    """name -> Column with a common row count."""

    def __init__(self, data=None, num_rows=None):
        self._columns = {}
        self._num_rows = num_rows
        self._initializers = {}
        self._default_initializer = None
        if data is not None:
            for k, v in data.items():
                col = v if isinstance(v, Column) else Column(v)
                if not isinstance(col.storage, LazyFeature):
                    n = len(col)
                    if self._num_rows is None:
                        self._num_rows = n
                    elif n != self._num_rows:
                        raise DGLError(f"Expected all columns to have {self._num_rows} rows, got {n} on column '{k}'")
                self._columns[k] = col
        if self._num_rows is None:
            self._num_rows = 0

    # mapping protocol
    def __getitem__(self, name):
        return self._columns[name].data

    def __setitem__(self, name, data):
        self.update_column(name, data)

    def __delitem__(self, name):
        del self._columns[name]

    def __iter__(self):
        return iter(self._columns)

    def __len__(self):
        return len(self._columns)

    def __contains__(self, name):
        return name in self._columns

    @property
    def num_rows(self):
        return self._num_rows

    @property
    def schemes(self):
        return {k: Scheme(tuple(c.shape[1:]), c.dtype) for k, c in self._columns.items()
                if not isinstance(c.storage, LazyFeature)}

    def set_initializer(self, initializer, column=None):
        if column is None:
            self._default_initializer = initializer
        else:
            self._initializers[column] = initializer

    def get_initializer(self, column=None):
        return self._initializers.get(column, self._default_initializer)

    def update_column(self, name, data):
        if isinstance(data, LazyFeature):
            self._columns[name] = Column(data)
            return
        if isinstance(data, Column):
            col = data
        else:
            if not torch.is_tensor(data):
                data = torch.as_tensor(data)
            col = Column(data)
        if len(col) != self._num_rows:
            raise DGLError(f"Expect number of features to match number of rows ({self._num_rows}). "
                           f"Got {len(col)} on column '{name}'.")
        self._columns[name] = col

    def update_row(self, rowids, data):
        """Scatter rows of ``data`` (dict) into the frame; new columns are
        created with the initializer (zeros by default)."""
        for k, v in data.items():
            if k not in self._columns:
                init = self.get_initializer(k)
                shape = (self._num_rows,) + tuple(v.shape[1:])
                if init is None:
                    base = torch.zeros(shape, dtype=v.dtype, device=v.device)
                else:
                    base = init(shape, v.dtype, v.device, slice(0, self._num_rows))
                self._columns[k] = Column(base)
            cur = self._columns[k].data
            new = cur.clone() if not cur.requires_grad else cur.clone()
            idx = rowids.long().to(new.device)
            new = new.index_copy(0, idx, v.to(new.dtype)) if v.requires_grad or new.requires_grad else _index_copy_(new, idx, v)
            self._columns[k] = Column(new)

    def add_rows(self, num_rows):
        for k, c in list(self._columns.items()):
            cur = c.data
            init = self.get_initializer(k)
            shape = (num_rows,) + tuple(cur.shape[1:])
            if init is None:
                ext = torch.zeros(shape, dtype=cur.dtype, device=cur.device)
            else:
                ext = init(shape, cur.dtype, cur.device, slice(self._num_rows, self._num_rows + num_rows))
            self._columns[k] = Column(torch.cat([cur, ext], 0))
        self._num_rows += num_rows

    def append(self, data):
        """Append rows given as dict of tensors (missing columns are filled)."""
        n = None
        for v in data.values():
            n = v.shape[0]
            break
        if n is None:
            return
        old = self._num_rows
        self.add_rows(n)
        ids = torch.arange(old, old + n)
        self.update_row(ids, {k: v for k, v in data.items()})

    def subframe(self, rowids):
        cols = {k: c.subcolumn(rowids) for k, c in self._columns.items()}
        f = Frame(num_rows=int(rowids.shape[0]))
        f._columns = cols
        f._initializers = dict(self._initializers)
        f._default_initializer = self._default_initializer
        return f

    # ---- reference frame.py helpers -------------------------------------------------------------------------
    @property
    def num_columns(self):
        return len(self._columns)

    def add_column(self, name, scheme=None, ctx=None):
        """Add an (initialiser-filled) column (reference frame.py add_column)."""
        if name in self._columns:
            return
        self._append_missing_or_init(name, scheme, ctx) if hasattr(self, "_append_missing_or_init") else None

    def _cast(self, dtype):
        out = self.clone()
        for k in list(out.keys()):
            v = out[k]
            if torch.is_tensor(v) and v.is_floating_point():
                out[k] = v.to(dtype)
        return out

    def half(self):
        return self._cast(torch.float16)

    def bfloat16(self):
        return self._cast(torch.bfloat16)

    def float(self):
        return self._cast(torch.float32)

    def double(self):
        return self._cast(torch.float64)

    def deepclone(self):
        """Clone with copied storage (reference frame.py deepclone)."""
        out = self.clone()
        for k in list(out.keys()):
            v = out[k]
            if torch.is_tensor(v):
                out[k] = v.clone()
        return out

    def clone(self):
        f = Frame(num_rows=self._num_rows)
        f._columns = {k: c.clone() for k, c in self._columns.items()}
        f._initializers = dict(self._initializers)
        f._default_initializer = self._default_initializer
        return f

    def to(self, device, **kwargs):
        f = Frame(num_rows=self._num_rows)
        f._columns = {k: c.to(device, **kwargs) for k, c in self._columns.items()}
        f._initializers = dict(self._initializers)
        f._default_initializer = self._default_initializer
        return f

    def pin_memory_(self):
        from .utils.pin_memory import pin_memory_inplace
        for c in self._columns.values():
            d = c.data
            if torch.is_tensor(d) and not d.is_cuda:
                pin_memory_inplace(d)

    def unpin_memory_(self):
        from .utils.pin_memory import unpin_memory_inplace
        for c in self._columns.values():
            d = c.data
            if torch.is_tensor(d) and not d.is_cuda:
                unpin_memory_inplace(d)

    def record_stream(self, stream):
        for c in self._columns.values():
            d = c.data
            if torch.is_tensor(d) and d.is_cuda:
                d.record_stream(stream)

    def __repr__(self):
        return f"Frame(num_rows={self._num_rows}, columns={list(self._columns)})"


def _index_copy_(dst, idx, src):
    dst[idx] = src.to(dst.dtype)
    return dst
