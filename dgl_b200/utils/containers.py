"""Container and id-relabelling helpers of ``dgl.utils`` (parity: /root/reference/python/dgl/utils/internal.py:
Index, LazyDict/HybridDict/ReadOnlyDict, FlattenedDict, relabel family, cached_member, retry_method_with_fix,
apply_each/recursive_apply_pair, thread-count and NUMA queries)."""
from __future__ import annotations

import functools
import glob
import os
from collections import defaultdict
from collections.abc import Iterable, Mapping, Sequence

import numpy as np
import torch

from ..base import DGLError, dgl_warning

__all__ = ["InconsistentDtypeException", "Index", "toindex", "zero_index", "set_diff", "LazyDict", "HybridDict",
           "ReadOnlyDict", "build_relabel_map", "build_relabel_dict", "CtxCachedObject", "cached_member",
           "is_dict_like", "is_iterable", "is_listlike", "reorder", "reorder_index", "to_dgl_context", "to_nbits_int",
           "make_invmap", "retry_method_with_fix", "group_as_dict", "FlattenedDict", "maybe_flatten_dict", "compensate",
           "relabel", "extract_node_subframes_for_block", "set_num_threads", "get_num_threads", "get_numa_nodes_cores",
           "alias_func", "apply_each", "recursive_apply_pair"]


def is_listlike(data):
    """Sequences that are not strings."""
    return isinstance(data, Sequence) and not isinstance(data, str)


def is_dict_like(obj):
    return isinstance(obj, Mapping)


def is_iterable(obj):
    return isinstance(obj, Iterable)


class InconsistentDtypeException(DGLError):
    """Raised when an index is built from data whose integer width disagrees with the requested one."""

    def __init__(self, msg="", *args, **kwargs):
        super().__init__(f"DGL now requires the input tensor to have the same dtype as the graph index's dtype"
                         f"(which you can get by g.idype). {msg}", *args, **kwargs)


class Index:
    """Immutable integer index backed by one tensor; contiguous ranges are remembered as slices so consumers can take
    the cheap path. Accepts ints, lists, numpy arrays, tensors, slices and other ``Index`` objects."""

    def __init__(self, data, dtype="int64"):
        self.dtype = {"int32": torch.int32, "int64": torch.int64}.get(dtype, dtype)
        self._slice = None
        if isinstance(data, Index):
            self._t, self._slice = data._t.to(self.dtype), data._slice
        elif isinstance(data, slice):
            if data.step not in (None, 1) or data.start is None or data.stop is None:
                raise DGLError("Only contiguous slices with explicit bounds can be used as an index.")
            self._slice = slice(int(data.start), int(data.stop))
            self._t = torch.arange(data.start, data.stop, dtype=self.dtype)
        elif torch.is_tensor(data):
            if data.dtype not in (torch.int32, torch.int64):
                raise DGLError(f"Index data must be an integer tensor, got {data.dtype}.")
            if data.dtype != self.dtype:
                raise InconsistentDtypeException(f"Index data specified as {self.dtype}, but got: {data.dtype}")
            if data.dim() > 1:
                raise DGLError("Index data must be 1D int32/int64 vector, but got shape: %s" % (tuple(data.shape),))
            self._t = data.reshape(-1)
        else:
            arr = np.asarray(data)
            if arr.size and not np.issubdtype(arr.dtype, np.integer):
                raise DGLError(f"Index data must be integers, got {arr.dtype}.")
            self._t = torch.as_tensor(arr.astype(np.int64).reshape(-1)).to(self.dtype)

    # ---- views -----------------------------------------------------------------------------
    def tonumpy(self):
        return self._t.cpu().numpy()

    def tousertensor(self, ctx=None):
        return self._t if ctx is None else self._t.to(ctx)

    def todgltensor(self):
        return self._t.cpu()

    def slice_data(self):
        return self._slice

    def is_slice(self, start, stop):
        return self._slice is not None and self._slice.start == start and self._slice.stop == stop

    def __iter__(self):
        return iter(self._t.tolist())

    def __len__(self):
        return int(self._t.shape[0])

    def __getitem__(self, i):
        return int(self._t[i])

    def __getstate__(self):
        return (self._slice, self.dtype) if self._slice is not None else (self._t.cpu(), self.dtype)

    def __setstate__(self, state):
        data, self.dtype = state
        self._slice = data if isinstance(data, slice) else None
        self._t = torch.arange(data.start, data.stop, dtype=self.dtype) if isinstance(data, slice) else data

    # ---- index algebra ---------------------------------------------------------------------
    def get_items(self, index):
        """``self[index]`` as a new Index; slicing a slice stays a slice."""
        if self._slice is not None and index._slice is not None:
            s = self._slice.start
            return Index(slice(s + index._slice.start, s + index._slice.stop), self.dtype)
        return Index(self._t[index._t.long()], self.dtype)

    def set_items(self, index, value):
        """Copy with ``self[index] = value`` (``value``: int or Index)."""
        out = self._t.clone()
        out[index._t.long()] = value if isinstance(value, int) else value._t.to(out.dtype)
        return Index(out, self.dtype)

    def append_zeros(self, num):
        if num == 0:
            return self
        return Index(torch.cat([self._t, torch.zeros(num, dtype=self.dtype, device=self._t.device)]), self.dtype)

    def nonzero(self):
        return Index(torch.nonzero(self._t).view(-1).to(self.dtype), self.dtype)

    def has_nonzero(self):
        return bool((self._t != 0).any())


def toindex(data, dtype="int64"):
    """Wrap as :class:`Index` (no copy if it already is one of the right width)."""
    if isinstance(data, Index) and data.dtype == {"int32": torch.int32, "int64": torch.int64}.get(dtype, dtype):
        return data
    return Index(data, dtype)


def zero_index(size, dtype="int64"):
    return Index(torch.zeros(size, dtype={"int32": torch.int32, "int64": torch.int64}.get(dtype, dtype)), dtype)


def set_diff(ar1, ar2):
    """Sorted unique ids of ``ar1`` absent from ``ar2``."""
    return toindex(np.setdiff1d(toindex(ar1).tonumpy(), toindex(ar2).tonumpy()))


class LazyDict(Mapping):
    """Read-only mapping whose values are produced on access by ``fn(key)``."""

    def __init__(self, fn, keys):
        self._fn, self._keys = fn, keys

    def __getitem__(self, key):
        if key not in self._keys:
            raise KeyError(key)
        return self._fn(key)

    def __contains__(self, key):
        return key in self._keys

    def __iter__(self):
        return iter(self._keys)

    def __len__(self):
        return len(self._keys)

    def keys(self):
        return self._keys


class HybridDict(Mapping):
    """Read-only union of several mappings; earlier mappings win on key clashes."""

    def __init__(self, *dict_like_list):
        self._parts = dict_like_list
        self._keys = set()
        for d in dict_like_list:
            self._keys.update(d.keys())

    def keys(self):
        return self._keys

    def __getitem__(self, key):
        for d in self._parts:
            if key in d:
                return d[key]
        raise KeyError(key)

    def __contains__(self, key):
        return key in self._keys

    def __iter__(self):
        return iter(self._keys)

    def __len__(self):
        return len(self._keys)


class ReadOnlyDict(Mapping):
    """Read-only view of a mapping."""

    def __init__(self, dict_like):
        self._d = dict_like

    def keys(self):
        return self._d.keys()

    def __getitem__(self, key):
        return self._d[key]

    def __contains__(self, key):
        return key in self._d

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)


def relabel(x):
    """(new_to_old, old_to_new): ids renumbered 0.. in ascending order; ``old_to_new`` has length max(x)+1 with zeros
    at ids that never occur."""
    uniq = torch.unique(x)
    old_to_new = torch.zeros(int(uniq.max()) + 1 if uniq.numel() else 0, dtype=x.dtype, device=x.device)
    old_to_new[uniq.long()] = torch.arange(uniq.shape[0], dtype=x.dtype, device=x.device)
    return uniq, old_to_new


def build_relabel_map(x, is_sorted=False):
    """:func:`relabel` on an :class:`Index` (``is_sorted``: already unique and ascending); the map is int64 on CPU."""
    t = toindex(x).tousertensor()
    uniq = t if is_sorted else torch.unique(t)
    old_to_new = torch.zeros(int(uniq.max()) + 1 if uniq.numel() else 0, dtype=torch.int64)
    old_to_new[uniq.long().cpu()] = torch.arange(uniq.shape[0])
    return uniq, old_to_new


def build_relabel_dict(x):
    """old id -> position in ``x``."""
    return {v: i for i, v in enumerate(x)}


class CtxCachedObject:
    """Memoise ``generator(ctx)`` per device/context."""

    def __init__(self, generator):
        self._generator = generator
        self._by_ctx = {}

    def __call__(self, ctx):
        if ctx not in self._by_ctx:
            self._by_ctx[ctx] = self._generator(ctx)
        return self._by_ctx[ctx]


def cached_member(cache, prefix):
    """Method decorator storing results in the dict attribute ``cache`` keyed by ``prefix`` and the arguments."""
    def deco(func):
        @functools.wraps(func)
        def wrapper(self, *args, **kwargs):
            store = getattr(self, cache)
            key = "%s-%s-%s" % (prefix, "-".join(str(a) for a in args),
                                "-".join(f"{k}:{v}" for k, v in kwargs.items()))
            if key not in store:
                store[key] = func(self, *args, **kwargs)
            return store[key]
        return wrapper
    return deco


def retry_method_with_fix(fix_method):
    """Method decorator: on any exception run ``fix_method(self)`` once and retry; a second failure propagates."""
    def deco(func):
        @functools.wraps(func)
        def wrapper(self, *args, **kwargs):
            try:
                return func(self, *args, **kwargs)
            except Exception:  # noqa: BLE001 - that is the contract
                fix_method(self)
                return func(self, *args, **kwargs)
        return wrapper
    return deco


def reorder(dict_like, index):
    """Every column gathered by ``index``."""
    return {k: v[toindex(index).tousertensor(v.device).long()] for k, v in dict_like.items()}


def reorder_index(idx, order):
    return toindex(toindex(idx).tousertensor()[toindex(order).tousertensor().long()])


def to_dgl_context(ctx):
    """The reference converts to its own DGLContext; tensors here live on torch devices, so this is the identity."""
    return torch.device(ctx)


def to_nbits_int(tensor, nbits):
    assert nbits in (32, 64), "nbits can either be 32 or 64"
    return tensor.to(torch.int32 if nbits == 32 else torch.int64)


def make_invmap(array, use_numpy=True):
    """(uniques, {value: rank}, array re-expressed as ranks)."""
    uniques = np.unique(array) if use_numpy else list(set(array))
    invmap = {x: i for i, x in enumerate(uniques)}
    return uniques, invmap, np.asarray([invmap[x] for x in array])


def group_as_dict(pairs):
    out = defaultdict(list)
    for k, v in pairs:
        out[k].append(v)
    return out


class FlattenedDict:
    """Random access over every ``(group key, item)`` of a dict of sequences, in group order."""

    def __init__(self, groups):
        self._groups = groups
        self._keys = list(groups.keys())
        sizes = [len(groups[k]) for k in self._keys]
        self._offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)

    def __len__(self):
        return int(self._offsets[-1])

    def __iter__(self):
        for k in self._keys:
            for item in self._groups[k]:
                yield k, item

    def __getitem__(self, idx):
        if idx < 0:
            idx += len(self)
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        g = int(np.searchsorted(self._offsets, idx, side="right")) - 1
        k = self._keys[g]
        return k, self._groups[k][int(idx - self._offsets[g])]


def maybe_flatten_dict(data):
    return FlattenedDict(data) if isinstance(data, Mapping) else data


def compensate(ids, origin_ids):
    """Members of the sorted id list ``origin_ids`` that are not in its subset ``ids``."""
    return origin_ids[~torch.isin(origin_ids, ids)].to(ids.dtype)


def extract_node_subframes_for_block(graph, srcnodes, dstnodes):
    """Source-side then destination-side node frames for a block built from ``graph`` (features are index views)."""
    from ..base import NID
    frames = []
    for side in (srcnodes, dstnodes):
        for i, ids in enumerate(side):
            sub = graph._node_frames[i].subframe(ids)
            sub[NID] = ids
            frames.append(sub)
    return frames


def set_num_threads(num_threads):
    """OpenMP thread count of the native host library (and torch's intra-op pool, which shares the cores)."""
    from .._native import loader
    loader.cpu_lib().fn("dglb200_cpu_set_num_threads", "i")(int(num_threads))
    torch.set_num_threads(int(num_threads))


def get_num_threads():
    from .._native import loader
    fn = loader.cpu_lib().raw("dglb200_cpu_num_threads", __import__("ctypes").c_int32, [])
    return int(fn())


def get_numa_nodes_cores():
    """{numa node: [(core id, [hardware threads]), ...]} from sysfs; {} when unavailable."""
    nodes = {}
    try:
        for node_path in glob.glob("/sys/devices/system/node/node[0-9]*"):
            cores = defaultdict(list)
            for cpu_dir in glob.glob(os.path.join(node_path, "cpu[0-9]*")):
                with open(os.path.join(cpu_dir, "topology", "core_id")) as f:
                    cores[int(f.read().strip())].append(int(os.path.basename(cpu_dir)[3:]))
            nodes[int(os.path.basename(node_path)[4:])] = sorted((c, sorted(t)) for c, t in cores.items())
    except (OSError, ValueError, IndexError):
        dgl_warning("Failed to read NUMA info")
        return {}
    return nodes


def alias_func(func):
    @functools.wraps(func)
    def _fn(*args, **kwargs):
        return func(*args, **kwargs)
    _fn.__doc__ = f"Alias of :func:`dgl.{func.__name__}`."
    return _fn


def apply_each(data, fn, *args, **kwargs):
    """One level: mapping -> dict, sequence -> list, anything else -> ``fn(data)``."""
    if isinstance(data, Mapping):
        return {k: fn(v, *args, **kwargs) for k, v in data.items()}
    if is_listlike(data):
        return [fn(v, *args, **kwargs) for v in data]
    return fn(data, *args, **kwargs)


def recursive_apply_pair(data1, data2, fn, *args, **kwargs):
    """``fn`` on matching leaves of two equally nested containers."""
    if isinstance(data1, Mapping) and isinstance(data2, Mapping):
        return {k: recursive_apply_pair(data1[k], data2[k], fn, *args, **kwargs) for k in data1}
    if isinstance(data1, tuple) and isinstance(data2, tuple):
        return tuple(recursive_apply_pair(a, b, fn, *args, **kwargs) for a, b in zip(data1, data2))
    if is_listlike(data1) and is_listlike(data2):
        return [recursive_apply_pair(a, b, fn, *args, **kwargs) for a, b in zip(data1, data2)]
    return fn(data1, data2, *args, **kwargs)
