"""Small container / id-manipulation helpers that user code reaches through ``dgl.utils`` (reference
python/dgl/utils/internal.py: dict views, list predicates, relabelling, thread-count control). Independent implementations
on plain Python and torch."""
from __future__ import annotations

import functools
import os
from collections.abc import Iterable, Mapping, MutableMapping

import torch

from ..base import DGLError


class InconsistentDtypeException(DGLError):
    """Raised when ids of two different integer widths are mixed."""


def is_listlike(x):
    return isinstance(x, (list, tuple))


def is_iterable(x):
    return isinstance(x, Iterable)


def is_dict_like(x):
    return isinstance(x, Mapping)


def group_as_dict(pairs):
    """[(k, v), ...] -> {k: [v, ...]} keeping first-seen key order."""
    out = {}
    for k, v in pairs:
        out.setdefault(k, []).append(v)
    return out


class ReadOnlyDict(Mapping):
    def __init__(self, data):
        self._data = data

    def __getitem__(self, k):
        return self._data[k]

    def __iter__(self):
        return iter(self._data)

    def __len__(self):
        return len(self._data)


class LazyDict(Mapping):
    """values are produced by ``fn(key)`` on access; ``keys`` fixes the key set."""

    def __init__(self, fn, keys):
        self._fn, self._keys = fn, tuple(keys)

    def __getitem__(self, k):
        if k not in self._keys:
            raise KeyError(k)
        return self._fn(k)

    def __iter__(self):
        return iter(self._keys)

    def __len__(self):
        return len(self._keys)


class HybridDict(Mapping):
    """read-only union of several mappings; the first one holding a key wins."""

    def __init__(self, *dicts):
        self._dicts = dicts

    def __getitem__(self, k):
        for d in self._dicts:
            if k in d:
                return d[k]
        raise KeyError(k)

    def __iter__(self):
        seen = set()
        for d in self._dicts:
            for k in d:
                if k not in seen:
                    seen.add(k)
                    yield k

    def __len__(self):
        return sum(1 for _ in self)


class FlattenedDict(Mapping):
    """``{group: tensor/list}`` viewed as one flat sequence of ``(group, element)`` style keys: key ``i`` maps to
    ``(group name, group[j])`` where groups are laid out one after another."""

    def __init__(self, groups):
        self._groups = groups
        self._names = list(groups)
        sizes = [len(groups[n]) for n in self._names]
        self._offsets = [0]
        for s in sizes:
            self._offsets.append(self._offsets[-1] + s)

    def __len__(self):
        return self._offsets[-1]

    def __iter__(self):
        return iter(range(len(self)))

    def __getitem__(self, i):
        i = int(i)
        if not 0 <= i < len(self):
            raise KeyError(i)
        import bisect
        g = bisect.bisect_right(self._offsets, i) - 1
        return self._names[g], self._groups[self._names[g]][i - self._offsets[g]]


def maybe_flatten_dict(data):
    return FlattenedDict(data) if isinstance(data, Mapping) else data


def apply_each(data, fn, *args, **kwargs):
    """``fn`` on every value of a mapping (keeping keys), or on ``data`` itself."""
    if isinstance(data, Mapping):
        return {k: fn(v, *args, **kwargs) for k, v in data.items()}
    return fn(data, *args, **kwargs)


def recursive_apply_pair(a, b, fn, *args, **kwargs):
    """``fn(x, y)`` over two identically nested structures of lists / tuples / dicts."""
    if isinstance(a, Mapping):
        return {k: recursive_apply_pair(a[k], b[k], fn, *args, **kwargs) for k in a}
    if is_listlike(a):
        return type(a)(recursive_apply_pair(x, y, fn, *args, **kwargs) for x, y in zip(a, b))
    return fn(a, b, *args, **kwargs)


def alias_func(func):
    """A distinct function object forwarding to ``func`` (used to publish one implementation under two names)."""
    @functools.wraps(func)
    def forward(*args, **kwargs):
        return func(*args, **kwargs)
    return forward


def cached_member(cache, prefix):
    """Memoise a method's result in ``getattr(self, cache)[prefix + args]``."""
    def deco(method):
        @functools.wraps(method)
        def wrapper(self, *args, **kwargs):
            store = getattr(self, cache)
            key = (prefix,) + tuple(args) + tuple(sorted(kwargs.items()))
            if key not in store:
                store[key] = method(self, *args, **kwargs)
            return store[key]
        return wrapper
    return deco


def retry_method_with_fix(fix):
    """Call the method; on an exception run ``fix(self)`` once and call it again."""
    def deco(method):
        @functools.wraps(method)
        def wrapper(self, *args, **kwargs):
            try:
                return method(self, *args, **kwargs)
            except Exception:  # noqa: BLE001 - by contract: any failure triggers the repair
                fix(self)
                return method(self, *args, **kwargs)
        return wrapper
    return deco


# ---------------------------------------------------------------------------------------------- id helpers
def relabel(x):
    """(distinct old ids in ascending order, map old -> new as a tensor of size max+1, -1 where unused)."""
    x = torch.as_tensor(x)
    uniq = torch.unique(x)
    size = int(uniq.max()) + 1 if uniq.numel() else 0
    old_to_new = torch.full((size,), -1, dtype=torch.int64, device=x.device)
    old_to_new[uniq.long()] = torch.arange(uniq.numel(), device=x.device)
    return uniq, old_to_new


def build_relabel_map(x, is_sorted=False):
    """Tensor map old id -> consecutive new id (ascending order of the old ids)."""
    x = torch.as_tensor(x)
    uniq = x if is_sorted else torch.unique(x)
    size = int(uniq.max()) + 1 if uniq.numel() else 0
    m = torch.zeros(size, dtype=torch.int64, device=x.device)
    m[uniq.long()] = torch.arange(uniq.numel(), device=x.device)
    return m


def build_relabel_dict(x):
    """{old id: new id} with new ids in order of first appearance."""
    out = {}
    for v in torch.as_tensor(x).tolist():
        out.setdefault(v, len(out))
    return out


def make_invmap(array, use_numpy=True):
    """(sorted distinct values, {value: rank}, the array rewritten as ranks)."""
    t = torch.as_tensor(array)
    uniq, inv = torch.unique(t, return_inverse=True)
    values = uniq.tolist()
    remapped = inv.numpy() if use_numpy else inv
    return (uniq.numpy() if use_numpy else uniq), {v: i for i, v in enumerate(values)}, remapped


def set_diff(ar1, ar2):
    """Sorted elements of ``ar1`` that are not in ``ar2``."""
    a, b = torch.unique(torch.as_tensor(ar1)), torch.as_tensor(ar2)
    return a[~torch.isin(a, b)]


def compensate(ids, n_total):
    """Sorted ids of ``range(n_total)`` missing from ``ids``."""
    mask = torch.ones(n_total, dtype=torch.bool)
    mask[torch.as_tensor(ids).long()] = False
    return mask.nonzero().squeeze(1)


def reorder_index(idx, order):
    return torch.as_tensor(idx)[torch.as_tensor(order).long()]


def zero_index(size):
    return torch.zeros(size, dtype=torch.int64)


def to_nbits_int(tensor, nbits):
    if nbits not in (32, 64):
        raise DGLError(f"nbits must be 32 or 64, got {nbits}")
    return tensor.to(torch.int32 if nbits == 32 else torch.int64)


# ---------------------------------------------------------------------------------------------- threads / topology
def set_num_threads(num_threads):
    """Threads used by the host kernels (OpenMP team of libdglb200_cpu) and by torch."""
    os.environ["OMP_NUM_THREADS"] = str(int(num_threads))
    torch.set_num_threads(int(num_threads))


def get_num_threads():
    return torch.get_num_threads()


def get_numa_nodes_cores():
    """{numa node id: [cpu ids]} from sysfs (empty dict where the topology is not exposed)."""
    base = "/sys/devices/system/node"
    out = {}
    try:
        for entry in sorted(os.listdir(base)):
            if entry.startswith("node") and entry[4:].isdigit():
                cpus = []
                for part in open(os.path.join(base, entry, "cpulist")).read().strip().split(","):
                    if "-" in part:
                        a, b = part.split("-")
                        cpus.extend(range(int(a), int(b) + 1))
                    elif part:
                        cpus.append(int(part))
                out[int(entry[4:])] = cpus
    except OSError:
        return {}
    return out
