"""FFI container types of the reference (parity: /root/reference/python/dgl/container.py: List, Map, StrMap, Value).

The reference needs them to pass Python lists / dicts through its PackedFunc ABI. Native calls here take raw
pointers through ctypes, so these are thin Python-side equivalents kept for API compatibility."""
from __future__ import annotations

from collections.abc import Mapping, Sequence

__all__ = ["List", "Map", "StrMap", "Value", "convert_to_strmap"]


class List(Sequence):
    """Immutable sequence."""

    def __init__(self, items=()):
        self._items = tuple(items)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._items[j] for j in range(*i.indices(len(self._items)))]
        if i < -len(self._items) or i >= len(self._items):
            raise IndexError(f"List index out of range. List size: {len(self._items)}, got index {i}")
        return self._items[i]

    def __len__(self):
        return len(self._items)

    def __repr__(self):
        return f"List({list(self._items)})"


class Map(Mapping):
    """Immutable mapping."""

    def __init__(self, items=None):
        self._d = dict(items or {})

    def __getitem__(self, k):
        return self._d[k]

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def items(self):
        return list(self._d.items())

    def __repr__(self):
        return f"Map({self._d})"


class StrMap(Map):
    """Mapping whose keys must be strings."""

    def __init__(self, items=None):
        super().__init__(items)
        for k in self._d:
            if not isinstance(k, str):
                raise TypeError(f"StrMap keys must be str, got {type(k).__name__}")


class Value:
    """Boxed scalar."""

    def __init__(self, data):
        self.data = data

    def __repr__(self):
        return f"Value({self.data!r})"


def convert_to_strmap(value):
    """dict -> StrMap."""
    return StrMap(value)
