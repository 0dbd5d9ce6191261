"""Membership filter over an id set (parity: /root/reference/python/dgl/utils/filter.py, src/array/filter.cc,
src/array/cuda/array_filter.cu).

The reference builds a GPU hash table; here the set is kept sorted and queries are one vectorised binary search
(``torch.searchsorted``), which is a single kernel on either device and needs no table sizing."""
from __future__ import annotations

import torch

__all__ = ["Filter"]


class Filter:
    def __init__(self, ids):
        ids = torch.as_tensor(ids)
        self._sorted = torch.sort(ids)[0]

    def _hit(self, test):
        test = torch.as_tensor(test, device=self._sorted.device)
        if self._sorted.numel() == 0:
            return torch.zeros(test.shape[0], dtype=torch.bool, device=test.device)
        pos = torch.searchsorted(self._sorted, test.to(self._sorted.dtype)).clamp(max=self._sorted.numel() - 1)
        return self._sorted[pos] == test

    def find_included_indices(self, test):
        """Positions of ``test`` whose id is in the filter."""
        return torch.nonzero(self._hit(test)).view(-1)

    def find_excluded_indices(self, test):
        """Positions of ``test`` whose id is not in the filter."""
        return torch.nonzero(~self._hit(test)).view(-1)
