"""ItemSet / HeteroItemSet and item samplers (parity: /root/reference/python/dgl/graphbolt/itemset.py,
item_sampler.py: ItemSampler :283, DistributedItemSampler :520)."""
from __future__ import annotations

import math
from collections.abc import Mapping

import torch
import torch.distributed as dist

from .minibatch import MiniBatch


class ItemSet:
    """An indexable collection of items: an int (range), a tensor, or a tuple of equally long tensors."""

    def __init__(self, items, names=None):
        if isinstance(items, int):
            self._items = torch.arange(items)
            self._single = True
        elif torch.is_tensor(items):
            self._items = items
            self._single = True
        else:
            self._items = tuple(items)
            self._single = False
            n = len(self._items[0])
            assert all(len(x) == n for x in self._items), "all item tensors must have the same length"
        if names is not None:
            names = (names,) if isinstance(names, str) else tuple(names)
            assert len(names) == (1 if self._single else len(self._items)), "Number of items and names must match."
        self._names = names

    def __len__(self):
        return len(self._items) if self._single else len(self._items[0])

    def __getitem__(self, idx):
        if self._single:
            return self._items[idx]
        return tuple(x[idx] for x in self._items)

    @property
    def names(self):
        return self._names

    @property
    def num_items(self):
        return 1 if self._single else len(self._items)

    def __repr__(self):
        return f"ItemSet(items={self._items}, names={self._names})"


class HeteroItemSet:
    """Dict of ItemSets keyed by node type / canonical edge type string."""

    def __init__(self, itemsets):
        self._itemsets = dict(itemsets)
        self._keys = list(self._itemsets.keys())
        self._offsets = [0]
        for k in self._keys:
            self._offsets.append(self._offsets[-1] + len(self._itemsets[k]))
        self._names = self._itemsets[self._keys[0]].names if self._keys else None

    def __len__(self):
        return self._offsets[-1]

    def __getitem__(self, idx):
        if isinstance(idx, int):
            for i, k in enumerate(self._keys):
                if idx < self._offsets[i + 1]:
                    return {k: self._itemsets[k][idx - self._offsets[i]]}
            raise IndexError(idx)
        idx = torch.as_tensor(idx)
        out = {}
        for i, k in enumerate(self._keys):
            m = (idx >= self._offsets[i]) & (idx < self._offsets[i + 1])
            if bool(m.any()):
                out[k] = self._itemsets[k][idx[m] - self._offsets[i]]
        return out

    @property
    def names(self):
        return self._names

    def __repr__(self):
        return f"HeteroItemSet(itemsets={self._itemsets}, names={self._names})"


ItemSetDict = HeteroItemSet


def minibatcher_default(batch, names):
    if names is None:
        return batch
    mb = MiniBatch()
    if isinstance(batch, Mapping):
        items = {k: (v if isinstance(v, tuple) else (v,)) for k, v in batch.items()}
        for i, name in enumerate(names):
            setattr(mb, name, {k: v[i] for k, v in items.items()})
    else:
        items = batch if isinstance(batch, tuple) else (batch,)
        for name, item in zip(names, items):
            setattr(mb, name, item)
    return mb


class ItemSampler:
    """Iterate minibatches of items; supports shuffle / drop_last and a custom minibatcher."""

    def __init__(self, item_set, batch_size, minibatcher=minibatcher_default, drop_last=False, shuffle=False,
                 use_indexing=True, buffer_size=-1, seed=None):
        self._item_set = item_set
        self._names = item_set.names
        self._batch_size = batch_size
        self._minibatcher = minibatcher
        self._drop_last = drop_last
        self._shuffle = shuffle
        self._epoch = 0
        self._seed = seed if seed is not None else 0
        self._rank, self._world = 0, 1
        self._drop_uneven_inputs = False
        self._worker, self._num_workers = 0, 1      # set inside loader worker processes: batch b goes to worker b % n

    def _positions(self):
        n = len(self._item_set)
        if self._shuffle:
            g = torch.Generator()
            g.manual_seed(self._seed + self._epoch)
            return torch.randperm(n, generator=g)
        return torch.arange(n)

    def __iter__(self):
        pos = self._positions()
        if self._world > 1:
            n = len(pos)
            per = math.ceil(n / self._world)
            lo, hi = self._rank * per, min(n, (self._rank + 1) * per)
            pos = pos[lo:hi]
            if self._drop_uneven_inputs:
                nb = (n // self._world) // self._batch_size if self._drop_last else math.ceil((n // self._world) / self._batch_size)
                pos = pos[: nb * self._batch_size]
        n = len(pos)
        nb = n // self._batch_size if self._drop_last else math.ceil(n / self._batch_size)
        for b in range(self._worker, nb, self._num_workers):
            idx = pos[b * self._batch_size: (b + 1) * self._batch_size]
            yield self._minibatcher(self._item_set[idx], self._names)
        self._epoch += 1

    def shard_for_worker(self, worker, num_workers, epoch=None):
        """Restrict iteration to the batches ``worker, worker + num_workers, ...`` (the permutation of an epoch is the
        same in every process because it is seeded by ``seed + epoch``)."""
        self._worker, self._num_workers = int(worker), max(1, int(num_workers))
        if epoch is not None:
            self._epoch = int(epoch)

    def __len__(self):
        n = len(self._item_set)
        if self._world > 1:
            n = math.ceil(n / self._world)
        return n // self._batch_size if self._drop_last else math.ceil(n / self._batch_size)

    # datapipe-style functional verbs -----------------------------------------------------------
    def sample_neighbor(self, graph, fanouts, **kw):
        from .impl.neighbor_sampler import NeighborSampler
        return NeighborSampler(self, graph, fanouts, **kw)

    def sample_layer_neighbor(self, graph, fanouts, **kw):
        from .impl.neighbor_sampler import LayerNeighborSampler
        return LayerNeighborSampler(self, graph, fanouts, **kw)

    def sample_uniform_negative(self, graph, negative_ratio):
        from .impl.negative_sampler import UniformNegativeSampler
        return UniformNegativeSampler(self, graph, negative_ratio)

    def copy_to(self, device, non_blocking=False):
        from .datapipes import CopyTo
        return CopyTo(self, device, non_blocking)

    def transform(self, fn):
        from .datapipes import MiniBatchTransformer
        return MiniBatchTransformer(self, fn)


class DistributedItemSampler(ItemSampler):
    """Shard the item set across torch.distributed ranks (reference item_sampler.py:520)."""

    def __init__(self, item_set, batch_size, minibatcher=minibatcher_default, drop_last=False, shuffle=False,
                 drop_uneven_inputs=False, buffer_size=-1, seed=None):
        super().__init__(item_set, batch_size, minibatcher, drop_last, shuffle, seed=seed)
        self._drop_uneven_inputs = drop_uneven_inputs
        if dist.is_initialized():
            self._rank, self._world = dist.get_rank(), dist.get_world_size()


def __getattr__(name):
    # names the reference keeps in this module but that live elsewhere in this package
    if name.startswith("__"):
        raise AttributeError(name)
    import importlib
    for modname in ('dgl_b200.graphbolt.compat',):
        mod = importlib.import_module(modname)
        if name in vars(mod):
            return vars(mod)[name]
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
