"""Index datasets behind the loader and the legacy collator API (parity:
/root/reference/python/dgl/dataloading/dataloader.py: TensorizedDataset :184, DDPTensorizedDataset :255,
create_tensorized_dataset :166, CollateWrapper :476, WorkerInitWrapper :777, remove/restore_parent_storage_columns
:558-620, and the pre-0.8 NodeCollator / EdgeCollator).

``DataLoader`` itself uses ``_IndexBatcher`` (batches cut on the target device); these classes expose the same
behaviour as stand-alone iterable datasets so user code that builds its own ``torch.utils.data.DataLoader`` around
them keeps working."""
from __future__ import annotations

import math
from collections.abc import Mapping

import torch
import torch.distributed as dist

from ..base import DGLError
from ..frame import LazyFeature

__all__ = ["TensorizedDataset", "DDPTensorizedDataset", "create_tensorized_dataset", "CollateWrapper",
           "WorkerInitWrapper", "remove_parent_storage_columns", "restore_parent_storage_columns", "NodeCollator",
           "EdgeCollator", "find_exclude_eids"]


def _flatten(indices):
    """(id tensor, type keys, type offsets): dict indices are concatenated in key order."""
    if isinstance(indices, Mapping):
        keys = list(indices.keys())
        parts = [torch.as_tensor(indices[k]) for k in keys]
        offsets = [0]
        for p in parts:
            offsets.append(offsets[-1] + p.shape[0])
        return torch.cat(parts), keys, offsets
    ids = torch.as_tensor(indices)
    if ids.dtype == torch.bool:
        ids = ids.nonzero().view(-1)
    return ids, None, None


class _BatchIter:
    def __init__(self, ids, positions, batch_size, drop_last, keys, offsets):
        self.ids, self.pos, self.bs, self.drop, self.keys, self.offsets = ids, positions, batch_size, drop_last, keys, offsets
        self.at = 0

    def __iter__(self):
        return self

    def __next__(self):
        n = self.pos.shape[0]
        lo, hi = self.at, min(self.at + self.bs, n)
        if lo >= n or (self.drop and hi - lo < self.bs):
            raise StopIteration
        self.at = hi
        p = self.pos[lo:hi]
        if self.keys is None:
            return self.ids[p]
        out = {}
        for i, k in enumerate(self.keys):
            m = (p >= self.offsets[i]) & (p < self.offsets[i + 1])
            if bool(m.any()):
                out[k] = self.ids[p[m]]
        return out


class TensorizedDataset(torch.utils.data.IterableDataset):
    """Mini-batches of an index tensor (or dict of them): ``shuffle()`` draws a new permutation, iteration slices it.
    With ``use_shared_memory`` the permutation lives in torch shared memory so forked workers see each reshuffle."""

    def __init__(self, indices, batch_size, drop_last, shuffle, use_shared_memory=False):
        self._ids, self._keys, self._offsets = _flatten(indices)
        self._perm = torch.arange(self._ids.shape[0])
        if use_shared_memory:
            self._perm.share_memory_()
        self.batch_size, self.drop_last, self._shuffle = batch_size, drop_last, shuffle

    def shuffle(self):
        """Called by the loader at the start of every epoch."""
        self._perm.copy_(torch.randperm(self._ids.shape[0]))

    def __iter__(self):
        return _BatchIter(self._ids, self._perm.clone(), self.batch_size, self.drop_last, self._keys, self._offsets)

    def __len__(self):
        n = self._ids.shape[0]
        return n // self.batch_size if self.drop_last else math.ceil(n / self.batch_size)


class DDPTensorizedDataset(torch.utils.data.IterableDataset):
    """Per-rank shard of the indices: every epoch all ranks draw the *same* seeded permutation (``ddp_seed + epoch``)
    and take a contiguous 1/world slice, padding by wrap-around unless ``drop_last`` — so shards are disjoint, equally
    sized and differ between epochs."""

    def __init__(self, indices, batch_size, drop_last, ddp_seed, shuffle):
        if not dist.is_available() or not dist.is_initialized():
            raise DGLError("DDPTensorizedDataset requires torch.distributed to be initialised")
        self._ids, self._keys, self._offsets = _flatten(indices)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.seed, self.epoch = ddp_seed, 0
        self.batch_size, self.drop_last, self._shuffle = batch_size, drop_last, shuffle
        n = self._ids.shape[0]
        self.num_samples = n // self.world if drop_last else math.ceil(n / self.world)
        self.total_size = self.num_samples * self.world
        self._local = None
        self.shuffle()

    def set_epoch(self, epoch):
        self.epoch = epoch

    def shuffle(self):
        n = self._ids.shape[0]
        if self._shuffle:
            gen = torch.Generator().manual_seed(self.seed + self.epoch)
            perm = torch.randperm(n, generator=gen)
        else:
            perm = torch.arange(n)
        if self.total_size > n:
            reps = math.ceil(self.total_size / max(n, 1))
            perm = perm.repeat(reps)[: self.total_size]
        else:
            perm = perm[: self.total_size]
        self._local = perm[self.rank * self.num_samples: (self.rank + 1) * self.num_samples]

    def __iter__(self):
        return _BatchIter(self._ids, self._local, self.batch_size, self.drop_last, self._keys, self._offsets)

    def __len__(self):
        return self.num_samples // self.batch_size if self.drop_last else math.ceil(self.num_samples / self.batch_size)


def create_tensorized_dataset(indices, batch_size, drop_last, use_ddp, ddp_seed, shuffle, use_shared_memory):
    if use_ddp:
        return DDPTensorizedDataset(indices, batch_size, drop_last, ddp_seed, shuffle)
    return TensorizedDataset(indices, batch_size, drop_last, shuffle, use_shared_memory)


class CollateWrapper:
    """``collate_fn`` for a torch DataLoader over index batches: samples, then (optionally) moves to ``device``."""

    def __init__(self, sample_func, g, use_uva, device):
        self.sample_func, self.g, self.use_uva, self.device = sample_func, g, use_uva, device

    def __call__(self, items):
        graph_device = getattr(self.g, "device", None)
        if self.use_uva or (graph_device is not None and graph_device != torch.device("cpu")):
            move = lambda x: x.to(self.device)  # noqa: E731
            items = {k: move(v) for k, v in items.items()} if isinstance(items, Mapping) else move(items)
        return self.sample_func(self.g, items)


class WorkerInitWrapper:
    """``worker_init_fn`` that caps each worker's intra-op threads to its share of the cores before the user hook."""

    def __init__(self, func):
        self.func = func

    def __call__(self, worker_id):
        info = torch.utils.data.get_worker_info()
        if info is not None and info.num_workers > 0:
            import os
            torch.set_num_threads(max(1, (os.cpu_count() or 1) // info.num_workers))
        if self.func is not None:
            self.func(worker_id)


def _frames(item):
    """Every Frame reachable from a sampler output (graphs / blocks nested in lists, tuples, dicts)."""
    from ..heterograph import DGLGraph
    if isinstance(item, DGLGraph):
        yield from item._node_frames
        yield from item._edge_frames
    elif isinstance(item, Mapping):
        for v in item.values():
            yield from _frames(v)
    elif isinstance(item, (list, tuple)):
        for v in item:
            yield from _frames(v)


def remove_parent_storage_columns(item, g):
    """Before a sampled subgraph crosses a process boundary: columns that are lazy views into the *parent graph's*
    storage are replaced by name-only :class:`LazyFeature` markers so the parent tensors are not pickled."""
    parent = {}
    for fr in list(g._node_frames) + list(g._edge_frames):
        for k in fr.keys():
            parent[id(fr._columns[k].storage)] = k
    for fr in _frames(item):
        for k in list(fr.keys()):
            col = fr._columns[k]
            if col.index is not None and id(col.storage) in parent:
                fr._columns[k] = type(col)(LazyFeature(parent[id(col.storage)], col.index))
    return item


def restore_parent_storage_columns(item, g):
    """Inverse of :func:`remove_parent_storage_columns` in the consumer process: markers become lazy views again."""
    from ..heterograph import DGLGraph

    def fix(graph):
        pools = [(graph._node_frames, g._node_frames if len(g._node_frames) == len(graph._node_frames) else None),
                 (graph._edge_frames, g._edge_frames if len(g._edge_frames) == len(graph._edge_frames) else None)]
        for frames, parents in pools:
            for i, fr in enumerate(frames):
                for k in list(fr.keys()):
                    col = fr._columns[k]
                    if isinstance(col.storage, LazyFeature) and col.storage.id_ is not None:
                        # blocks stack src + dst node frames: fall back to a by-name search
                        cands = [parents[i]] if parents is not None else list(g._node_frames) + list(g._edge_frames)
                        for pf in cands:
                            if col.storage.name in pf:
                                fr._columns[k] = pf._columns[col.storage.name].subcolumn(col.storage.id_)
                                break

    def walk(x):
        if isinstance(x, DGLGraph):
            fix(x)
        elif isinstance(x, Mapping):
            for v in x.values():
                walk(v)
        elif isinstance(x, (list, tuple)):
            for v in x:
                walk(v)

    walk(item)
    return item


def find_exclude_eids(g, exclude_mode, eids, **kwargs):
    """Edges to leave out of neighbour sampling for a batch of seed edges (base.py:find_exclude_eids).

    ``exclude_mode``: None | 'self' | 'reverse_id' (needs ``reverse_eid_map``) | 'reverse_types' (needs
    ``reverse_etype_map``) | callable. ``output_device`` moves the result."""
    from .base import _find_exclude_eids
    out = _find_exclude_eids(g, exclude_mode, eids, kwargs.get("reverse_eid_map"), kwargs.get("reverse_etype_map"))
    dev = kwargs.get("output_device")
    if out is not None and dev is not None:
        out = {k: v.to(dev) for k, v in out.items()} if isinstance(out, Mapping) else out.to(dev)
    return out


class NodeCollator:
    """Pre-0.8 API: wraps (graph, seed node ids, sampler); ``collate(items)`` returns
    ``(input_nodes, output_nodes, blocks)`` and ``dataset`` is what a torch DataLoader iterates."""

    def __init__(self, g, nids, graph_sampler):
        self.g, self.nids, self.graph_sampler = g, nids, graph_sampler
        if isinstance(nids, Mapping):
            self._dataset = [(k, int(i)) for k, v in nids.items() for i in torch.as_tensor(v).tolist()]
        else:
            self._dataset = torch.as_tensor(nids)

    @property
    def dataset(self):
        return self._dataset

    def _gather(self, items):
        if isinstance(self.nids, Mapping):
            by = {}
            for k, i in items:
                by.setdefault(k, []).append(i)
            return {k: torch.tensor(v, dtype=self.g.idtype) for k, v in by.items()}
        return torch.as_tensor(items).to(self.g.idtype) if not torch.is_tensor(items) else items.to(self.g.idtype)

    def collate(self, items):
        return self.graph_sampler.sample(self.g, self._gather(items))


class EdgeCollator(NodeCollator):
    """Pre-0.8 edge-wise counterpart: builds the edge-prediction sampler from the keyword options and returns
    ``(input_nodes, pair_graph, [neg_graph,] blocks)``."""

    def __init__(self, g, eids, graph_sampler, g_sampling=None, exclude=None, reverse_eids=None, reverse_etypes=None,
                 negative_sampler=None):
        from .base import as_edge_prediction_sampler
        super().__init__(g, eids, as_edge_prediction_sampler(graph_sampler, exclude=exclude, reverse_eids=reverse_eids,
                                                             reverse_etypes=reverse_etypes,
                                                             negative_sampler=negative_sampler))
        self.g_sampling = g_sampling if g_sampling is not None else g

    def collate(self, items):
        return self.graph_sampler.sample(self.g_sampling, self._gather(items))
