"""DataLoader / GraphDataLoader.

Parity: /root/reference/python/dgl/dataloading/dataloader.py (DataLoader :788, DDPTensorizedDataset :255,
_PrefetchingIter :623, _prefetch :487, GraphDataLoader :1376). Differences by design: the loader is a
plain iterator (no torch DataLoader subclass): index batching happens on the target device, DDP sharding
is a seeded permutation slice per rank, and feature prefetch runs on a side CUDA stream with an optional
producer thread. Sampling itself runs on the GPU when the graph is on the GPU or pinned (UVA).
"""
from __future__ import annotations

import math
import os
import queue
import threading
from collections.abc import Mapping

import torch
import torch.distributed as dist

from ..base import DGLError, dgl_warning
from ..batch import batch as batch_graphs
from .base import assign_lazy_features

PREFETCHER_TIMEOUT = int(os.environ.get("DGL_PREFETCHER_TIMEOUT", "30"))


def _numel(indices):
    if isinstance(indices, Mapping):
        return sum(v.shape[0] for v in indices.values())
    return indices.shape[0]


class _IndexBatcher:
    """Yields index batches (tensor or {type: tensor}); shuffles per epoch; shards across DDP ranks."""

    def __init__(self, indices, batch_size, shuffle, drop_last, use_ddp, ddp_seed, device):
        self.is_dict = isinstance(indices, Mapping)
        if self.is_dict:
            self.types = list(indices.keys())
            sizes = [indices[t].shape[0] for t in self.types]
            self.type_offsets = [0]
            for s in sizes:
                self.type_offsets.append(self.type_offsets[-1] + s)
            self.ids = torch.cat([torch.as_tensor(indices[t]) for t in self.types])
        else:
            self.ids = torch.as_tensor(indices)
            if self.ids.dtype == torch.bool:
                self.ids = self.ids.nonzero().squeeze(1)
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.use_ddp = use_ddp
        self.seed = ddp_seed
        self.epoch = 0
        self.device = device
        if use_ddp:
            if not dist.is_initialized():
                raise DGLError("use_ddp=True requires torch.distributed to be initialised")
            self.rank = dist.get_rank()
            self.world = dist.get_world_size()
        else:
            self.rank, self.world = 0, 1
        n = self.ids.shape[0]
        if use_ddp:
            if drop_last:
                self.num_local = n // self.world
            else:
                self.num_local = math.ceil(n / self.world)
        else:
            self.num_local = n

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        if self.drop_last:
            return self.num_local // self.batch_size
        return math.ceil(self.num_local / self.batch_size)

    def _local_positions(self):
        n = self.ids.shape[0]
        if self.use_ddp:
            if self.shuffle:
                g = torch.Generator()
                g.manual_seed(self.seed + self.epoch)
                perm = torch.randperm(n, generator=g)
            else:
                perm = torch.arange(n)
            total = self.num_local * self.world
            if total > n:
                perm = torch.cat([perm, perm[: total - n]])
            else:
                perm = perm[:total]
            return perm[self.rank * self.num_local: (self.rank + 1) * self.num_local]
        if self.shuffle:
            return torch.randperm(n)
        return torch.arange(n)

    def __iter__(self):
        pos = self._local_positions()
        nb = len(self)
        for b in range(nb):
            p = pos[b * self.batch_size: (b + 1) * self.batch_size]
            if self.is_dict:
                out = {}
                for ti, t in enumerate(self.types):
                    lo, hi = self.type_offsets[ti], self.type_offsets[ti + 1]
                    m = (p >= lo) & (p < hi)
                    if bool(m.any()):
                        out[t] = self.ids[p[m]].to(self.device)
                yield out
            else:
                yield self.ids[p.to(self.ids.device)].to(self.device)
        self.epoch += 1 if not self.use_ddp else 0


class DataLoader:
    """Sampled-graph loader: iterates index batches and calls ``graph_sampler.sample(graph, batch)``."""

    def __init__(self, graph, indices, graph_sampler, device=None, use_ddp=False, ddp_seed=0, batch_size=1,
                 drop_last=False, shuffle=False, use_prefetch_thread=None, use_alternate_streams=None,
                 pin_prefetcher=None, use_uva=False, gpu_cache=None, num_workers=0, **kwargs):
        self.graph = graph
        self.graph_sampler = graph_sampler
        if device is None:
            device = graph.device if not use_uva else torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if use_uva:
            if graph.device.type != "cpu":
                raise DGLError("Graph must be on CPU if UVA sampling is enabled.")
            if num_workers > 0:
                raise DGLError("num_workers must be 0 if UVA sampling is enabled.")
            if not graph.is_pinned():
                graph.create_formats_()
                graph.pin_memory_()
            indices_device = self.device
        else:
            if graph.device.type == "cuda" and num_workers > 0:
                raise DGLError("num_workers must be 0 if graph and indices are on CUDA.")
            indices_device = graph.device
        # num_workers > 0 on a host graph: real sampling processes (forked, graph shared copy-on-write, batches returned
        # through torch.multiprocessing's shared-memory tensors) - see _WorkerPool. DGLB200_LOADER_THREADS=1 keeps the
        # lighter single prefetch thread instead.
        self._use_processes = num_workers > 0 and graph.device.type == "cpu" and not use_uva and \
            os.environ.get("DGLB200_LOADER_THREADS", "0") != "1"
        if num_workers > 0 and not self._use_processes:
            use_prefetch_thread = True
        self._pool = None
        self.persistent_workers = bool(kwargs.pop("persistent_workers", False))
        self.prefetch_factor = int(kwargs.pop("prefetch_factor", 2) or 2)
        self.use_uva = use_uva
        self.batcher = _IndexBatcher(indices, batch_size, shuffle, drop_last, use_ddp, ddp_seed, indices_device)
        if use_alternate_streams is None:
            use_alternate_streams = self.device.type == "cuda" and graph.device.type == "cpu" and not use_uva
        self.use_alternate_streams = use_alternate_streams and torch.cuda.is_available()
        self.use_prefetch_thread = bool(use_prefetch_thread)
        self.pin_prefetcher = pin_prefetcher
        self.gpu_cache = gpu_cache or {}
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.num_workers = num_workers
        self.kwargs = kwargs

    def attach_data(self, name, data):
        """Add a data source other than node / edge features for prefetching (reference dataloader.py:1258): samplers can
        request rows of it through ``LazyFeature`` and the loader resolves them alongside the graph features."""
        if not hasattr(self, "other_storages"):
            self.other_storages = {}
        from ..storages import wrap_storage
        self.other_storages[name] = wrap_storage(data)

    def set_epoch(self, epoch):
        self.batcher.set_epoch(epoch)

    def __len__(self):
        return len(self.batcher)

    def _produce(self, idx):
        batch = self.graph_sampler.sample(self.graph, idx)
        batch = assign_lazy_features(batch, self.graph, self.device)
        batch = _to_device(batch, self.device)
        return batch

    def _iter_processes(self, it):
        pool = self._pool
        if pool is None or not pool.alive():
            pool = self._pool = _WorkerPool(self.graph, self.graph_sampler, self.num_workers, self.prefetch_factor)
        try:
            for batch in pool.run(it):
                yield _to_device(batch, self.device)
        finally:
            if not self.persistent_workers:
                pool.shutdown()
                self._pool = None

    def __del__(self):
        pool = getattr(self, "_pool", None)
        if pool is not None:
            pool.shutdown()

    def __iter__(self):
        it = iter(self.batcher)
        if self._use_processes:
            yield from self._iter_processes(it)
            return
        if not self.use_prefetch_thread:
            stream = torch.cuda.Stream(self.device) if self.use_alternate_streams else None
            for idx in it:
                if stream is not None:
                    cur = torch.cuda.current_stream(self.device)
                    with torch.cuda.stream(stream):
                        batch = self._produce(idx)
                    cur.wait_stream(stream)
                    _record_stream(batch, cur)
                    yield batch
                else:
                    yield self._produce(idx)
            return
        q = queue.Queue(maxsize=2)
        done = object()
        stop = threading.Event()          # set when the consumer abandons the iterator: the producer must not block forever

        def offer(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.2)
                    return True
                except queue.Full:
                    continue
            return False

        def worker():
            try:
                if self.device.type == "cuda":
                    torch.cuda.set_device(self.device)
                stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
                for idx in it:
                    if stop.is_set():
                        return
                    if stream is not None:
                        with torch.cuda.stream(stream):
                            batch = self._produce(idx)
                        ev = torch.cuda.Event()
                        ev.record(stream)
                        if not offer((batch, ev, None)):
                            return
                    elif not offer((self._produce(idx), None, None)):
                        return
                offer((done, None, None))
            except Exception as e:  # noqa: BLE001
                offer((None, None, e))

        th = threading.Thread(target=worker, daemon=True)
        th.start()
        try:
            while True:
                try:
                    batch, ev, exc = q.get(timeout=PREFETCHER_TIMEOUT * 20)
                except queue.Empty:
                    raise RuntimeError(f"Prefetcher thread timed out at {PREFETCHER_TIMEOUT * 20} seconds.")
                if exc is not None:
                    raise exc
                if batch is done:
                    break
                if ev is not None:
                    torch.cuda.current_stream(self.device).wait_event(ev)
                    _record_stream(batch, torch.cuda.current_stream(self.device))
                yield batch
        finally:                          # normal end, exception, or the generator being closed early (break / GC)
            stop.set()
            while True:                   # drop whatever is queued so GPU batches are released
                try:
                    q.get_nowait()
                except queue.Empty:
                    break
            th.join(timeout=5)

    # CPU affinity helpers kept for API parity (reference dataloader.py:1159-1257)
    def enable_cpu_affinity(self, loader_cores=None, compute_cores=None, verbose=True):
        from contextlib import nullcontext
        return nullcontext()


def _worker_loop(graph, sampler, index_q, result_q, wid, base_seed):
    """Body of one sampling process: receives (task id, index batch), returns (task id, batch | exception)."""
    import random as _random
    if os.environ.get("DGLB200_DEBUG_WORKER"):            # dump every thread's stack if the worker stalls
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["DGLB200_DEBUG_WORKER"]), exit=True)
    torch.set_num_threads(1)
    # GNU OpenMP does not survive fork(): the child inherits the parent's thread-pool bookkeeping without the threads,
    # and the first parallel region with more than one thread waits for them forever. One-thread teams run inline and
    # never touch the pool, so the native library is pinned to one thread before the first native call of this process.
    try:
        from .._native import loader as _loader
        _loader.cpu_lib().fn("dglb200_cpu_set_num_threads", "i")(1)
    except Exception:  # noqa: BLE001 - a build without the setter has no OpenMP teams to worry about
        pass
    torch.manual_seed(base_seed + wid)
    _random.seed(base_seed + wid)
    try:
        from .. import random as dgl_random
        dgl_random.seed(base_seed + wid)
    except Exception:  # noqa: BLE001
        pass
    while True:
        task = index_q.get()
        if task is None:
            return
        tid, idx = task
        try:
            batch = sampler.sample(graph, idx)
            batch = assign_lazy_features(batch, graph, torch.device("cpu"))
            result_q.put((tid, batch, None))
        except Exception as e:  # noqa: BLE001
            import traceback
            result_q.put((tid, None, f"{type(e).__name__}: {e}\n{traceback.format_exc()}"))


class _WorkerPool:
    """``num_workers`` forked sampling processes. The (host) graph is inherited copy-on-write, index batches go out
    round-robin with ``prefetch_factor`` tasks in flight per worker, results come back through one queue (tensors travel
    in shared memory) and are re-ordered, so iteration order equals the single-process order.
    Reference: the torch DataLoader worker pool behind python/dgl/dataloading/dataloader.py:925-1090."""

    def __init__(self, graph, sampler, num_workers, prefetch_factor=2):
        import torch.multiprocessing as mp
        ctx = mp.get_context("fork")
        self.n, self.depth = num_workers, max(1, prefetch_factor)
        self.index_qs = [ctx.Queue() for _ in range(num_workers)]
        self.result_q = ctx.Queue()
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) & 0x7FFFFFFF
        self.procs = [ctx.Process(target=_worker_loop, args=(graph, sampler, self.index_qs[w], self.result_q, w, seed),
                                  daemon=True) for w in range(num_workers)]
        for p in self.procs:
            p.start()

    def alive(self):
        return all(p.is_alive() for p in self.procs)

    def run(self, index_iter):
        sent = done = 0
        pending = {}
        exhausted = False

        def dispatch():
            nonlocal sent, exhausted
            while not exhausted and sent - done < self.n * self.depth:
                try:
                    idx = next(index_iter)
                except StopIteration:
                    exhausted = True
                    return
                self.index_qs[sent % self.n].put((sent, idx.cpu() if torch.is_tensor(idx) else idx))
                sent += 1

        dispatch()
        while done < sent:
            while done not in pending:
                try:
                    tid, batch, err = self.result_q.get(timeout=PREFETCHER_TIMEOUT * 20)
                except queue.Empty:
                    if not self.alive():
                        raise RuntimeError("a DataLoader worker process died unexpectedly") from None
                    raise RuntimeError(f"DataLoader workers timed out at {PREFETCHER_TIMEOUT * 20} seconds.") from None
                if err is not None:
                    raise RuntimeError(f"DataLoader worker failed on batch {tid}:\n{err}")
                pending[tid] = batch
            batch = pending.pop(done)
            done += 1
            dispatch()
            yield batch

    def shutdown(self):
        for q_ in self.index_qs:
            try:
                q_.put(None)
            except Exception:  # noqa: BLE001
                pass
        for p in self.procs:
            p.join(timeout=2)
            if p.is_alive():
                p.terminate()
        self.procs = []


def _to_device(x, device):
    if hasattr(x, "_node_frames"):
        return x.to(device) if x.device != device else x
    if torch.is_tensor(x):
        return x.to(device, non_blocking=True)
    if isinstance(x, (list, tuple)):
        return type(x)(_to_device(y, device) for y in x)
    if isinstance(x, dict):
        return {k: _to_device(v, device) for k, v in x.items()}
    return x


def _record_stream(x, stream):
    if hasattr(x, "record_stream") and not torch.is_tensor(x):
        x.record_stream(stream)
    elif torch.is_tensor(x):
        if x.is_cuda:
            x.record_stream(stream)
    elif isinstance(x, (list, tuple)):
        for y in x:
            _record_stream(y, stream)
    elif isinstance(x, dict):
        for y in x.values():
            _record_stream(y, stream)


class GraphCollator:
    """Collate samples containing graphs by batching them (reference dataloader.py:GraphCollator)."""

    def collate(self, items):
        elem = items[0]
        if hasattr(elem, "_node_frames"):
            return batch_graphs(items)
        if torch.is_tensor(elem):
            return torch.stack(items, 0)
        if isinstance(elem, (int, float)):
            return torch.tensor(items)
        if isinstance(elem, Mapping):
            return {k: self.collate([d[k] for d in items]) for k in elem}
        if isinstance(elem, (tuple, list)):
            return type(elem)(self.collate(list(s)) for s in zip(*items)) if isinstance(elem, tuple) else \
                [self.collate(list(s)) for s in zip(*items)]
        return items


class GraphDataLoader(torch.utils.data.DataLoader):
    """Batched graph-property-prediction loader (reference dataloader.py:1376)."""

    def __init__(self, dataset, collate_fn=None, use_ddp=False, ddp_seed=0, **kwargs):
        self.use_ddp = use_ddp
        if collate_fn is None:
            collate_fn = GraphCollator().collate
        if use_ddp:
            self.dist_sampler = torch.utils.data.distributed.DistributedSampler(
                dataset, shuffle=kwargs.pop("shuffle", False), drop_last=kwargs.get("drop_last", False), seed=ddp_seed)
            kwargs["sampler"] = self.dist_sampler
        super().__init__(dataset=dataset, collate_fn=collate_fn, **kwargs)

    def set_epoch(self, epoch):
        if self.use_ddp:
            self.dist_sampler.set_epoch(epoch)
        else:
            raise DGLError("set_epoch is only available when use_ddp is True.")


class DistNodeDataLoaderStub:
    """Placeholder kept for import parity; see dgl_b200.distributed.DistNodeDataLoader."""


def __getattr__(name):
    # names the reference keeps in this module but that live elsewhere in this package
    if name.startswith("__"):
        raise AttributeError(name)
    import importlib
    for modname in ('dgl_b200.dataloading.tensorized', 'dgl_b200.distributed.dist_graph', 'dgl_b200.distributed.compat'):
        mod = importlib.import_module(modname)
        if name in vars(mod):
            return vars(mod)[name]
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
