"""``CopyTo`` and the GraphBolt ``DataLoader``.

Parity: /root/reference/python/dgl/graphbolt/dataloader.py:78-191 and base.py (CopyTo). The reference rewrites the
torchdata datapipe graph to insert worker processes and prefetch buffers; sampling and gathering are GPU kernels here, so
the loader overlaps the host-side stages with compute through ONE producer thread that owns a side CUDA stream (events
order the hand-over, ``record_stream`` keeps the allocator honest) whenever a stage touches the GPU; chains of host-only
stages are moved into real worker processes in front of the feature fetcher (``MultiprocessingWrapper``), as the
reference does; ``max_uva_threads`` caps the UVA gather grid so that
it leaves SMs to the model."""
from __future__ import annotations

import queue
import threading

import torch

from .minibatch_transformer import _Stage

__all__ = ["CopyTo", "DataLoader", "MultiprocessingWrapper"]


class CopyTo(_Stage):
    def __init__(self, datapipe, device, non_blocking=False):
        super().__init__(datapipe)
        self.device = torch.device(device)
        self.non_blocking = non_blocking

    def __iter__(self):
        for data in self.datapipe:
            yield data.to(self.device, self.non_blocking) if hasattr(data, "to") else data


def _record_stream(obj, stream, depth=0):
    """Best-effort ``record_stream`` over every CUDA tensor reachable from a mini-batch (tensors, graphs / blocks,
    lists, tuples, dicts, plain objects) - the reference does the same for prefetched batches
    (python/dgl/dataloading/dataloader.py:473-485)."""
    if obj is None or depth > 6:
        return
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
        return
    if hasattr(obj, "record_stream") and not isinstance(obj, type):
        try:
            obj.record_stream(stream)
        except Exception:  # noqa: BLE001
            pass
        return
    if isinstance(obj, dict):
        for v in obj.values():
            _record_stream(v, stream, depth + 1)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_stream(v, stream, depth + 1)
    elif hasattr(obj, "__dict__"):
        for v in vars(obj).values():
            _record_stream(v, stream, depth + 1)


def _pipeline_root(datapipe):
    """The item sampler a chain of stages starts from (stages link to their input through ``.datapipe``)."""
    node = datapipe
    while hasattr(node, "datapipe"):
        node = node.datapipe
    return node


def _uses_cuda(datapipe):
    """Whether any stage of the chain works on the GPU (a device graph, a device feature store or a CopyTo)."""
    node = datapipe
    while node is not None:
        if isinstance(node, CopyTo) and torch.device(node.device).type == "cuda":
            return True
        graph = getattr(node, "graph", None)
        dev = getattr(graph, "device", None)
        if dev is not None and torch.device(dev).type == "cuda":
            return True
        node = getattr(node, "datapipe", None)
    return False


def _mp_worker(datapipe, wid, num_workers, epoch_q, result_q):
    """Body of one loader process: for every epoch announced on ``epoch_q`` iterate this worker's share of the batches
    and send ``(epoch, batch index, minibatch)``; ``None`` on the queue ends the process."""
    torch.set_num_threads(1)
    root = _pipeline_root(datapipe)
    while True:
        epoch = epoch_q.get()
        if epoch is None:
            return
        try:
            if hasattr(root, "shard_for_worker"):
                root.shard_for_worker(wid, num_workers, epoch)
            torch.manual_seed(1_000_003 * (epoch + 1) + wid)
            for i, mb in enumerate(datapipe):
                result_q.put((epoch, wid + i * num_workers, mb, None))
            result_q.put((epoch, -1 - wid, None, None))                        # this worker's end-of-epoch marker
        except Exception as e:  # noqa: BLE001
            import traceback
            result_q.put((epoch, -1 - wid, None, f"{type(e).__name__}: {e}\n{traceback.format_exc()}"))


class MultiprocessingWrapper(_Stage):
    """Run the wrapped chain of host stages (item sampling, neighbour sampling, ...) in ``num_workers`` processes
    (reference dataloader.py:48-75, which delegates to a torch DataLoader over the datapipe).

    Own design: the processes are forked once (the sampling graph and item set are shared copy-on-write), each one
    iterates the *same* chain but its item sampler only produces batches ``w, w + n, ...``; minibatches travel back
    through one queue (tensors in shared memory) and are handed out in batch order, so results are identical in
    order to the single-process loader. With ``persistent_workers`` the pool survives between epochs."""

    def __init__(self, datapipe, num_workers=0, persistent_workers=True, timeout=600.0):
        super().__init__(datapipe)
        self.num_workers, self.persistent, self.timeout = int(num_workers), bool(persistent_workers), timeout
        self._procs, self._epoch = [], 0

    def _start(self):
        import torch.multiprocessing as mp
        ctx = mp.get_context("fork")
        self._epoch_qs = [ctx.Queue() for _ in range(self.num_workers)]
        self._result_q = ctx.Queue()
        self._procs = [ctx.Process(target=_mp_worker, daemon=True,
                                   args=(self.datapipe, w, self.num_workers, self._epoch_qs[w], self._result_q))
                       for w in range(self.num_workers)]
        for p in self._procs:
            p.start()

    def shutdown(self):
        for q_ in getattr(self, "_epoch_qs", []):
            try:
                q_.put(None)
            except Exception:  # noqa: BLE001
                pass
        for p in self._procs:
            p.join(timeout=2)
            if p.is_alive():
                p.terminate()
        self._procs = []

    def __del__(self):
        try:
            self.shutdown()
        except Exception:  # noqa: BLE001
            pass

    def __iter__(self):
        if self.num_workers <= 0:
            yield from self.datapipe
            return
        if not self._procs or not all(p.is_alive() for p in self._procs):
            self.shutdown()
            self._start()
        root = _pipeline_root(self.datapipe)
        epoch = getattr(root, "_epoch", self._epoch)
        for q_ in self._epoch_qs:
            q_.put(epoch)
        pending, finished, nxt = {}, 0, 0
        try:
            while finished < self.num_workers or pending:
                while nxt in pending:
                    yield pending.pop(nxt)
                    nxt += 1
                if finished == self.num_workers:
                    if pending:
                        raise RuntimeError(f"loader workers skipped batch {nxt}")
                    break
                try:
                    ep, idx, mb, err = self._result_q.get(timeout=self.timeout)
                except queue.Empty:
                    raise RuntimeError("GraphBolt loader workers timed out or died") from None
                if ep != epoch:
                    continue                                                  # left over from an abandoned epoch
                if err is not None:
                    raise RuntimeError(f"GraphBolt loader worker failed:\n{err}")
                if idx < 0:
                    finished += 1
                else:
                    pending[idx] = mb
        finally:
            if hasattr(root, "_epoch"):
                root._epoch = epoch + 1
            self._epoch = epoch + 1
            if not self.persistent or finished < self.num_workers:
                self.shutdown()                                               # abandoned mid-epoch: restart cleanly


class DataLoader:
    """Iterate a datapipe with optional background prefetch (``num_workers`` > 0 or ``overlap`` enables a
    producer thread with its own CUDA stream; ``max_uva_threads`` caps the UVA gather grid)."""

    def __init__(self, datapipe, num_workers=0, persistent_workers=True, max_uva_threads=10240, prefetch=2):
        self.datapipe = datapipe
        self.num_workers = num_workers
        self.persistent_workers = persistent_workers
        self.prefetch = prefetch
        if max_uva_threads:
            try:
                from ..ops.cuda_struct import set_max_uva_threads
                set_max_uva_threads(max_uva_threads)
            except Exception:  # noqa: BLE001
                pass

    def __len__(self):
        return len(self.datapipe)

    def _insert_worker_processes(self):
        """Datapipe rewriting (reference dataloader.py:150-175): everything before the feature fetcher - or the whole
        chain when there is none - moves into worker processes, provided those stages are host-only. Chains that touch
        the GPU keep the producer-thread path: CUDA work cannot be spread over forked processes."""
        from .feature_fetcher import FeatureFetcher
        if getattr(self, "_rewritten", False) or _uses_cuda(self.datapipe):
            return getattr(self, "_rewritten", False)
        node, fetcher = self.datapipe, None
        while node is not None:
            if isinstance(node, FeatureFetcher):
                fetcher = node
            node = getattr(node, "datapipe", None)
        if fetcher is not None:
            fetcher.datapipe = MultiprocessingWrapper(fetcher.datapipe, self.num_workers, self.persistent_workers)
        else:
            self.datapipe = MultiprocessingWrapper(self.datapipe, self.num_workers, self.persistent_workers)
        self._rewritten = True
        return True

    def __iter__(self):
        if self.num_workers <= 0:
            yield from self.datapipe
            return
        if self._insert_worker_processes():
            yield from self.datapipe
            return
        q = queue.Queue(maxsize=self.prefetch)
        done = object()
        stop = threading.Event()
        use_cuda = torch.cuda.is_available()
        dev = torch.cuda.current_device() if use_cuda else None

        def hand_over(entry):
            """put with a stop check: a consumer that went away must not leave this thread blocked on a full queue"""
            while not stop.is_set():
                try:
                    q.put(entry, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def worker():
            try:
                stream = None
                if use_cuda:
                    torch.cuda.set_device(dev)
                    stream = torch.cuda.Stream()
                it = iter(self.datapipe)
                while not stop.is_set():
                    ev = None
                    try:
                        if stream is not None:
                            with torch.cuda.stream(stream):
                                item = next(it)
                            ev = torch.cuda.Event()
                            ev.record(stream)
                        else:
                            item = next(it)
                    except StopIteration:
                        break
                    if not hand_over((item, ev, None)):
                        return
                hand_over((done, None, None))
            except Exception as e:  # noqa: BLE001
                hand_over((None, None, e))

        th = threading.Thread(target=worker, daemon=True)
        th.start()
        try:
            while True:
                item, ev, exc = q.get()
                if exc is not None:
                    raise exc
                if item is done:
                    break
                if ev is not None:
                    cur = torch.cuda.current_stream()
                    cur.wait_event(ev)
                    _record_stream(item, cur)   # memory allocated on the producer stream is consumed on this one
                yield item
        finally:
            # normal end, an exception, or an abandoned iteration (generator closed / collected): stop the producer and wait
            # for it, so that no thread is inside a CUDA call when the interpreter tears down
            stop.set()
            th.join(timeout=10)
