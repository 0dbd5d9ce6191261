"""``CopyTo`` and the GraphBolt ``DataLoader``.

Parity: /root/reference/python/dgl/graphbolt/dataloader.py:78-191 and base.py (CopyTo). The reference rewrites the
torchdata datapipe graph to insert worker processes and prefetch buffers; sampling and gathering are GPU kernels here, so
the loader overlaps the host-side stages with compute through ONE producer thread that owns a side CUDA stream (events
order the hand-over, ``record_stream`` keeps the allocator honest); ``max_uva_threads`` caps the UVA gather grid so that
it leaves SMs to the model."""
from __future__ import annotations

import queue
import threading

import torch

from .minibatch_transformer import _Stage

__all__ = ["CopyTo", "DataLoader"]


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


class DataLoader:
    """Iterate a datapipe with optional background prefetch (``num_workers`` > 0 or ``overlap`` enables a
    producer thread with its own CUDA stream; ``max_uva_threads`` caps the UVA gather grid)."""

    def __init__(self, datapipe, num_workers=0, persistent_workers=True, max_uva_threads=10240, prefetch=2):
        self.datapipe = datapipe
        self.num_workers = num_workers
        self.prefetch = prefetch
        if max_uva_threads:
            try:
                from ..ops.cuda_struct import set_max_uva_threads
                set_max_uva_threads(max_uva_threads)
            except Exception:  # noqa: BLE001
                pass

    def __len__(self):
        return len(self.datapipe)

    def __iter__(self):
        if self.num_workers <= 0:
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
