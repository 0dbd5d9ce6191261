"""Asynchronous execution for the data pipeline: futures over a shared I/O thread pool and over side CUDA streams.

Counterpart of the reference's ``Future<T>`` machinery (/root/reference/graphbolt/include/graphbolt/async.h:40-220, the
``*Async`` entry points of graphbolt/src/python_binding.cc:98-233 and the ``read_async`` staging of
python/dgl/graphbolt/feature_store.py:60-120), re-designed around two facts of this framework:

* every native host routine is reached through ``ctypes``, which drops the GIL for the duration of the call, so a plain
  Python thread pool already runs gathers, disk reads, cache lookups and samplers truly in parallel with the trainer
  thread - no second C++ executor is needed;
* device-side work is ordered by CUDA streams, so a GPU future does not block the host: ``wait()`` records a
  dependency of the *current* stream on the producing side stream (event wait) and returns the tensors immediately.

``Future.wait()`` is the only synchronisation point users see, as in the reference."""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor

import torch

_lock = threading.Lock()
_pool = None
_pool_threads = None


def get_num_io_threads():
    return _pool_threads or min(max((torch.get_num_threads() + 1) // 2, 1), 8)


def set_num_io_threads(n):
    """Size of the shared I/O pool (reference ``set_num_io_uring_threads``); takes effect for work launched afterwards."""
    global _pool, _pool_threads
    n = int(n)
    if n < 1:
        raise ValueError("the I/O pool needs at least one thread")
    with _lock:
        old, _pool, _pool_threads = _pool, None, n
    if old is not None:
        old.shutdown(wait=False)


def _init_worker():
    """Pool threads run native calls side by side (lookup of batch k, disk read of k-1, combine of k-2): each takes a
    third of the CPU budget for its OpenMP team instead of a full team, so three stages in flight do not oversubscribe
    the container's quota (spinning OpenMP workers count against it)."""
    try:
        from ..._native import loader
        from ...utils.cpu_quota import effective_cpu_count
        loader.cpu_lib().fn("dglb200_cpu_set_thread_team", "i")(max(2, effective_cpu_count() // 3))
    except Exception:  # noqa: BLE001 - sizing only
        pass


def _executor():
    global _pool
    if _pool is None:
        with _lock:
            if _pool is None:
                _pool = ThreadPoolExecutor(max_workers=get_num_io_threads(), thread_name_prefix="dglb200-io",
                                           initializer=_init_worker)
    return _pool


def _after_fork_in_child():
    # threads do not survive fork(): a worker process builds its own pool on first use
    global _pool
    _pool = None


os.register_at_fork(after_in_child=_after_fork_in_child)


class Future:
    """Handle of a value that is being produced; ``wait()`` returns it (re-raising the producer's exception)."""

    __slots__ = ("_inner", "_event", "_value", "_ready", "_then")

    def __init__(self, inner=None, event=None, value=None, ready=False, then=None):
        self._inner, self._event, self._value, self._ready, self._then = inner, event, value, ready, then

    def done(self):
        if self._ready:
            return self._event is None or self._event.query()
        return self._inner.done()

    def wait(self):
        if not self._ready:
            self._value = self._inner.result()
            if self._then is not None:
                self._value = self._then(self._value)
            self._inner, self._then, self._ready = None, None, True
        if self._event is not None:
            # device result: order the caller's stream after the producer, do not stall the host
            torch.cuda.current_stream().wait_event(self._event)
            self._event = None
        return self._value


def completed(value):
    """A future that already holds its value (used for one-stage reads)."""
    return Future(value=value, ready=True)


def launch(fn, *args, **kwargs):
    """Run ``fn`` on the I/O pool. The CUDA device and the caller's stream are carried over so device work issued by
    ``fn`` is ordered exactly as if the caller had issued it."""
    if torch.cuda.is_available() and torch.cuda.is_initialized():
        dev, stream = torch.cuda.current_device(), torch.cuda.current_stream()

        def run():
            with torch.cuda.device(dev), torch.cuda.stream(stream):
                return fn(*args, **kwargs)

        return Future(inner=_executor().submit(run))
    return Future(inner=_executor().submit(fn, *args, **kwargs))


def then(future, fn):
    """Future of ``fn(future.wait())`` evaluated lazily in the waiter's thread (cheap continuations only)."""
    return Future(inner=_Chained(future), then=fn)


class _Chained:
    __slots__ = ("f",)

    def __init__(self, f):
        self.f = f

    def done(self):
        return self.f.done()

    def result(self):
        return self.f.wait()


_side_streams = {}


def side_stream(device=None, key="copy"):
    """One long-lived side stream per (device, purpose)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    k = (dev.index, key)
    s = _side_streams.get(k)
    if s is None:
        s = _side_streams[k] = torch.cuda.Stream(device=dev)
    return s


def on_side_stream(fn, *args, key="copy", device=None, **kwargs):
    """Issue ``fn`` on a side stream that first waits for the caller's stream; the returned future's ``wait()`` makes the
    caller's stream wait for the side stream's work. Tensors produced are tagged for the waiting stream."""
    cur = torch.cuda.current_stream(device)
    s = side_stream(device, key)
    s.wait_stream(cur)
    with torch.cuda.stream(s):
        out = fn(*args, **kwargs)
        ev = torch.cuda.Event()
        ev.record(s)
    _record(out, cur)
    return Future(value=out, ready=True, event=ev)


def _record(obj, stream):
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _record(o, stream)
    elif isinstance(obj, dict):
        for o in obj.values():
            _record(o, stream)
