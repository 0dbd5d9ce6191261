"""Tracing / profiling built into the native-op dispatcher.

The reference has no tracing at all (no NVTX, no profiler hooks). Here every native call goes through
``_native.loader._Lib.fn``; when tracing is on, the wrapper

  * opens an NVTX range named after the C symbol (visible in ncu / nsys timelines), and
  * brackets the call with CUDA events on the current stream (CUDA ops) or ``perf_counter`` (CPU ops),

and ``report()`` prints a per-op table (calls, total / mean device time). Enable with ``DGLB200_TRACE=1`` in the
environment, ``dgl_b200.utils.trace.enable()``, or the ``trace()`` context manager; ``range(name)`` adds user ranges
(e.g. per layer / per pipeline stage). Overhead when disabled: one attribute test per native call.

    with dgl.utils.trace.trace() as t:
        model(g, x).sum().backward()
    print(t.report())
"""
from __future__ import annotations

import contextlib
import os
import time
from collections import OrderedDict

import torch

_state = {"on": False, "nvtx": True}
_records = OrderedDict()       # name -> [calls, cpu_seconds, [pending (start_event, end_event)], device_ms]


def enabled():
    return _state["on"]


def enable(nvtx=True):
    _state["on"] = True
    _state["nvtx"] = bool(nvtx)


def disable():
    _state["on"] = False


def reset():
    _records.clear()


def _rec(name):
    r = _records.get(name)
    if r is None:
        r = _records[name] = [0, 0.0, [], 0.0]
    return r


def call(name, fn, args, is_cuda):
    """Run ``fn(*args)`` under a range + timers (used by the loader when tracing is on)."""
    r = _rec(name)
    use_cuda = is_cuda and torch.cuda.is_available()
    if use_cuda and _state["nvtx"]:
        torch.cuda.nvtx.range_push(name)
    t0 = time.perf_counter()
    if use_cuda:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
    try:
        return fn(*args)
    finally:
        if use_cuda:
            e.record()
            r[2].append((s, e))
            if _state["nvtx"]:
                torch.cuda.nvtx.range_pop()
        r[0] += 1
        r[1] += time.perf_counter() - t0


@contextlib.contextmanager
def range(name):
    """User range: NVTX + timers, nested freely (``with trace.range("layer1"): ...``)."""
    if not _state["on"]:
        yield
        return
    use_cuda = torch.cuda.is_available()
    r = _rec("range:" + name)
    if use_cuda and _state["nvtx"]:
        torch.cuda.nvtx.range_push(name)
    t0 = time.perf_counter()
    if use_cuda:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
    try:
        yield
    finally:
        if use_cuda:
            e.record()
            r[2].append((s, e))
            if _state["nvtx"]:
                torch.cuda.nvtx.range_pop()
        r[0] += 1
        r[1] += time.perf_counter() - t0


def _flush():
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    for r in _records.values():
        for s, e in r[2]:
            r[3] += s.elapsed_time(e)
        r[2] = []


def summary():
    """name -> dict(calls, host_ms, device_ms) (synchronises the device to resolve pending events)."""
    _flush()
    return {k: dict(calls=v[0], host_ms=v[1] * 1e3, device_ms=v[3]) for k, v in _records.items()}


def report(sort_by="device_ms"):
    rows = sorted(summary().items(), key=lambda kv: -kv[1][sort_by])
    lines = [f"{'op':48s} {'calls':>7s} {'device ms':>11s} {'mean us':>10s} {'host ms':>10s}"]
    for k, v in rows:
        mean = v["device_ms"] * 1e3 / max(v["calls"], 1)
        lines.append(f"{k[:48]:48s} {v['calls']:7d} {v['device_ms']:11.3f} {mean:10.1f} {v['host_ms']:10.3f}")
    return "\n".join(lines)


class trace(contextlib.AbstractContextManager):
    """``with trace() as t: ...; t.report()``: enables tracing for the block (records are reset on entry)."""

    def __init__(self, nvtx=True):
        self._nvtx = nvtx

    def __enter__(self):
        self._was = _state["on"]
        reset()
        enable(self._nvtx)
        return self

    def __exit__(self, *exc):
        _state["on"] = self._was
        return False

    summary = staticmethod(summary)
    report = staticmethod(report)


if os.environ.get("DGLB200_TRACE"):
    enable()
