"""Process helpers (parity: /root/reference/python/dgl/multiprocessing/pytorch.py:59-104): tensors shared across the
local ranks through named POSIX shared memory; the segment name is broadcast over torch.distributed."""
from __future__ import annotations

import os
import uuid

import torch
import torch.distributed as dist

from ..utils.shared_mem import create_shared_tensor, open_shared_tensor


def call_once_and_share(func, shape, dtype, rank=0):
    """Rank ``rank`` computes ``func()``; everybody gets a shared-memory view of the result."""
    current = dist.get_rank() if dist.is_initialized() else 0
    name_list = [None]
    if current == rank:
        name_list[0] = "/dglb200_" + uuid.uuid4().hex[:16]
        t = create_shared_tensor(name_list[0], shape, dtype)
        t.copy_(func())
    if dist.is_initialized():
        dist.broadcast_object_list(name_list, src=rank)
        dist.barrier()
    if current != rank:
        t = open_shared_tensor(name_list[0], shape, dtype)
    return t


def shared_tensor(shape, dtype=torch.float32):
    """Uninitialised tensor shared by all local ranks."""
    return call_once_and_share(lambda: torch.zeros(shape, dtype=dtype), shape, dtype)


from torch.multiprocessing import *  # noqa: F401,F403,E402


def thread_wrapped_func(func):
    """Run a process entry point on a fresh thread and hand back its result (reference multiprocessing/pytorch.py:13).
    OpenMP runtimes keep per-thread state that does not survive ``fork``; entering through a new thread gives the
    child a clean OpenMP team. Exceptions are re-raised in the caller with the worker's traceback as message."""
    import functools
    import threading
    import traceback

    @functools.wraps(func)
    def decorated(*args, **kwargs):
        box = {}

        def run():
            try:
                box["res"] = func(*args, **kwargs)
            except Exception as exc:  # noqa: BLE001
                box["exc"], box["trace"] = exc, traceback.format_exc()

        t = threading.Thread(target=run)
        t.start()
        t.join()
        if "exc" in box:
            try:
                raise box["exc"].__class__(box["trace"])
            except TypeError:
                raise RuntimeError(box["trace"]) from box["exc"]
        return box.get("res")

    return decorated


class Process(torch.multiprocessing.Process):
    """``torch.multiprocessing.Process`` whose target is wrapped by :func:`thread_wrapped_func`."""

    def __init__(self, group=None, target=None, name=None, args=(), kwargs=None, *, daemon=None):
        if target is not None:
            target = thread_wrapped_func(target)
        super().__init__(group, target, name, args, kwargs or {}, daemon=daemon)
