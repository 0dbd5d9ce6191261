"""Inspecting and rewriting a datapipe chain (parity: /root/reference/python/dgl/graphbolt/datapipes/utils.py:
datapipe_graph_to_adjlist, find_dps, replace_dp) plus the thread-prefetch stage (reference
``torchdata`` ``Prefetcher`` wrapped as PrefetcherIterDataPipe).

A datapipe here is any iterable object; its upstream datapipes are the attributes that are themselves datapipes
(``datapipe`` for every built-in stage, but user stages may hold several). That attribute scan is the whole graph
traversal, so these helpers work on user-defined stages without registration."""
from __future__ import annotations

import queue
import threading

__all__ = ["datapipe_graph_to_adjlist", "find_dps", "replace_dp", "PrefetcherIterDataPipe", "traverse_dps"]


def _is_datapipe(obj):
    from . import DataLoader, _Stage
    from ..itemset import ItemSampler
    return isinstance(obj, (_Stage, ItemSampler, PrefetcherIterDataPipe)) and not isinstance(obj, DataLoader)


def _upstreams(dp):
    """(attribute name, datapipe) pairs feeding ``dp``."""
    out = []
    for name, val in list(vars(dp).items()):
        if _is_datapipe(val):
            out.append((name, None, val))
        elif isinstance(val, (list, tuple)):
            out.extend((name, i, v) for i, v in enumerate(val) if _is_datapipe(v))
    return out


def traverse_dps(datapipe):
    """Nested ``{id: (datapipe, {upstream id: (...)})}`` like ``torch.utils.data.graph.traverse_dps``."""
    return {id(datapipe): (datapipe, {k: v for _, _, up in _upstreams(datapipe) for k, v in traverse_dps(up).items()})}


def datapipe_graph_to_adjlist(datapipe_graph):
    """Flatten :func:`traverse_dps` output to ``{id: (datapipe, [upstream ids])}``."""
    adj = {}

    def walk(level):
        for key, (dp, ups) in level.items():
            if key not in adj:
                adj[key] = (dp, list(ups.keys()))
                walk(ups)

    walk(datapipe_graph)
    return adj


def find_dps(graph, dp_type):
    """Every datapipe of type ``dp_type`` in a :func:`traverse_dps` graph (most downstream first)."""
    found, seen = [], set()

    def walk(level):
        for key, (dp, ups) in level.items():
            if key in seen:
                continue
            seen.add(key)
            if isinstance(dp, dp_type):
                found.append(dp)
            walk(ups)

    walk(graph)
    return found


def replace_dp(graph, old_datapipe, new_datapipe):
    """Rewire every consumer of ``old_datapipe`` to read from ``new_datapipe``; returns the refreshed graph.
    When ``old_datapipe`` is the sink itself the new sink's graph is returned."""
    roots = [dp for dp, _ in graph.values()]
    if any(dp is old_datapipe for dp in roots):
        return traverse_dps(new_datapipe)
    seen = set()

    def walk(dp):
        if id(dp) in seen or dp is new_datapipe:
            return
        seen.add(id(dp))
        for name, pos, up in _upstreams(dp):
            if up is old_datapipe:
                if pos is None:
                    setattr(dp, name, new_datapipe)
                else:
                    seq = list(getattr(dp, name))
                    seq[pos] = new_datapipe
                    setattr(dp, name, type(getattr(dp, name))(seq))
            else:
                walk(up)

    for root in roots:
        walk(root)
    return {id(r): traverse_dps(r)[id(r)] for r in roots}


class PrefetcherIterDataPipe:
    """Run the upstream datapipe in a background thread, keeping up to ``buffer_size`` items ready."""

    _DONE = object()

    def __init__(self, source_datapipe, buffer_size=10):
        if buffer_size <= 0:
            raise ValueError("'buffer_size' is required to be a positive integer.")
        self.source_datapipe = self.datapipe = source_datapipe
        self.buffer_size = buffer_size

    def __len__(self):
        return len(self.datapipe)

    def __iter__(self):
        q = queue.Queue(maxsize=self.buffer_size)
        stop = threading.Event()

        def produce():
            try:
                for item in self.datapipe:
                    while not stop.is_set():
                        try:
                            q.put((item, None), timeout=0.05)
                            break
                        except queue.Full:
                            continue
                    if stop.is_set():
                        return
                q.put((self._DONE, None))
            except BaseException as exc:  # noqa: BLE001 - forwarded to the consumer
                q.put((None, exc))

        th = threading.Thread(target=produce, daemon=True)
        th.start()
        try:
            while True:
                item, exc = q.get()
                if exc is not None:
                    raise exc
                if item is self._DONE:
                    return
                yield item
        finally:
            stop.set()
