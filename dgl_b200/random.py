"""Global RNG seeding (parity: /root/reference/python/dgl/random.py, src/random/random.cc).
Native samplers use counter-based RNG streams derived from this seed, so results are
deterministic for a given (seed, call index) regardless of thread count."""
from __future__ import annotations

import itertools
import threading

import torch

_state = threading.local()
_global_seed = [0x5EED]
_counter = itertools.count()


def seed(val):
    """Seed every native random generator (samplers, random walks, negative sampling)."""
    _global_seed[0] = int(val) & 0x7FFFFFFFFFFFFFFF
    global _counter
    _counter = itertools.count()
    torch.manual_seed(int(val))


def next_seed():
    """A fresh 63-bit seed for one native call."""
    c = next(_counter)
    x = (_global_seed[0] * 0x9E3779B97F4A7C15 + (c + 1) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 31
    return x & 0x7FFFFFFFFFFFFFFF
