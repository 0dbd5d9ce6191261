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


def choice(a, size, replace=True, prob=None):
    """Counterpart of :func:`numpy.random.choice` on tensors (reference python/dgl/random.py:21).

    ``a``: population size or a 1-D tensor to draw elements from; ``size``: int or shape tuple; ``prob``: optional
    1-D non-negative weights (need not be normalised). The draw runs in the native host library
    (csrc/cpu/random.cc) and is a pure function of :func:`seed` and the call index."""
    import math

    from ._native import loader
    from .base import dtype_code
    num = int(math.prod(size)) if isinstance(size, (tuple, list)) else int(size)
    population = int(a.shape[0]) if torch.is_tensor(a) else int(a)
    p = None
    if prob is not None:
        p = torch.as_tensor(prob).detach().cpu().contiguous()
        if not p.dtype.is_floating_point:
            p = p.float()
        if p.shape[0] != population:
            raise ValueError(f"prob has {p.shape[0]} entries but the population has {population}")
    out = torch.empty(num, dtype=torch.int64)
    loader.cpu_lib().fn("dglb200_cpu_choice", "llipilp")(
        num, population, 0 if p is None else dtype_code(p.dtype), 0 if p is None else p.data_ptr(), int(bool(replace)),
        next_seed(), out.data_ptr())
    if torch.is_tensor(a):
        out = a[out.to(a.device)]
    return out.view(*size) if isinstance(size, (tuple, list)) else out
