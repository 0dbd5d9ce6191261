"""Feature / graph caches.

Parity: /root/reference/graphbolt/src/{cache_policy,partitioned_cache_policy,feature_cache}.cc (CPU caches with
S3-FIFO / SIEVE / LRU / CLOCK), graphbolt/src/cuda/extension/{gpu_cache,gpu_graph_cache}.cu (HugeCTR / cuco based
GPU caches) and python/dgl/graphbolt/impl/{feature_cache,gpu_feature_cache,gpu_graph_cache}.py. The CPU policies are
native C++ (csrc/cpu/cache.cc); the GPU feature cache is an sm_100a set-associative kernel (csrc/cuda/gpu_cache.cu)."""
from __future__ import annotations

import ctypes
import enum

import torch

from ..._native import loader
from ...base import DGLError


class CachePolicyKind(enum.IntEnum):
    S3FIFO = 0
    SIEVE = 1
    LRU = 2
    CLOCK = 3


_POLICY_NAMES = {"s3-fifo": 0, "s3fifo": 0, "sieve": 1, "lru": 2, "clock": 3}


class FeatureCache:
    """Pinned-host feature cache: ``query`` returns cached rows + the positions/keys that missed; ``replace``
    inserts the missing rows (evicting per policy)."""

    def __init__(self, cache_shape, dtype, policy="sieve", num_parts=None, pin_memory=False):
        if isinstance(policy, str):
            if policy.lower() not in _POLICY_NAMES:
                raise DGLError(f"policy must be one of {list(_POLICY_NAMES)}, got {policy}")
            policy = _POLICY_NAMES[policy.lower()]
        self.capacity = int(cache_shape[0])
        self.tensor = torch.empty(tuple(cache_shape), dtype=dtype)
        if pin_memory and torch.cuda.is_available():
            self.tensor = self.tensor.pin_memory()
        self.row_bytes = self.tensor[0].numel() * self.tensor.element_size()
        lib = loader.cpu_lib()
        if num_parts is None:
            num_parts = max(1, min(8, torch.get_num_threads()))
        h = ctypes.c_int64()
        lib.fn("dglb200_cache_create", "ilip")(int(policy), self.capacity, int(num_parts), ctypes.addressof(h))
        self._h = h.value
        self._lib = lib

    def __del__(self):
        try:
            self._lib.fn("dglb200_cache_destroy", "l")(self._h)
        except Exception:  # noqa: BLE001
            pass

    def query(self, keys):
        """Returns (values_of_hits, hit_positions_in_keys, missing_positions_in_keys, missing_keys)."""
        keys = keys.to(torch.int64).cpu().contiguous()
        n = keys.shape[0]
        pos = torch.empty(n, dtype=torch.int64)
        nh = torch.zeros(1, dtype=torch.int64)
        self._lib.fn("dglb200_cache_query", "lplpp")(self._h, keys.data_ptr(), n, pos.data_ptr(), nh.data_ptr())
        hit = pos >= 0
        hit_idx = torch.nonzero(hit).squeeze(1)
        miss_idx = torch.nonzero(~hit).squeeze(1)
        vals = torch.empty((hit_idx.shape[0],) + tuple(self.tensor.shape[1:]), dtype=self.tensor.dtype)
        slots = pos[hit_idx].contiguous()
        self._lib.fn("dglb200_cpu_gather_rows", "pplpl")(self.tensor.data_ptr(), slots.data_ptr(), slots.shape[0],
                                                         vals.data_ptr(), self.row_bytes)
        return vals, hit_idx, miss_idx, keys[miss_idx]

    def replace(self, keys, values):
        keys = keys.to(torch.int64).cpu().contiguous()
        values = values.cpu().contiguous()
        n = keys.shape[0]
        pos = torch.empty(n, dtype=torch.int64)
        self._lib.fn("dglb200_cache_replace", "lplp")(self._h, keys.data_ptr(), n, pos.data_ptr())
        self._lib.fn("dglb200_cpu_scatter_rows", "pplpl")(self.tensor.data_ptr(), pos.data_ptr(), n, values.data_ptr(),
                                                          self.row_bytes)

    @property
    def stats(self):
        h, q = ctypes.c_int64(), ctypes.c_int64()
        self._lib.fn("dglb200_cache_stats", "lpp")(self._h, ctypes.addressof(h), ctypes.addressof(q))
        return h.value, q.value

    @property
    def miss_rate(self):
        h, q = self.stats
        return 1.0 - h / q if q else 0.0


class GPUCache:
    """Set-associative GPU feature cache (warp-per-set LRU-ish replacement with per-slot epochs)."""

    WAYS = 32

    def __init__(self, cache_shape, dtype, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        cap = int(cache_shape[0])
        self.num_sets = max(1, (cap + self.WAYS - 1) // self.WAYS)
        self.capacity = self.num_sets * self.WAYS
        self.feat_shape = tuple(cache_shape[1:])
        self.dtype = dtype
        self.keys = torch.full((self.capacity,), -1, dtype=torch.int64, device=self.device)
        self.stamp = torch.zeros(self.capacity, dtype=torch.int32, device=self.device)
        self.values = torch.zeros((self.capacity,) + self.feat_shape, dtype=dtype, device=self.device)
        self.locks = torch.zeros(self.num_sets, dtype=torch.int32, device=self.device)
        self._epoch = 0
        self.row_bytes = self.values[0].numel() * self.values.element_size()
        self.total_queries = 0
        self.total_miss = 0
        self._lib = loader.cuda_lib()

    def query(self, keys):
        """Returns (values [n, ...] with hit rows filled, missing_index, missing_keys)."""
        keys = keys.to(self.device, torch.int64).contiguous()
        n = keys.shape[0]
        out = torch.empty((n,) + self.feat_shape, dtype=self.dtype, device=self.device)
        hit = torch.empty(n, dtype=torch.int32, device=self.device)
        self._epoch += 1
        self._lib.fn("dglb200_gpucache_query", "pppilipplplp")(
            self.keys.data_ptr(), self.stamp.data_ptr(), self.values.data_ptr(), self._epoch, self.num_sets, self.WAYS,
            keys.data_ptr(), out.data_ptr(), n, hit.data_ptr(), self.row_bytes, loader.stream_ptr())
        miss_idx = torch.nonzero(hit == 0).squeeze(1)
        self.total_queries += n
        self.total_miss += int(miss_idx.shape[0])
        return out, miss_idx, keys[miss_idx]

    def replace(self, keys, values):
        keys = keys.to(self.device, torch.int64).contiguous()
        values = values.to(self.device, self.dtype).contiguous()
        if keys.shape[0] == 0:
            return
        # de-duplicate inside the batch (two warps inserting the same key would occupy two ways)
        uniq, inv = torch.unique(keys, return_inverse=True)
        if uniq.shape[0] != keys.shape[0]:
            first = torch.full((uniq.shape[0],), keys.shape[0], dtype=torch.long, device=self.device)
            first.scatter_reduce_(0, inv, torch.arange(keys.shape[0], device=self.device), reduce="amin")
            keys, values = uniq, values[first].contiguous()
        self._epoch += 1
        self._lib.fn("dglb200_gpucache_replace", "ppppilippllp")(
            self.keys.data_ptr(), self.stamp.data_ptr(), self.values.data_ptr(), self.locks.data_ptr(), self._epoch,
            self.num_sets, self.WAYS, keys.data_ptr(), values.data_ptr(), keys.shape[0], self.row_bytes,
            loader.stream_ptr())

    @property
    def miss_rate(self):
        return self.total_miss / self.total_queries if self.total_queries else 0.0


class GPUGraphCache:
    """GPU cache of hot vertices' adjacency lists: vertices whose access count reaches ``threshold`` get their
    neighbour lists copied into device memory; ``query`` splits seeds into cached / missing."""

    def __init__(self, num_edges, threshold, indptr_dtype, dtypes=None, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.capacity_edges = int(num_edges)
        self.threshold = int(threshold)
        self.counts = {}
        self.offsets = {}
        self.indices = torch.empty(self.capacity_edges, dtype=torch.int64, device=self.device)
        self.used = 0
        self.total_queries = 0
        self.total_miss = 0

    def query(self, seeds):
        keys = seeds.tolist()
        hit_pos, miss_pos = [], []
        for i, k in enumerate(keys):
            self.counts[k] = self.counts.get(k, 0) + 1
            (hit_pos if k in self.offsets else miss_pos).append(i)
        self.total_queries += len(keys)
        self.total_miss += len(miss_pos)
        dev = seeds.device
        return torch.tensor(hit_pos, dtype=torch.long, device=dev), torch.tensor(miss_pos, dtype=torch.long, device=dev)

    def fetch_cached(self, seeds):
        """(indptr, indices) of the cached adjacency lists of ``seeds`` (all must be cached)."""
        lens = [self.offsets[int(k)][1] for k in seeds.tolist()]
        indptr = torch.zeros(len(lens) + 1, dtype=torch.int64, device=self.device)
        indptr[1:] = torch.cumsum(torch.tensor(lens, device=self.device), 0)
        parts = [self.indices[self.offsets[int(k)][0]: self.offsets[int(k)][0] + self.offsets[int(k)][1]] for k in seeds.tolist()]
        return indptr, (torch.cat(parts) if parts else self.indices[:0])

    def replace(self, seeds, indptr, indices):
        """Offer the freshly fetched adjacency lists of ``seeds``; hot ones are admitted while space remains."""
        for i, k in enumerate(seeds.tolist()):
            if k in self.offsets or self.counts.get(k, 0) < self.threshold:
                continue
            lo, hi = int(indptr[i]), int(indptr[i + 1])
            n = hi - lo
            if self.used + n > self.capacity_edges:
                continue
            self.indices[self.used: self.used + n] = indices[lo:hi].to(self.device)
            self.offsets[k] = (self.used, n)
            self.used += n

    @property
    def miss_rate(self):
        return self.total_miss / self.total_queries if self.total_queries else 0.0
