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

    def is_pinned(self):
        return self.tensor.is_pinned()

    @property
    def max_size_in_bytes(self):
        return self.capacity * self.row_bytes

    def query_and_replace(self, keys, reader_fn):
        """Rows of ``keys``: hits come from the cache, misses are read with ``reader_fn(missing_keys)`` and inserted
        (reference impl/cpu_feature_cache.py:query_and_replace)."""
        values, hit_pos, miss_pos, miss_keys = self.query(keys)
        out = torch.empty((keys.shape[0],) + tuple(self.tensor.shape[1:]), dtype=self.tensor.dtype)
        out[hit_pos] = values
        if miss_keys.numel():
            fresh = reader_fn(miss_keys)
            out[miss_pos] = fresh.to(out.dtype)
            self.replace(miss_keys, fresh)
        return out

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
    """GPU cache of hot vertices' adjacency lists (reference: graphbolt/src/cuda/extension/gpu_graph_cache.cu,
    python/dgl/graphbolt/impl/gpu_graph_cache.py). Device-native: an open-addressing vertex table (key, access count,
    offset, length) plus an edge pool, all plain CUDA tensors (csrc/cuda/graph_cache.cu); ``query`` / ``replace`` run
    as kernels without host synchronisation, vertices are admitted once their access count reaches ``threshold``.

        hit_pos, miss_pos = cache.query(seeds)
        indptr_h, indices_h = cache.fetch_cached(seeds[hit_pos])          # compact CSC from the pool
        indptr_m, indices_m = index_select_csc(graph_indptr, graph_indices, seeds[miss_pos])   # UVA / device graph
        cache.replace(seeds[miss_pos], indptr_m, indices_m)
    """

    def __init__(self, num_edges, threshold, indptr_dtype=torch.int64, dtypes=None, device=None, num_slots=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.capacity_edges = int(num_edges)
        self.threshold = int(threshold)
        self.idtype = (dtypes[0] if dtypes else torch.int64)
        slots = int(num_slots) if num_slots else max(1024, 2 * self.capacity_edges // 8)
        self.slots = 1 << (slots - 1).bit_length()
        dev = self.device
        self.keys = torch.full((self.slots,), -1, dtype=torch.int64, device=dev)
        self.counts = torch.zeros(self.slots, dtype=torch.int32, device=dev)
        self.offs = torch.zeros(self.slots, dtype=torch.int64, device=dev)
        self.lens = torch.full((self.slots,), -1, dtype=torch.int32, device=dev)
        self.pool = torch.empty(self.capacity_edges, dtype=self.idtype, device=dev)
        self.used_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.stats = torch.zeros(2, dtype=torch.int64, device=dev)      # [queries, misses]
        self._last = None

    def _lib(self):
        from ..._native import loader
        return loader.cuda_lib(), loader

    def query(self, seeds):
        """Bump the access counters of ``seeds`` and split their positions into (cached, missing)."""
        from ...base import id_code
        seeds = seeds.to(self.device).to(self.idtype).contiguous()
        n = seeds.shape[0]
        hit = torch.empty(n, dtype=torch.uint8, device=self.device)
        off = torch.empty(n, dtype=torch.int64, device=self.device)
        ln = torch.empty(n, dtype=torch.int32, device=self.device)
        lib, loader = self._lib()
        lib.fn("dglb200_graphcache_query", "ipppplplppppp")(
            id_code(self.idtype), self.keys.data_ptr(), self.counts.data_ptr(), self.offs.data_ptr(), self.lens.data_ptr(),
            self.slots, seeds.data_ptr(), n, hit.data_ptr(), off.data_ptr(), ln.data_ptr(), self.stats.data_ptr(),
            loader.stream_ptr())
        hb = hit.bool()
        self._last = (seeds, off, ln)
        return torch.nonzero(hb).squeeze(1), torch.nonzero(~hb).squeeze(1)

    def fetch_cached(self, seeds):
        """(indptr, indices) compact CSC of cached ``seeds`` (as returned by the hit positions of ``query``)."""
        from ...base import id_code
        seeds = seeds.to(self.device).to(self.idtype).contiguous()
        n = seeds.shape[0]
        hit = torch.empty(n, dtype=torch.uint8, device=self.device)
        off = torch.empty(n, dtype=torch.int64, device=self.device)
        ln = torch.empty(n, dtype=torch.int32, device=self.device)
        lib, loader = self._lib()
        scratch = torch.zeros(2, dtype=torch.int64, device=self.device)
        cnt = self.counts.clone()      # a pure lookup must not change the admission counters
        lib.fn("dglb200_graphcache_query", "ipppplplppppp")(
            id_code(self.idtype), self.keys.data_ptr(), cnt.data_ptr(), self.offs.data_ptr(), self.lens.data_ptr(),
            self.slots, seeds.data_ptr(), n, hit.data_ptr(), off.data_ptr(), ln.data_ptr(), scratch.data_ptr(),
            loader.stream_ptr())
        indptr = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(ln.long(), 0, out=indptr[1:])
        total = int(indptr[-1])
        indices = torch.empty(total, dtype=self.idtype, device=self.device)
        if total:
            lib.fn("dglb200_cuda_seg_copy", "iipplpppppp")(
                id_code(self.idtype), 1, off.data_ptr(), 0, n, indptr.data_ptr(), self.pool.data_ptr(), indices.data_ptr(),
                0, 0, loader.stream_ptr())
        return indptr, indices

    def replace(self, seeds, indptr, indices):
        """Offer the freshly fetched adjacency lists of ``seeds``; hot ones are admitted while the pool has room."""
        from ...base import id_code
        seeds = seeds.to(self.device).to(self.idtype).contiguous()
        indptr = indptr.to(self.device).long().contiguous()
        indices = indices.to(self.device).to(self.idtype).contiguous()
        lib, loader = self._lib()
        lib.fn("dglb200_graphcache_replace", "ippppliplppplpp")(
            id_code(self.idtype), self.keys.data_ptr(), self.counts.data_ptr(), self.offs.data_ptr(), self.lens.data_ptr(),
            self.slots, self.threshold, seeds.data_ptr(), seeds.shape[0], indptr.data_ptr(), indices.data_ptr(),
            self.pool.data_ptr(), self.capacity_edges, self.used_dev.data_ptr(), loader.stream_ptr())

    @property
    def used(self):
        return int(self.used_dev.item())

    @property
    def total_queries(self):
        return int(self.stats[0].item())

    @property
    def total_miss(self):
        return int(self.stats[1].item())

    @property
    def miss_rate(self):
        q = self.total_queries
        return self.total_miss / q if q else 0.0


def index_select_csc(indptr, indices, seeds, edge_ids=None):
    """Compact CSC (indptr int64, indices[, edge ids]) of the adjacency lists of ``seeds``. ``indptr`` / ``indices`` may
    live in host-pinned memory (UVA) or on the device; one sub-warp copies one list (reference:
    graphbolt/src/cuda/index_select_csc_impl.cu)."""
    from ..._native import loader
    from ...base import id_code
    dev = seeds.device
    if dev.type != "cuda":
        s = indptr[seeds.long()].long()
        deg = indptr[seeds.long() + 1].long() - s
        out_indptr = torch.zeros(seeds.shape[0] + 1, dtype=torch.int64)
        torch.cumsum(deg, 0, out=out_indptr[1:])
        pos = torch.repeat_interleave(s - out_indptr[:-1], deg) + torch.arange(int(out_indptr[-1]))
        return (out_indptr, indices[pos]) + ((edge_ids[pos],) if edge_ids is not None else ())
    idt = indptr.dtype
    seeds = seeds.to(idt).contiguous()
    n = seeds.shape[0]
    lib = loader.cuda_lib()

    def dptr(t):
        if t.is_cuda:
            return t.data_ptr()
        from ...utils.pin_memory import host_device_pointer
        return host_device_pointer(t)
    deg = torch.empty(n, dtype=torch.int64, device=dev)
    lib.fn("dglb200_cuda_csc_degree", "ipplpp")(id_code(idt), dptr(indptr), seeds.data_ptr(), n, deg.data_ptr(), loader.stream_ptr())
    out_indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=out_indptr[1:])
    total = int(out_indptr[-1])
    out = torch.empty(total, dtype=indices.dtype, device=dev)
    out2 = torch.empty(total, dtype=indices.dtype, device=dev) if edge_ids is not None else None
    if total:
        lib.fn("dglb200_cuda_seg_copy", "iipplpppppp")(
            id_code(idt), 0, dptr(indptr), seeds.data_ptr(), n, out_indptr.data_ptr(), dptr(indices), out.data_ptr(),
            0 if edge_ids is None else dptr(edge_ids.to(indices.dtype)), 0 if out2 is None else out2.data_ptr(), loader.stream_ptr())
    return (out_indptr, out) + ((out2,) if out2 is not None else ())
