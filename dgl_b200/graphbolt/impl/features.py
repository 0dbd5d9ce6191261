"""Feature stores (parity: /root/reference/python/dgl/graphbolt/{feature_store.py, impl/torch_based_feature_store.py,
impl/basic_feature_store.py, impl/cpu_cached_feature.py, impl/gpu_cached_feature.py, impl/disk_based_feature?}).
Reads with GPU indices on pinned storage run as UVA gather kernels."""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np
import torch

from ..._native import loader
from ...base import DGLError
from ..base import index_select
from ..internal import futures as _fut
from .caches import FeatureCache, GPUCache


def _drain(feature, ids):
    """Run a nested ``read_async`` to completion stage by stage, yielding each stage's future to the caller."""
    gen = feature.read_async(ids)
    for _ in range(feature.read_async_num_stages(ids.device)):
        yield next(gen)


class Feature:
    def read(self, ids=None):
        raise NotImplementedError

    def read_async_num_stages(self, ids_device=None):
        """How many futures the ``read_async`` generator yields; the last one's ``wait()`` is the result
        (reference feature_store.py:60-120). Call ``next()`` that many times, interleaving other work between calls."""
        return 1

    def read_async(self, ids):
        """Staged asynchronous read; the default is one stage running ``read`` on the I/O pool."""
        yield _fut.launch(self.read, ids)

    def size(self):
        raise NotImplementedError

    def count(self):
        raise NotImplementedError

    def update(self, value, ids=None):
        raise NotImplementedError

    def metadata(self):
        return {}


class TorchBasedFeature(Feature):
    """In-memory (optionally pinned / GPU) tensor feature."""

    def __init__(self, torch_feature, metadata=None):
        assert isinstance(torch_feature, torch.Tensor), f"torch_feature in TorchBasedFeature must be torch.Tensor, but got {type(torch_feature)}."
        assert torch_feature.dim() > 1, f"dimension of torch_feature in TorchBasedFeature must be greater than 1, but got {torch_feature.dim()} dimension."
        self._tensor = torch_feature
        self._metadata = metadata or {}

    def read(self, ids=None):
        if ids is None:
            return self._tensor
        return index_select(self._tensor, ids)

    def _host_table_device_ids(self, ids_device):
        return (ids_device is not None and torch.device(ids_device).type == "cuda" and not self._tensor.is_cuda
                and not self._tensor.is_pinned())

    def read_async_num_stages(self, ids_device=None):
        # pageable host table read with device ids: host gather into pinned memory, then the H2D copy
        return 2 if self._host_table_device_ids(ids_device) else 1

    def read_async(self, ids):
        if ids is None:
            yield _fut.completed(self._tensor)
        elif self._host_table_device_ids(ids.device):
            dev = ids.device

            def gather():
                idx = ids.cpu()
                out = torch.empty((idx.shape[0],) + tuple(self._tensor.shape[1:]), dtype=self._tensor.dtype,
                                  pin_memory=True)  # caching host allocator: no cudaHostAlloc in steady state
                torch.index_select(self._tensor, 0, idx, out=out)
                return out

            host = _fut.launch(gather)
            yield host
            yield _fut.on_side_stream(lambda: host.wait().to(dev, non_blocking=True), device=dev)
        elif ids.is_cuda:
            # device table or pinned table (UVA gather kernel): enqueue on the copy stream
            yield _fut.on_side_stream(self.read, ids, device=ids.device)
        else:
            yield _fut.launch(self.read, ids)

    def size(self):
        return self._tensor.size()[1:]

    def count(self):
        return self._tensor.size()[0]

    def update(self, value, ids=None):
        if ids is None:
            self._tensor = value
        else:
            assert ids.shape[0] == value.shape[0], f"ids and value must have the same length, but got {ids.shape[0]} and {value.shape[0]}."
            if self._tensor.is_pinned() and ids.is_cuda:
                from ...ops.index import index_scatter
                index_scatter(self._tensor, ids, value.to(ids.device))
            else:
                self._tensor[ids.to(self._tensor.device)] = value.to(self._tensor.device)

    def metadata(self):
        return self._metadata

    def pin_memory_(self):
        from ...utils.pin_memory import pin_memory_inplace
        if not self._tensor.is_cuda:
            pin_memory_inplace(self._tensor)
        return self

    def is_pinned(self):
        return self._tensor.is_pinned()

    def to(self, device):
        return TorchBasedFeature(self._tensor.pin_memory() if device == "pinned" else self._tensor.to(device), self._metadata)

    def __repr__(self):
        return f"TorchBasedFeature(feature={self._tensor}, metadata={self._metadata})"


class DiskBasedFeature(Feature):
    """Feature stored in a ``.npy`` file. Rows are fetched through io_uring rings driven by raw syscalls
    (csrc/cpu/uring.cc: one ring per worker thread, ``queue_depth`` reads in flight each, short reads resumed) - the
    counterpart of the reference's OnDiskNpyArray (graphbolt/src/cnumpy.cc:146-330) - and through an OpenMP team of
    ``pread`` calls (csrc/cpu/diskio.cc) where the kernel or the container's seccomp profile refuses io_uring.
    ``io_backend``: "auto" (io_uring when available; env ``DGLB200_DISK_IO`` overrides), "uring" or "pread"."""

    def __init__(self, path, metadata=None, num_threads=None, io_backend="auto", queue_depth=128):
        self._path = path
        self._metadata = metadata or {}
        arr = np.load(path, mmap_mode="r")
        self._shape = tuple(arr.shape)
        self._np_dtype = arr.dtype
        self._dtype = torch.from_numpy(np.zeros(1, dtype=arr.dtype)).dtype
        self._offset = arr.offset
        if not arr.flags["C_CONTIGUOUS"]:
            raise DGLError("DiskBasedFeature needs a C-contiguous .npy file")
        self._row_bytes = int(np.prod(self._shape[1:], dtype=np.int64)) * arr.dtype.itemsize
        self._lib = loader.cpu_lib()
        fd = ctypes.c_int64()
        self._lib.fn("dglb200_disk_open", "sp")(path.encode(), ctypes.addressof(fd))
        self._fd = fd.value
        self._threads = num_threads or 0
        self._queue_depth = int(queue_depth)
        self.io_backend = self._pick_backend(io_backend)

    def _pick_backend(self, wanted):
        import os
        wanted = os.environ.get("DGLB200_DISK_IO", wanted) if wanted == "auto" else wanted
        if wanted not in ("auto", "uring", "pread"):
            raise DGLError(f"unknown io_backend {wanted!r}")
        if wanted == "pread":
            return "pread"
        ok, why = ctypes.c_int(), ctypes.c_int()
        self._lib.fn("dglb200_uring_available", "pp")(ctypes.addressof(ok), ctypes.addressof(why))
        if ok.value:
            return "uring"
        if wanted == "uring":
            raise DGLError(f"io_uring is not available to this process: {os.strerror(why.value)}")
        return "pread"

    def __del__(self):
        try:
            self._lib.fn("dglb200_disk_close", "l")(self._fd)
        except Exception:  # noqa: BLE001
            pass

    def read_async_num_stages(self, ids_device=None):
        return 2 if ids_device is not None and torch.device(ids_device).type == "cuda" else 1

    def read_async(self, ids):
        """Stage 1: the io_uring / pread row fetch on the I/O pool (into pinned memory when the ids are on a GPU);
        stage 2 (GPU ids only): the H2D copy on the copy stream."""
        if ids is None or not ids.is_cuda:
            yield _fut.launch(self.read, ids)
            return
        dev = ids.device
        host = _fut.launch(self._read_host, ids, True)
        yield host
        yield _fut.on_side_stream(lambda: host.wait().to(dev, non_blocking=True), device=dev)

    def read(self, ids=None):
        if ids is None:
            return torch.from_numpy(np.load(self._path))
        dev = ids.device
        out = self._read_host(ids, dev.type == "cuda")
        return out.to(dev, non_blocking=True) if dev.type == "cuda" else out

    def _read_host(self, ids, pinned):
        idx = ids.to(torch.int64).cpu().contiguous()
        # pinned staging comes from torch's caching host allocator (``.pin_memory()`` would allocate AND copy every call)
        out = torch.empty((idx.shape[0],) + self._shape[1:], dtype=self._dtype, pin_memory=bool(pinned))
        if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= self._shape[0]):
            raise DGLError(f"row index out of range for a feature with {self._shape[0]} rows")
        if self.io_backend == "uring":
            self._lib.fn("dglb200_disk_read_rows_uring", "lllplpii")(self._fd, self._offset, self._row_bytes, idx.data_ptr(),
                                                                    idx.shape[0], out.data_ptr(), self._threads, self._queue_depth)
        else:
            self._lib.fn("dglb200_disk_read_rows", "lllplpi")(self._fd, self._offset, self._row_bytes, idx.data_ptr(),
                                                              idx.shape[0], out.data_ptr(), self._threads)
        return out

    def to(self, device):
        """Disk-backed features stay on disk; only an explicit "cpu"/"pinned" target is accepted (reference
        torch_based_feature_store.py: DiskBasedFeature.to raises for GPUs — wrap it with a GPU cache instead)."""
        if str(device) not in ("cpu", "pinned"):
            raise DGLError("`DiskBasedFeature` does not support `to()` with a non-CPU device: reads land in (pinned) "
                           "host memory; put a GPU cache in front of it instead.")
        return self

    def pin_memory_(self):
        raise DGLError("`DiskBasedFeature` does not support `pin_memory_()`: there is no resident host copy to pin.")

    def is_pinned(self):
        return False

    def size(self):
        return torch.Size(self._shape[1:])

    def count(self):
        return self._shape[0]

    def update(self, value, ids=None):
        raise DGLError("DiskBasedFeature is read-only")

    def metadata(self):
        return self._metadata

    def read_into_memory(self):
        return TorchBasedFeature(self.read(), self._metadata)

    def __repr__(self):
        return f"DiskBasedFeature(path={self._path}, shape={self._shape})"


class CPUCachedFeature(Feature):
    """Wrap a (slow) fallback feature with a host cache (policy: s3-fifo | sieve | lru | clock)."""

    def __init__(self, fallback_feature, max_cache_size_in_bytes, policy=None, pin_memory=False):
        super().__init__()
        assert isinstance(fallback_feature, Feature), f"The fallback_feature must be an instance of Feature, but got {type(fallback_feature)}."
        self._fallback_feature = fallback_feature
        self.max_cache_size_in_bytes = max_cache_size_in_bytes
        feat0 = fallback_feature.read(torch.tensor([0]))
        row_bytes = feat0.element_size() * feat0[0].numel()
        cache_rows = max(1, max_cache_size_in_bytes // max(row_bytes, 1))
        self._feature = FeatureCache((cache_rows,) + tuple(feat0.shape[1:]), feat0.dtype, policy or "sieve",
                                     pin_memory=pin_memory)
        self._mutex = threading.Lock()  # lookups and insertions of overlapping mini-batches are serialised
        self._cache_bytes = cache_rows * row_bytes

    @property
    def cache_size_in_bytes(self):
        return self._cache_bytes

    def read(self, ids=None):
        if ids is None:
            return self._fallback_feature.read()
        dev = ids.device
        with self._mutex:
            values, hit_idx, miss_idx, miss_keys = self._feature.query(ids)
        out = torch.empty((ids.shape[0],) + tuple(values.shape[1:]), dtype=values.dtype, pin_memory=dev.type == "cuda")
        out[hit_idx] = values
        if miss_keys.numel():
            missing = self._fallback_feature.read(miss_keys).cpu()
            out[miss_idx] = missing
            with self._mutex:
                self._feature.replace(miss_keys, missing)
        return out.to(dev, non_blocking=True) if dev.type == "cuda" else out

    def read_async_num_stages(self, ids_device=None):
        # cache lookup, the fallback's own stages for the missing rows (host ids), combine + insert (+ H2D)
        return 2 + self._fallback_feature.read_async_num_stages(torch.device("cpu"))

    def read_async(self, ids):
        if ids is None:
            for _ in range(self.read_async_num_stages() - 1):
                yield _fut.completed(None)
            yield _fut.launch(self._fallback_feature.read)
            return
        dev = ids.device

        def lookup():
            with self._mutex:
                return self._feature.query(ids)

        looked = _fut.launch(lookup)
        yield looked
        values, hit_idx, miss_idx, miss_keys = looked.wait()
        last = None
        for last in _drain(self._fallback_feature, miss_keys.cpu()):
            yield last

        def combine():
            out = torch.empty((ids.shape[0],) + tuple(values.shape[1:]), dtype=values.dtype, pin_memory=dev.type == "cuda")
            out[hit_idx] = values
            if miss_keys.numel():
                missing = last.wait().cpu()
                out[miss_idx] = missing
                with self._mutex:
                    self._feature.replace(miss_keys, missing)
            return out

        host = _fut.launch(combine)
        if dev.type == "cuda":
            yield _fut.then(host, lambda t: t.to(dev, non_blocking=True))
        else:
            yield host

    def size(self):
        return self._fallback_feature.size()

    def count(self):
        return self._fallback_feature.count()

    def update(self, value, ids=None):
        self._fallback_feature.update(value, ids)
        if ids is not None:
            self._feature.replace(ids, value)

    def metadata(self):
        return self._fallback_feature.metadata()

    @property
    def miss_rate(self):
        return self._feature.miss_rate

    def is_pinned(self):
        return self._feature.tensor.is_pinned()


class GPUCachedFeature(Feature):
    """Wrap a fallback feature with the sm_100a set-associative GPU cache."""

    def __init__(self, fallback_feature, max_cache_size_in_bytes):
        super().__init__()
        assert isinstance(fallback_feature, Feature), f"The fallback_feature must be an instance of Feature, but got {type(fallback_feature)}."
        self._fallback_feature = fallback_feature
        self.max_cache_size_in_bytes = max_cache_size_in_bytes
        feat0 = fallback_feature.read(torch.tensor([0]))
        row_bytes = feat0.element_size() * feat0[0].numel()
        cache_rows = max(32, max_cache_size_in_bytes // max(row_bytes, 1))
        self._feature = GPUCache((cache_rows,) + tuple(feat0.shape[1:]), feat0.dtype)
        self._cache_bytes = cache_rows * row_bytes

    @property
    def cache_size_in_bytes(self):
        return self._cache_bytes

    def read(self, ids=None):
        if ids is None:
            return self._fallback_feature.read()
        values, missing_index, missing_keys = self._feature.query(ids)
        if missing_keys.numel():
            fb = self._fallback_feature
            src = missing_keys if (isinstance(fb, TorchBasedFeature) and fb.is_pinned()) or fb.read(torch.tensor([0])).is_cuda else missing_keys.cpu()
            missing_values = fb.read(src).to(values.device)
            values[missing_index] = missing_values
            self._feature.replace(missing_keys, missing_values)
        return values

    def _fallback_ids(self, missing_keys):
        fb = self._fallback_feature
        direct = (isinstance(fb, TorchBasedFeature) and (fb.is_pinned() or fb._tensor.is_cuda))
        return missing_keys if direct else missing_keys.cpu()

    def read_async_num_stages(self, ids_device=None):
        # device cache lookup, the fallback's stages for the missing rows, scatter + insert
        fb = self._fallback_feature
        direct = isinstance(fb, TorchBasedFeature) and (fb.is_pinned() or fb._tensor.is_cuda)
        return 2 + fb.read_async_num_stages(ids_device if direct else torch.device("cpu"))

    def read_async(self, ids):
        if ids is None:
            for _ in range(self.read_async_num_stages() - 1):
                yield _fut.completed(None)
            yield _fut.launch(self._fallback_feature.read)
            return
        looked = _fut.on_side_stream(self._feature.query, ids, device=ids.device, key="cache")
        yield looked
        values, missing_index, missing_keys = looked.wait()
        last = None
        for last in _drain(self._fallback_feature, self._fallback_ids(missing_keys)):
            yield last

        def fill():
            if missing_keys.numel():
                missing_values = last.wait().to(values.device, non_blocking=True)
                values[missing_index] = missing_values
                self._feature.replace(missing_keys, missing_values)
            return values

        yield _fut.on_side_stream(fill, device=ids.device, key="cache")

    def size(self):
        return self._fallback_feature.size()

    def count(self):
        return self._fallback_feature.count()

    def update(self, value, ids=None):
        self._fallback_feature.update(value, ids)
        if ids is not None:
            self._feature.replace(ids, value)

    def metadata(self):
        return self._fallback_feature.metadata()

    @property
    def miss_rate(self):
        return self._feature.miss_rate


def cpu_cached_feature(fallback_features, max_cache_size_in_bytes, policy=None, pin_memory=False):
    if isinstance(fallback_features, Feature):
        return CPUCachedFeature(fallback_features, max_cache_size_in_bytes, policy, pin_memory)
    return {k: CPUCachedFeature(v, max_cache_size_in_bytes, policy, pin_memory) for k, v in fallback_features.items()}


def gpu_cached_feature(fallback_features, max_cache_size_in_bytes):
    if isinstance(fallback_features, Feature):
        return GPUCachedFeature(fallback_features, max_cache_size_in_bytes)
    return {k: GPUCachedFeature(v, max_cache_size_in_bytes) for k, v in fallback_features.items()}


class BasicFeatureStore:
    """(domain, type_name, feature_name) -> Feature."""

    def __init__(self, features):
        self._features = dict(features)

    def __getitem__(self, feature_key):
        return self._features[feature_key]

    def __setitem__(self, feature_key, feature):
        self._features[feature_key] = feature

    def __contains__(self, feature_key):
        return feature_key in self._features

    def read(self, domain, type_name, feature_name, ids=None):
        return self._features[(domain, type_name, feature_name)].read(ids)

    def size(self, domain, type_name, feature_name):
        return self._features[(domain, type_name, feature_name)].size()

    def count(self, domain, type_name, feature_name):
        return self._features[(domain, type_name, feature_name)].count()

    def metadata(self, domain, type_name, feature_name):
        return self._features[(domain, type_name, feature_name)].metadata()

    def update(self, domain, type_name, feature_name, value, ids=None):
        self._features[(domain, type_name, feature_name)].update(value, ids)

    def keys(self):
        return list(self._features.keys())

    def __len__(self):
        return len(self._features)


class TorchBasedFeatureStore(BasicFeatureStore):
    """Feature store built from OnDiskFeatureData-like specs or a dict of tensors."""

    def __init__(self, feat_data):
        features = {}
        if isinstance(feat_data, dict):
            for key, t in feat_data.items():
                features[key] = t if isinstance(t, Feature) else TorchBasedFeature(t)
        else:
            for spec in feat_data:
                key = (spec["domain"], spec.get("type"), spec["name"])
                path = spec["path"]
                fmt = spec.get("format", "numpy")
                in_memory = spec.get("in_memory", True)
                if fmt == "torch":
                    features[key] = TorchBasedFeature(torch.load(path), spec.get("extra", {}))
                elif in_memory:
                    features[key] = TorchBasedFeature(torch.from_numpy(np.load(path)), spec.get("extra", {}))
                else:
                    features[key] = DiskBasedFeature(path, spec.get("extra", {}))
        super().__init__(features)

    def pin_memory_(self):
        for f in self._features.values():
            if hasattr(f, "pin_memory_"):
                f.pin_memory_()
        return self

    def is_pinned(self):
        """True iff every in-memory feature is pinned (disk-backed features have no host copy to pin)."""
        return all(f.is_pinned() for f in self._features.values() if hasattr(f, "is_pinned"))

    def to(self, device):
        return TorchBasedFeatureStore({k: (f.to(device) if hasattr(f, "to") else f) for k, f in self._features.items()})
