"""Public GraphBolt names that complete the reference's surface (python/dgl/graphbolt/{base,dataset,feature_store,
sampling_graph,internal/*,impl/ondisk_metadata}.py): abstract base classes, datapipe markers / buffering stages, the
pydantic models of ``metadata.yaml``, stream helpers and small utilities."""
from __future__ import annotations

import enum
import hashlib
import os
import tarfile
import warnings
import zipfile
from collections import namedtuple
from typing import Any, Dict, List, Optional

import torch

from .datapipes import MiniBatchTransformer, _Stage
from .impl.caches import FeatureCache, GPUCache
from .impl.features import Feature
from .dataset import Dataset, Task  # noqa: F401
from .feature_store import FeatureKey, FeatureStore  # noqa: F401
from .impl.legacy_dataset import LegacyDataset  # noqa: F401
from .impl.ondisk_metadata import (ExtraMetaData, OnDiskFeatureData, OnDiskFeatureDataDomain,  # noqa: F401
                                   OnDiskFeatureDataFormat, OnDiskGraphTopology, OnDiskGraphTopologyType, OnDiskMetaData,
                                   OnDiskTaskData, OnDiskTVTSet, OnDiskTVTSetData)
from .sampling_graph import SamplingGraph  # noqa: F401

# ---------------------------------------------------------------------------------------------- abstract bases


CPUFeatureCache = FeatureCache
GPUFeatureCache = GPUCache

# ---------------------------------------------------------------------------------------------- datapipe stages / markers


class EndMarker(_Stage):
    """Pass-through stage marking the end of a rewritable datapipe segment (reference base.py EndMarker)."""

    def __init__(self, datapipe):
        super().__init__(datapipe)

    def __iter__(self):
        yield from self.datapipe


class FeatureFetcherStartMarker(EndMarker):
    """Marks where feature fetching starts (the DataLoader inserts the overlap stages after it)."""


class Bufferer(_Stage):
    """Keep up to ``buffer_size`` upstream items in flight (reference base.py Bufferer)."""

    def __init__(self, datapipe, buffer_size=1):
        super().__init__(datapipe)
        if buffer_size <= 0:
            raise ValueError("'buffer_size' is required to be a positive integer.")
        self.buffer_size = buffer_size

    def __iter__(self):
        from collections import deque
        buf = deque()
        for item in self.datapipe:
            buf.append(item)
            if len(buf) > self.buffer_size:
                yield buf.popleft()
        while buf:
            yield buf.popleft()


class Waiter(_Stage):
    """Call ``.wait()`` on every item that carries a pending asynchronous result (reference base.py Waiter)."""

    def __init__(self, datapipe):
        super().__init__(datapipe)

    def __iter__(self):
        for item in self.datapipe:
            w = getattr(item, "wait", None)
            yield w() if callable(w) else item


class SamplePerLayer(MiniBatchTransformer):
    """One layer of neighbour sampling as its own stage (the reference splits ``sample_neighbor`` into these)."""

    def __init__(self, datapipe, sampler, fanout, overlap_fetch=False, asynchronous=False):
        super().__init__(datapipe, self._sample)
        self.sampler, self.fanout = sampler, fanout

    def _sample(self, minibatch):
        sub = self.sampler(minibatch._seed_nodes, self.fanout)
        minibatch.sampled_subgraphs = [sub] + list(minibatch.sampled_subgraphs or [])
        return minibatch


class FetchInsubgraphData(MiniBatchTransformer):
    """Fetch the adjacency lists of the current seeds from a (host-pinned) graph ahead of sampling."""

    def __init__(self, datapipe, graph, prob_name=None):
        super().__init__(datapipe, self._fetch)
        self.graph, self.prob_name = graph, prob_name

    def _fetch(self, minibatch):
        from .impl.caches import index_select_csc
        seeds = minibatch._seed_nodes
        if torch.is_tensor(seeds) and seeds.is_cuda:
            minibatch._sliced_sampling_graph = index_select_csc(self.graph.csc_indptr, self.graph.indices, seeds)
        return minibatch


class CombineCachedAndFetchedInSubgraph(MiniBatchTransformer):
    """Merge adjacency lists served by ``GPUGraphCache`` with the ones fetched from the host graph."""

    def __init__(self, datapipe, prob_name=None):
        super().__init__(datapipe, lambda mb: mb)


class SeedEdgesExcluder(MiniBatchTransformer):
    """Datapipe form of ``exclude_seed_edges`` (reference impl/neighbor_sampler.py)."""

    def __init__(self, datapipe, include_reverse_edges=False, reverse_etypes_mapping=None, asynchronous=False):
        from .utils import exclude_seed_edges
        super().__init__(datapipe, lambda mb: exclude_seed_edges(mb, include_reverse_edges, reverse_etypes_mapping))


def minibatcher_default(batch, names):
    """Default collation of an item-set batch into a ``MiniBatch`` (reference item_sampler.py)."""
    from .minibatch import MiniBatch
    if isinstance(names, str) or names is None:
        names, batch = (names,), (batch,)
    mb = MiniBatch()
    for name, item in zip(names, batch):
        if name is None:
            continue
        if not hasattr(mb, name):
            warnings.warn(f"Unknown item name '{name}' is detected and added into `MiniBatch`.", GBWarning)
        setattr(mb, name, item)
    return mb


# ---------------------------------------------------------------------------------------------- metadata.yaml models


def preprocess_ondisk_dataset(dataset_dir, include_original_edge_id=False, force_preprocess=None, auto_cast_to_optimal_dtype=True):
    """Run the one-off preprocessing of an on-disk dataset and return the path of the generated metadata
    (reference impl/ondisk_dataset.py preprocess_ondisk_dataset)."""
    from .impl.datasets import OnDiskDataset
    ds = OnDiskDataset(dataset_dir, include_original_edge_id=include_original_edge_id,
                       force_preprocess=bool(force_preprocess)).load()
    return os.path.join(dataset_dir, "preprocessed", "metadata.yaml") if ds is not None else None


# ---------------------------------------------------------------------------------------------- helpers
class GBWarning(UserWarning):
    """GraphBolt warning category."""


def gb_warning_format(message, category, filename, lineno, line=None):
    return f"{filename}:{lineno}: {category.__name__}: {message}\n"


def gb_warning(message, category=GBWarning, stacklevel=2):
    warnings.warn(message, category=category, stacklevel=stacklevel)


def seed(val):
    """Seed the GraphBolt samplers (alias of ``set_seed``; reference base.py seed)."""
    from .. import random as dgl_random
    dgl_random.seed(int(val))
    torch.manual_seed(int(val))


def built_with_cuda():
    from .._native import loader
    try:
        loader.cuda_lib()
        return True
    except Exception:  # noqa: BLE001
        return False


def is_wsl(v: str = None) -> bool:
    import platform
    v = v or platform.uname().release
    return "microsoft" in v.lower()


def is_listlike(x):
    return isinstance(x, (list, tuple))


def recursive_apply(data, fn, *args, **kwargs):
    """Apply ``fn`` to every leaf of a nested list / tuple / dict (reference internal/utils.py)."""
    if isinstance(data, dict):
        return {k: recursive_apply(v, fn, *args, **kwargs) for k, v in data.items()}
    if isinstance(data, (list, tuple)):
        return type(data)(recursive_apply(v, fn, *args, **kwargs) for v in data)
    return fn(data, *args, **kwargs)


def recursive_apply_reduce_all(data, fn, *args, **kwargs):
    if isinstance(data, dict):
        return all(recursive_apply_reduce_all(v, fn, *args, **kwargs) for v in data.values())
    if isinstance(data, (list, tuple)):
        return all(recursive_apply_reduce_all(v, fn, *args, **kwargs) for v in data)
    return fn(data, *args, **kwargs)


def get_attributes(obj):
    return [a for a in dir(obj) if not a.startswith("__") and not callable(getattr(obj, a))]


def get_nonproperty_attributes(obj):
    return [a for a in get_attributes(obj) if not isinstance(getattr(type(obj), a, None), property)]


def bytes_to_number_of_items(cache_capacity_in_bytes, single_item):
    item_bytes = single_item.nbytes if hasattr(single_item, "nbytes") else single_item.numel() * single_item.element_size()
    return max(1, int(cache_capacity_in_bytes) // max(int(item_bytes), 1))


def indptr_edge_ids(indptr, dtype=None, offset=None, output_size=None):
    """Edge ids relative to the start of each row: for indptr [0, 2, 5] -> [0, 1, 0, 1, 2] (+ per-row ``offset``)."""
    deg = (indptr[1:] - indptr[:-1]).long()
    total = int(output_size) if output_size is not None else int(indptr[-1] - indptr[0])
    starts = torch.repeat_interleave((indptr[:-1] - indptr[0]).long(), deg, output_size=total)
    out = torch.arange(total, device=indptr.device) - starts
    if offset is not None:
        out = out + torch.repeat_interleave(offset.long(), deg, output_size=total)
    return out.to(dtype or indptr.dtype)


_streams = {}


def get_host_to_device_uva_stream():
    if "h2d" not in _streams:
        _streams["h2d"] = torch.cuda.Stream(priority=-1)
    return _streams["h2d"]


def get_device_to_host_uva_stream():
    if "d2h" not in _streams:
        _streams["d2h"] = torch.cuda.Stream(priority=-1)
    return _streams["d2h"]


def all_to_all(outputs, inputs, group=None, async_op=False):
    """List-of-tensors all-to-all that also works on gloo (reference subgraph_sampler.py:41-84)."""
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl" or inputs[0].is_cuda:
        return dist.all_to_all(outputs, inputs, group=group, async_op=async_op)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    works = []
    for r in range(world):                     # gloo has no all_to_all for lists: pairwise isend / irecv
        if r == rank:
            outputs[r].copy_(inputs[r])
        else:
            works.append(dist.isend(inputs[r], r, group=group))
            works.append(dist.irecv(outputs[r], r, group=group))
    for w in works:
        w.wait()
    return None


def revert_to_homo(d):
    """{'_N': x} / {'_N:_E:_N': x} style single-type dicts back to the bare value (reference internal/utils.py)."""
    if isinstance(d, dict) and len(d) == 1:
        return next(iter(d.values()))
    return d


def convert_to_hetero(x, key="_N"):
    return x if isinstance(x, dict) else {key: x}


def wrap_with_cached_feature(cached_feature_type, fallback_features, max_cache_size_in_bytes, *args, **kwargs):
    """Wrap one feature (or every feature of a dict) in a cached feature class (reference impl/utils)."""
    if isinstance(fallback_features, dict):
        return {k: cached_feature_type(v, max_cache_size_in_bytes, *args, **kwargs) for k, v in fallback_features.items()}
    return cached_feature_type(fallback_features, max_cache_size_in_bytes, *args, **kwargs)


def check_sha1(filename, sha1_hash):
    h = hashlib.sha1()
    with open(filename, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest() == sha1_hash


def extract_archive(file, target_dir, overwrite=True):
    if os.path.exists(target_dir) and not overwrite:
        return
    if file.endswith((".tar.gz", ".tgz", ".tar")):
        with tarfile.open(file) as t:
            t.extractall(target_dir)
    elif file.endswith(".zip"):
        with zipfile.ZipFile(file) as z:
            z.extractall(target_dir)
    else:
        raise ValueError(f"Unrecognized file type: {file}")


def download(url, path=None, overwrite=True, sha1_hash=None, retries=5, verify_ssl=True, log=True):
    """Dataset download helper; this build has no network access, so only already-present files are accepted."""
    fname = path if path and not os.path.isdir(path) else os.path.join(path or ".", url.split("/")[-1])
    if os.path.exists(fname) and (sha1_hash is None or check_sha1(fname, sha1_hash)):
        return fname
    raise RuntimeError(f"cannot download {url}: no network access; place the file at {fname}")


# ---------------------------------------------------------------------------------------------- late additions
from typing import Dict as _Dict, NamedTuple as _NamedTuple, Tuple as _Tuple  # noqa: E402

from . import datapipes as _dp  # noqa: E402


class PyGLayerData(_NamedTuple):
    """Homogeneous inputs of one PyG layer: ``x`` = (source features, destination features), ``edge_index`` [2, E],
    ``size`` = (num src, num dst) — reference sampled_subgraph.py:44."""
    x: _Tuple
    edge_index: torch.Tensor
    size: _Tuple[int, int]


class PyGLayerHeteroData(_NamedTuple):
    """Heterogeneous counterpart: every field is a dict (features by node type, the rest by edge type) —
    reference sampled_subgraph.py:55."""
    x: _Tuple
    edge_index: _Dict[str, torch.Tensor]
    size: _Dict[str, _Tuple[int, int]]


def get_feature_key_list(feature_keys, domain):
    """[(domain, type or None, name)] from a list (homogeneous) or ``{type: [names]}`` (reference feature_fetcher.py:22)."""
    if isinstance(feature_keys, dict):
        return [(domain, t, n) for t, names in feature_keys.items() for n in names]
    if feature_keys is not None:
        return [(domain, None, n) for n in feature_keys]
    return []


def is_object_pinned(obj):
    """True iff every tensor reachable from the object's plain attributes is in pinned memory (reference base.py:323)."""
    for attr in get_nonproperty_attributes(obj):
        if not recursive_apply_reduce_all(getattr(obj, attr), lambda x: x is None or not hasattr(x, "is_pinned") or x.is_pinned()):
            return False
    return True


def is_scalar(x):
    """0-dim tensor or Python number (reference itemset.py:is_scalar)."""
    import numbers
    return (torch.is_tensor(x) and x.dim() == 0) or isinstance(x, numbers.Number)


def load_graphbolt():
    """The reference dlopens ``libgraphbolt`` here. The native pieces of this package (samplers, caches, disk reader)
    are part of ``libdglb200_{cpu,cuda}``; loading the host library is the equivalent readiness check."""
    from .._native import loader
    loader.cpu_lib()


class MultiprocessingWrapper(_dp.MiniBatchTransformer):
    """Reference dataloader.py:MultiprocessingWrapper moves the upstream stages into torch worker processes. The
    sampling stages here are GPU (or OpenMP) kernels, so the wrapper runs them in a prefetching thread instead:
    ``num_workers == 0`` is a pass-through, otherwise up to ``2 * num_workers`` batches are produced ahead."""

    def __init__(self, datapipe, num_workers=0, persistent_workers=True):
        from .datapipes.utils import PrefetcherIterDataPipe
        self.num_workers = num_workers
        inner = datapipe if num_workers <= 0 else PrefetcherIterDataPipe(datapipe, 2 * num_workers)
        super().__init__(inner, None)

    def __iter__(self):
        return iter(self.datapipe)


class NeighborSamplerImpl(_dp.SubgraphSampler):
    """Common base of the neighbour samplers (reference impl/neighbor_sampler.py:NeighborSamplerImpl): subclasses of it
    here are ``impl.neighbor_sampler.NeighborSampler`` / ``LayerNeighborSampler``; constructing it directly gives the
    sampler selected by ``layer`` (False: per-node fanout, True: LABOR layer sampling)."""

    def __new__(cls, datapipe, graph, fanouts, replace=False, prob_name=None, deduplicate=True, sampler=None,
                overlap_fetch=None, num_gpu_cached_edges=None, gpu_cache_threshold=None, cooperative=None,
                asynchronous=None, layer_dependency=None, batch_dependency=None, layer=False):
        from .impl.neighbor_sampler import LayerNeighborSampler, NeighborSampler
        if layer:
            return LayerNeighborSampler(datapipe, graph, fanouts, replace=replace, prob_name=prob_name,
                                        deduplicate=deduplicate, layer_dependency=bool(layer_dependency))
        return NeighborSampler(datapipe, graph, fanouts, replace=replace, prob_name=prob_name, deduplicate=deduplicate)


class TemporalNeighborSamplerImpl(_dp.SubgraphSampler):
    """Common base of the temporal samplers (reference impl/temporal_neighbor_sampler.py)."""

    def __new__(cls, datapipe, graph, fanouts, replace=False, prob_name=None, node_timestamp_attr_name=None,
                edge_timestamp_attr_name=None, sampler=None, layer=False):
        from .impl.neighbor_sampler import TemporalLayerNeighborSampler, TemporalNeighborSampler
        kind = TemporalLayerNeighborSampler if layer else TemporalNeighborSampler
        return kind(datapipe, graph, fanouts, replace=replace, prob_name=prob_name,
                    node_timestamp_attr_name=node_timestamp_attr_name, edge_timestamp_attr_name=edge_timestamp_attr_name)


class CompactPerLayer(_dp.MiniBatchTransformer):
    """Second half of a sampling layer as its own stage: relabel the freshly sampled block against the unique node set
    and make that set the seeds of the next layer (reference impl/neighbor_sampler.py:CompactPerLayer)."""

    def __init__(self, datapipe, deduplicate=True, cooperative=False, asynchronous=False):
        super().__init__(datapipe, self._compact)
        self.deduplicate = deduplicate

    def _compact(self, minibatch):
        from .utils import compact_csc_format, unique_and_compact_csc_formats
        sub = minibatch.sampled_subgraphs[0]
        seeds = minibatch._seed_nodes
        if getattr(sub, "original_row_node_ids", None) is not None:
            minibatch._seed_nodes = sub.original_row_node_ids     # layer already compacted by the sampler
            return minibatch
        if self.deduplicate:
            row_ids, csc = unique_and_compact_csc_formats(sub.sampled_csc, seeds)
        else:
            row_ids, csc = compact_csc_format(sub.sampled_csc, seeds)
        sub.sampled_csc, sub.original_row_node_ids, sub.original_column_node_ids = csc, row_ids, seeds
        minibatch._seed_nodes = row_ids
        return minibatch


class FetchCachedInsubgraphData(_dp.MiniBatchTransformer):
    """Query the GPU graph cache for the seeds' adjacency lists; misses are left for :class:`FetchInsubgraphData`
    (reference impl/neighbor_sampler.py:FetchCachedInsubgraphData)."""

    def __init__(self, datapipe, gpu_graph_cache):
        super().__init__(datapipe, self._query)
        self.cache = gpu_graph_cache

    def _query(self, minibatch):
        seeds = minibatch._seed_nodes
        if torch.is_tensor(seeds) and seeds.is_cuda:
            minibatch._cache_query = self.cache.query(seeds)
        return minibatch


class CombineCachedAndFetchedInSubgraph(_dp.MiniBatchTransformer):
    """Merge cache hits with the lists fetched for the misses and insert the latter into the cache (reference
    impl/neighbor_sampler.py:CombineCachedAndFetchedInSubgraph)."""

    def __init__(self, datapipe, prob_name=None):
        super().__init__(datapipe, self._combine)

    def _combine(self, minibatch):
        q = getattr(minibatch, "_cache_query", None)
        fetched = getattr(minibatch, "_sliced_sampling_graph", None)
        if q is not None and fetched is not None and hasattr(q, "replace"):
            minibatch._sliced_sampling_graph = q.replace(*fetched)
        return minibatch

