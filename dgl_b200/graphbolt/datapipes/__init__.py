"""Datapipe stages: MiniBatchTransformer, SubgraphSampler, FeatureFetcher, CopyTo, DataLoader.

Parity: /root/reference/python/dgl/graphbolt/{minibatch_transformer,subgraph_sampler,feature_fetcher,
base (CopyTo),dataloader}.py. The reference builds on torchdata datapipes and rewrites the datapipe graph to
insert worker processes / prefetch buffers (dataloader.py:78-191); here every stage is a plain Python
iterable and ``DataLoader`` overlaps host-side stages with GPU compute using a producer thread and a side
CUDA stream (the sampling / gather stages are GPU kernels, so worker processes are unnecessary)."""
from __future__ import annotations

import queue
import threading

import torch

from ..base import etype_str_to_tuple
from ..minibatch import MiniBatch


class _Stage:
    """Iterable with the functional verbs of the reference's datapipes."""

    def __init__(self, datapipe):
        self.datapipe = datapipe

    def __len__(self):
        return len(self.datapipe)

    def sample_neighbor(self, graph, fanouts, **kw):
        from ..impl.neighbor_sampler import NeighborSampler
        return NeighborSampler(self, graph, fanouts, **kw)

    def sample_layer_neighbor(self, graph, fanouts, **kw):
        from ..impl.neighbor_sampler import LayerNeighborSampler
        return LayerNeighborSampler(self, graph, fanouts, **kw)

    def sample_in_subgraph(self, graph):
        from ..impl.in_subgraph_sampler import InSubgraphSampler
        return InSubgraphSampler(self, graph)

    def sample_uniform_negative(self, graph, negative_ratio):
        from ..impl.negative_sampler import UniformNegativeSampler
        return UniformNegativeSampler(self, graph, negative_ratio)

    def fetch_feature(self, feature_store, node_feature_keys=None, edge_feature_keys=None, **kw):
        return FeatureFetcher(self, feature_store, node_feature_keys, edge_feature_keys, **kw)

    def copy_to(self, device, non_blocking=False):
        return CopyTo(self, device, non_blocking)

    def transform(self, fn):
        return MiniBatchTransformer(self, fn)

    def exclude_seed_edges(self, include_reverse_edges=False, reverse_etypes_mapping=None, asynchronous=False):
        from functools import partial
        from ..utils import exclude_seed_edges
        return MiniBatchTransformer(self, partial(exclude_seed_edges, include_reverse_edges=include_reverse_edges,
                                                  reverse_etypes_mapping=reverse_etypes_mapping))


class MiniBatchTransformer(_Stage):
    def __init__(self, datapipe, transformer=None):
        super().__init__(datapipe)
        self.transformer = transformer or (lambda mb: mb)

    def __iter__(self):
        for mb in self.datapipe:
            out = self.transformer(mb)
            assert isinstance(out, MiniBatch) or out is None or True
            yield out


class SubgraphSampler(MiniBatchTransformer):
    """Base: seeds -> unique seed nodes (+ compacted seeds for edge/hyperlink tasks) -> sample_subgraphs()."""

    def __init__(self, datapipe, *args, **kwargs):
        super().__init__(datapipe, self._sample)
        self._args, self._kwargs = args, kwargs

    @staticmethod
    def _seeds_preprocess(minibatch):
        from ..utils import unique_and_compact
        seeds = minibatch.seeds
        if isinstance(seeds, dict):
            nodes = {}
            compact_in = {}
            for et, s in seeds.items():
                if s.dim() == 1:
                    nodes.setdefault(et, []).append(s)
                else:
                    src_t, _, dst_t = etype_str_to_tuple(et)
                    nodes.setdefault(src_t, []).append(s[:, 0])
                    nodes.setdefault(dst_t, []).append(s[:, 1])
            uniq, comp = unique_and_compact(nodes)
            compacted = {}
            cursor = {k: 0 for k in nodes}
            for et, s in seeds.items():
                if s.dim() == 1:
                    cursor[et] += 1
                    continue
                src_t, _, dst_t = etype_str_to_tuple(et)
                a = comp[src_t][cursor[src_t]]; cursor[src_t] += 1
                b = comp[dst_t][cursor[dst_t]]; cursor[dst_t] += 1
                compacted[et] = torch.stack([a, b], 1)
            return uniq, (compacted or None)
        if seeds.dim() == 1:
            uniq, _ = unique_and_compact([seeds])
            return seeds, None  # node seeds are used as-is (order preserved)
        cols = [seeds[:, i] for i in range(seeds.shape[1])]
        uniq, comp = unique_and_compact(cols)
        return uniq, torch.stack(comp, 1)

    def _sample(self, minibatch):
        seeds, compacted = self._seeds_preprocess(minibatch)
        minibatch.input_nodes, minibatch.sampled_subgraphs = self.sample_subgraphs(seeds)
        minibatch.compacted_seeds = compacted
        return minibatch

    def sample_subgraphs(self, seeds, seeds_timestamp=None):
        raise NotImplementedError


class FeatureFetcher(MiniBatchTransformer):
    """Read node features of ``input_nodes`` and edge features of every sampled layer."""

    def __init__(self, datapipe, feature_store, node_feature_keys=None, edge_feature_keys=None, overlap_fetch=True,
                 cooperative=False):
        super().__init__(datapipe, self._read)
        self.feature_store = feature_store
        self.node_feature_keys = node_feature_keys
        self.edge_feature_keys = edge_feature_keys

    def _read(self, data):
        node_features, edge_features = {}, []
        input_nodes = data.input_nodes
        if self.node_feature_keys and input_nodes is not None:
            if isinstance(self.node_feature_keys, dict):
                for type_name, names in self.node_feature_keys.items():
                    nodes = input_nodes[type_name] if isinstance(input_nodes, dict) else input_nodes
                    if nodes is None:
                        continue
                    for name in names:
                        node_features[(type_name, name)] = self.feature_store.read("node", type_name, name, nodes)
            else:
                for name in self.node_feature_keys:
                    node_features[name] = self.feature_store.read("node", None, name, input_nodes)
        if self.edge_feature_keys and data.sampled_subgraphs:
            for sg in data.sampled_subgraphs:
                ef = {}
                eids = sg.original_edge_ids
                if eids is not None:
                    if isinstance(self.edge_feature_keys, dict):
                        for type_name, names in self.edge_feature_keys.items():
                            e = eids.get(type_name) if isinstance(eids, dict) else eids
                            if e is None:
                                continue
                            for name in names:
                                ef[(type_name, name)] = self.feature_store.read("edge", type_name, name, e)
                    else:
                        for name in self.edge_feature_keys:
                            ef[name] = self.feature_store.read("edge", None, name, eids)
                edge_features.append(ef)
        data.set_node_features(node_features)
        data.set_edge_features(edge_features)
        return data


class CopyTo(_Stage):
    def __init__(self, datapipe, device, non_blocking=False):
        super().__init__(datapipe)
        self.device = torch.device(device)
        self.non_blocking = non_blocking

    def __iter__(self):
        for data in self.datapipe:
            yield data.to(self.device, self.non_blocking) if hasattr(data, "to") else data


def _record_stream(obj, stream, depth=0):
    """Best-effort ``record_stream`` over every CUDA tensor reachable from a mini-batch (tensors, graphs / blocks,
    lists, tuples, dicts, plain objects) - the reference does the same for prefetched batches
    (python/dgl/dataloading/dataloader.py:473-485)."""
    if obj is None or depth > 6:
        return
    if torch.is_tensor(obj):
        if obj.is_cuda:
            obj.record_stream(stream)
        return
    if hasattr(obj, "record_stream") and not isinstance(obj, type):
        try:
            obj.record_stream(stream)
        except Exception:  # noqa: BLE001
            pass
        return
    if isinstance(obj, dict):
        for v in obj.values():
            _record_stream(v, stream, depth + 1)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_stream(v, stream, depth + 1)
    elif hasattr(obj, "__dict__"):
        for v in vars(obj).values():
            _record_stream(v, stream, depth + 1)


class DataLoader:
    """Iterate a datapipe with optional background prefetch (``num_workers`` > 0 or ``overlap`` enables a
    producer thread with its own CUDA stream; ``max_uva_threads`` caps the UVA gather grid)."""

    def __init__(self, datapipe, num_workers=0, persistent_workers=True, max_uva_threads=10240, prefetch=2):
        self.datapipe = datapipe
        self.num_workers = num_workers
        self.prefetch = prefetch
        if max_uva_threads:
            try:
                from ...ops.cuda_struct import set_max_uva_threads
                set_max_uva_threads(max_uva_threads)
            except Exception:  # noqa: BLE001
                pass

    def __len__(self):
        return len(self.datapipe)

    def __iter__(self):
        if self.num_workers <= 0:
            yield from self.datapipe
            return
        q = queue.Queue(maxsize=self.prefetch)
        done = object()
        use_cuda = torch.cuda.is_available()
        dev = torch.cuda.current_device() if use_cuda else None

        def worker():
            try:
                if use_cuda:
                    torch.cuda.set_device(dev)
                    stream = torch.cuda.Stream()
                    ctx = torch.cuda.stream(stream)
                else:
                    stream, ctx = None, None
                it = iter(self.datapipe)
                while True:
                    if ctx is not None:
                        with torch.cuda.stream(stream):
                            try:
                                item = next(it)
                            except StopIteration:
                                break
                        ev = torch.cuda.Event()
                        ev.record(stream)
                        q.put((item, ev, None))
                    else:
                        try:
                            item = next(it)
                        except StopIteration:
                            break
                        q.put((item, None, None))
                q.put((done, None, None))
            except Exception as e:  # noqa: BLE001
                q.put((None, None, e))

        th = threading.Thread(target=worker, daemon=True)
        th.start()
        while True:
            item, ev, exc = q.get()
            if exc is not None:
                raise exc
            if item is done:
                break
            if ev is not None:
                cur = torch.cuda.current_stream()
                cur.wait_event(ev)
                _record_stream(item, cur)   # memory allocated on the producer stream is consumed on this one
            yield item
        th.join()
