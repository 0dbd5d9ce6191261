"""Futures, staged ``read_async`` and the overlapped FeatureFetcher (graphbolt/internal/futures.py, impl/features.py,
feature_fetcher.py): every asynchronous route must return exactly what the synchronous one returns."""
import threading
import time

import numpy as np
import pytest
import torch

import dgl_b200.graphbolt as gb
from dgl_b200.graphbolt.impl.features import CPUCachedFeature, DiskBasedFeature, TorchBasedFeature
from dgl_b200.graphbolt.internal import futures


def _run_stages(feature, ids):
    gen = feature.read_async(ids)
    fut = None
    n = feature.read_async_num_stages(ids.device)
    for _ in range(n):
        fut = next(gen)
    return fut.wait(), n


def test_future_launch_runs_off_thread_and_propagates_errors():
    main = threading.get_ident()
    f = futures.launch(lambda: threading.get_ident())
    assert f.wait() != main and f.done() and f.wait() != main  # wait() is idempotent

    def boom():
        raise ValueError("producer failed")

    with pytest.raises(ValueError, match="producer failed"):
        futures.launch(boom).wait()
    assert futures.then(futures.launch(lambda: 20), lambda v: v + 1).wait() == 21
    assert futures.completed(5).wait() == 5


def test_io_pool_overlaps_blocking_calls():
    futures.set_num_io_threads(4)
    t0 = time.perf_counter()
    fs = [futures.launch(time.sleep, 0.2) for _ in range(4)]
    for f in fs:
        f.wait()
    assert time.perf_counter() - t0 < 0.6
    with pytest.raises(ValueError):
        futures.set_num_io_threads(0)


def test_read_async_equals_read_for_every_feature_kind(tmp_path):
    torch.manual_seed(0)
    table = torch.randn(500, 12)
    ids = torch.randint(0, 500, (200,))
    mem = TorchBasedFeature(table)
    out, stages = _run_stages(mem, ids)
    assert stages == 1 and torch.equal(out, table[ids])
    path = str(tmp_path / "feat.npy")
    np.save(path, table.numpy())
    for backend in ("pread", "auto"):
        disk = DiskBasedFeature(path, io_backend=backend)
        out, stages = _run_stages(disk, ids)
        assert stages == 1 and torch.equal(out, table[ids])
        cached = CPUCachedFeature(disk, 100 * 12 * 4)
        assert cached.read_async_num_stages(ids.device) == 3
        for _ in range(3):  # cold, then partially warm
            out, _ = _run_stages(cached, ids)
            assert torch.equal(out, table[ids])
        assert cached.miss_rate < 1.0
        nested = CPUCachedFeature(cached, 50 * 12 * 4)
        assert nested.read_async_num_stages(ids.device) == 5
        out, _ = _run_stages(nested, ids)
        assert torch.equal(out, table[ids])


def _pipeline(table, etable, overlap, tmp_path):
    indptr = torch.tensor([0, 2, 4, 6, 8, 10, 12, 14, 16])
    indices = torch.tensor([1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 0, 0, 1])
    g = gb.fused_csc_sampling_graph(indptr, indices)
    path = str(tmp_path / f"n{int(overlap)}.npy")
    np.save(path, table.numpy())
    store = gb.BasicFeatureStore({("node", None, "x"): CPUCachedFeature(DiskBasedFeature(path), 4 * table.shape[1] * 4),
                                  ("edge", None, "w"): TorchBasedFeature(etable)})
    items = gb.ItemSet(torch.arange(8), names="seeds")
    dp = gb.ItemSampler(items, batch_size=2).sample_neighbor(g, [-1, -1])
    return dp.fetch_feature(store, node_feature_keys=["x"], edge_feature_keys=["w"], overlap_fetch=overlap)


def test_overlapped_fetcher_matches_synchronous_fetcher(tmp_path):
    torch.manual_seed(1)
    table, etable = torch.randn(8, 6), torch.randn(16, 3)
    got = list(_pipeline(table, etable, True, tmp_path))
    want = list(_pipeline(table, etable, False, tmp_path))
    assert len(got) == len(want) == 4
    for a, b in zip(got, want):
        assert torch.equal(a.input_nodes, b.input_nodes)
        assert torch.equal(a.node_features["x"], table[a.input_nodes]) and torch.equal(a.node_features["x"], b.node_features["x"])
        assert len(a.edge_features) == len(b.edge_features) == 2
        for ea, eb, sg in zip(a.edge_features, b.edge_features, a.sampled_subgraphs):
            assert torch.equal(ea["w"], eb["w"]) and torch.equal(ea["w"], etable[sg.original_edge_ids])


def test_sampling_async_op_returns_a_future_of_the_same_subgraph():
    indptr = torch.tensor([0, 2, 4, 6, 8])
    indices = torch.tensor([1, 2, 2, 3, 3, 0, 0, 1])
    g = gb.fused_csc_sampling_graph(indptr, indices)
    seeds = torch.tensor([0, 2, 3])
    fut = g.sample_neighbors(seeds, [-1], async_op=True)
    assert isinstance(fut, futures.Future)
    a, b = fut.wait(), g.sample_neighbors(seeds, [-1])
    assert torch.equal(a.sampled_csc.indptr, b.sampled_csc.indptr)
    assert torch.equal(a.sampled_csc.indices.sort()[0], b.sampled_csc.indices.sort()[0])
    sub = g.in_subgraph(seeds, async_op=True).wait()
    assert torch.equal(sub.sampled_csc.indptr, g.in_subgraph(seeds).sampled_csc.indptr)


@pytest.mark.gpu
def test_device_side_futures_match_synchronous_reads(tmp_path):
    from dgl_b200.graphbolt.impl.features import GPUCachedFeature
    torch.manual_seed(0)
    table = torch.randn(4000, 32)
    ids = torch.randint(0, 4000, (1500,), device="cuda")
    want = table.cuda()[ids]
    pageable = TorchBasedFeature(table.clone())
    out, stages = _run_stages(pageable, ids)
    assert stages == 2 and out.is_cuda and torch.equal(out, want)
    pinned = TorchBasedFeature(table.clone()).pin_memory_()
    out, stages = _run_stages(pinned, ids)
    assert stages == 1 and torch.equal(out, want)
    resident = TorchBasedFeature(table.cuda())
    out, _ = _run_stages(resident, ids)
    assert torch.equal(out, want)
    path = str(tmp_path / "f.npy")
    np.save(path, table.numpy())
    disk = DiskBasedFeature(path)
    out, stages = _run_stages(disk, ids)
    assert stages == 2 and out.is_cuda and torch.equal(out, want)
    for fallback in (pinned, disk):
        cached = GPUCachedFeature(fallback, 1024 * 32 * 4)
        for _ in range(3):
            out, _ = _run_stages(cached, ids)
            torch.cuda.synchronize()
            assert torch.equal(out, want)
        assert cached.miss_rate < 1.0
    # consumer on another stream: wait() must order it after the producer
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out, _ = _run_stages(pinned, ids)
        doubled = out * 2
    s.synchronize()
    assert torch.equal(doubled, want * 2)


@pytest.mark.gpu
def test_sampling_future_on_the_device():
    g = gb.fused_csc_sampling_graph(torch.tensor([0, 2, 4, 6, 8]), torch.tensor([1, 2, 2, 3, 3, 0, 0, 1])).to("cuda")
    seeds = torch.tensor([0, 2, 3], device="cuda")
    sub = g.sample_neighbors(seeds, [-1], async_op=True).wait()
    ref = g.sample_neighbors(seeds, [-1])
    assert torch.equal(sub.sampled_csc.indptr, ref.sampled_csc.indptr)
    assert torch.equal(sub.sampled_csc.indices.sort()[0], ref.sampled_csc.indices.sort()[0])
