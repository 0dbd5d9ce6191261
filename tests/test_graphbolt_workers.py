"""GraphBolt ``DataLoader(num_workers > 0)``: host-only stages in front of the feature fetcher run in worker processes
(``MultiprocessingWrapper`` + datapipe rewriting) and the stream of minibatches equals the single-process one."""
import pytest
import torch

import dgl_b200 as dgl
import dgl_b200.graphbolt as gb
from dgl_b200.graphbolt.dataloader import MultiprocessingWrapper


def pipeline(num_workers, fetch=True, shuffle=True):
    torch.manual_seed(0)
    g = dgl.rand_graph(2000, 20000)
    graph = gb.from_dglgraph(g)
    feat = torch.arange(2000, dtype=torch.float32).unsqueeze(1).repeat(1, 4)
    items = gb.ItemSet((torch.arange(1000), torch.arange(1000) % 3), names=("seeds", "labels"))
    dp = gb.ItemSampler(items, batch_size=64, shuffle=shuffle, seed=5).sample_neighbor(graph, [4, 4])
    if fetch:
        dp = dp.fetch_feature(gb.BasicFeatureStore({("node", None, "feat"): gb.TorchBasedFeature(feat)}), ["feat"])
    return gb.DataLoader(dp, num_workers=num_workers), feat


def test_worker_processes_match_single_process_order_over_epochs():
    single, feat = pipeline(0)
    multi, _ = pipeline(3)
    for epoch in range(3):
        want = [(mb.seeds, mb.labels) for mb in single]
        got = list(multi)
        assert len(got) == len(want) == 16
        for (s, lab), mb in zip(want, got):
            assert torch.equal(mb.seeds, s) and torch.equal(mb.labels, lab)
            assert torch.equal(mb.node_features["feat"][:, 0], mb.input_nodes.float())     # fetched in the main process
            assert len(mb.blocks) == 2 and mb.blocks[-1].num_dst_nodes() == len(s)
    # the rewrite put the processes in front of the feature fetcher, and they persist between epochs
    assert type(multi.datapipe).__name__ == "FeatureFetcher" and isinstance(multi.datapipe.datapipe, MultiprocessingWrapper)
    wrapper = multi.datapipe.datapipe
    assert len(wrapper._procs) == 3 and all(p.is_alive() for p in wrapper._procs)
    first_pids = [p.pid for p in wrapper._procs]
    list(multi)
    assert [p.pid for p in wrapper._procs] == first_pids
    wrapper.shutdown()
    assert wrapper._procs == []


def test_without_fetcher_the_whole_chain_moves_and_early_exit_restarts_cleanly():
    multi, _ = pipeline(2, fetch=False, shuffle=False)
    it = iter(multi)
    first = next(it)
    assert first.seeds.tolist() == list(range(64))
    it.close()                                                   # abandoned mid-epoch: the pool is torn down
    assert isinstance(multi.datapipe, MultiprocessingWrapper) and multi.datapipe._procs == []
    seeds = torch.cat([mb.seeds for mb in multi])                # next epoch starts fresh and complete
    assert torch.equal(seeds, torch.arange(1000))
    multi.datapipe.shutdown()


def test_worker_errors_surface_in_the_parent():
    def boom(mb):
        if int(mb.seeds[0]) >= 128:
            raise ValueError("bad batch")
        return mb
    items = gb.ItemSet(torch.arange(256), names="seeds")
    dl = gb.DataLoader(gb.ItemSampler(items, batch_size=64).transform(boom), num_workers=2)
    with pytest.raises(RuntimeError, match="bad batch"):
        list(dl)


def test_item_sampler_worker_sharding():
    items = gb.ItemSet(torch.arange(100), names="seeds")
    s = gb.ItemSampler(items, batch_size=10)
    s.shard_for_worker(1, 3)
    assert [int(mb.seeds[0]) for mb in s] == [10, 40, 70]
    s.shard_for_worker(0, 1)
    assert len(list(s)) == 10


@pytest.mark.gpu
def test_gpu_pipelines_keep_the_producer_thread():
    g = dgl.rand_graph(500, 4000)
    graph = gb.from_dglgraph(g)
    items = gb.ItemSet(torch.arange(200), names="seeds")
    dp = gb.ItemSampler(items, batch_size=50).sample_neighbor(graph, [3]).copy_to("cuda")
    dl = gb.DataLoader(dp, num_workers=2)
    out = list(dl)
    assert len(out) == 4 and out[0].seeds.is_cuda and not isinstance(dl.datapipe, MultiprocessingWrapper)
    assert not getattr(dl, "_rewritten", False)                 # no processes were inserted in front of CUDA stages
