"""Edge-wise sampling adaptor and prefetch markers (parity: /root/reference/tests/python/common/dataloading/
test_dataloader.py edge-dataloader cases; checks here are value-level)."""
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200.dataloading import (DataLoader, MultiLayerFullNeighborSampler, NeighborSampler, as_edge_prediction_sampler,
                                  negative_sampler)
from dgl_b200.dataloading.base import _find_exclude_eids


def bidirected(n=40, m=120, seed=0):
    gen = torch.Generator().manual_seed(seed)
    u = torch.randint(0, n, (m,), generator=gen)
    v = torch.randint(0, n, (m,), generator=gen)
    g = dgl.graph((torch.cat([u, v]), torch.cat([v, u])), num_nodes=n)
    rev = torch.cat([torch.arange(m, 2 * m), torch.arange(0, m)])
    return g, rev, m


def block_parent_eids(block):
    return set(block.edata[dgl.EID].tolist())


def test_pair_graph_matches_seed_edges_and_reverse_exclusion():
    g, rev, m = bidirected()
    g.edata["label"] = torch.arange(g.num_edges()).float()
    sampler = as_edge_prediction_sampler(MultiLayerFullNeighborSampler(2), exclude="reverse_id", reverse_eids=rev,
                                         prefetch_labels=["label"])
    seeds = torch.tensor([3, 17, 50])
    input_nodes, pair, blocks = sampler.sample(g, seeds)
    # the pair graph holds exactly the seed edges, expressed in the compacted id space
    assert pair.edata[dgl.EID].tolist() == seeds.tolist()
    pu, pv = pair.edges()
    gu, gv = g.find_edges(seeds)
    assert pair.ndata[dgl.NID][pu].tolist() == gu.tolist() and pair.ndata[dgl.NID][pv].tolist() == gv.tolist()
    # the seed nodes of the neighbourhood sampling are the pair graph's nodes
    assert blocks[-1].dstdata[dgl.NID].tolist() == pair.ndata[dgl.NID].tolist()
    assert blocks[0].srcdata[dgl.NID].tolist() == input_nodes.tolist()
    # neither the seeds nor their reverses survive in any message-flow graph
    banned = set(seeds.tolist()) | set(rev[seeds].tolist())
    for b in blocks:
        assert not (block_parent_eids(b) & banned)
    # the label column is a placeholder until the loader resolves it
    resolved = dgl.dataloading.assign_lazy_features((input_nodes, pair, blocks), g)
    assert resolved[1].edata["label"].tolist() == seeds.float().tolist()


def test_negative_graph_shares_the_compacted_node_space():
    g, rev, m = bidirected(seed=1)
    sampler = as_edge_prediction_sampler(NeighborSampler([3, 3]), negative_sampler=negative_sampler.Uniform(2))
    seeds = torch.arange(8)
    input_nodes, pos, neg, blocks = sampler.sample(g, seeds)
    assert neg.num_edges() == 2 * seeds.numel()
    assert pos.num_nodes() == neg.num_nodes()
    assert torch.equal(pos.ndata[dgl.NID], neg.ndata[dgl.NID])
    # negative sources repeat the positive sources (uniform corruption of the destination)
    gu, _ = g.find_edges(seeds)
    nu, _ = neg.edges()
    assert neg.ndata[dgl.NID][nu].tolist() == gu.repeat_interleave(2).tolist()
    assert blocks[-1].num_dst_nodes() == pos.num_nodes()


def test_exclusion_policies():
    g, rev, m = bidirected(seed=2)
    seeds = torch.tensor([1, 2])
    assert _find_exclude_eids(g, None, seeds) is None
    assert _find_exclude_eids(g, "self", seeds) is seeds
    assert _find_exclude_eids(g, "reverse_id", seeds, reverse_eids=rev).tolist() == [1, 2, m + 1, m + 2]
    assert _find_exclude_eids(g, lambda s: s + 1, seeds).tolist() == [2, 3]
    with pytest.raises(ValueError):
        _find_exclude_eids(g, "bogus", seeds)
    with pytest.raises(ValueError):
        as_edge_prediction_sampler(NeighborSampler([2]), exclude="reverse_id")
    with pytest.raises(ValueError):
        as_edge_prediction_sampler(NeighborSampler([2]), exclude="reverse_types")


def test_reverse_types_on_heterograph():
    u = torch.tensor([0, 1, 2, 3, 0, 2])
    i = torch.tensor([0, 1, 2, 0, 2, 1])
    hg = dgl.heterograph({("user", "buys", "item"): (u, i), ("item", "bought-by", "user"): (i, u)})
    pol = _find_exclude_eids(hg, "reverse_types", {"buys": torch.tensor([0, 4])},
                             reverse_etypes={"buys": "bought-by", "bought-by": "buys"})
    assert pol[("user", "buys", "item")].tolist() == [0, 4] and pol[("item", "bought-by", "user")].tolist() == [0, 4]
    sampler = as_edge_prediction_sampler(MultiLayerFullNeighborSampler(1), exclude="reverse_types",
                                         reverse_etypes={"buys": "bought-by", "bought-by": "buys"})
    input_nodes, pair, blocks = sampler.sample(hg, {"buys": torch.tensor([0, 4])})
    assert pair.num_edges("buys") == 2 and pair.num_edges("bought-by") == 0
    for rel in blocks[0].canonical_etypes:
        assert not (set(blocks[0].edges[rel].data[dgl.EID].tolist()) & {0, 4})


def test_edge_dataloader_end_to_end():
    g, rev, m = bidirected(seed=3)
    g.ndata["x"] = torch.randn(g.num_nodes(), 4)
    sampler = as_edge_prediction_sampler(NeighborSampler([4], prefetch_node_feats=["x"]), exclude="reverse_id",
                                         reverse_eids=rev, negative_sampler=negative_sampler.Uniform(1))
    seen = 0
    for input_nodes, pos, neg, blocks in DataLoader(g, torch.arange(m), sampler, batch_size=32, shuffle=False):
        assert torch.equal(blocks[0].srcdata["x"], g.ndata["x"][input_nodes])
        seen += pos.num_edges()
    assert seen == m
