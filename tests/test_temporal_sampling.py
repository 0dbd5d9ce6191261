"""Time-respecting sampling on FusedCSCSamplingGraph (parity: /root/reference/tests/python/pytorch/graphbolt/impl/
test_fused_csc_sampling_graph.py temporal cases; the checks are exhaustive on small random graphs)."""
import pytest
import torch

import dgl_b200.graphbolt as gb
from dgl_b200.graphbolt.impl.temporal import _hash_uniform, expand_candidates, pick_per_owner


def random_csc(n=30, e=300, seed=0):
    gen = torch.Generator().manual_seed(seed)
    dst = torch.randint(0, n, (e,), generator=gen).sort().values
    src = torch.randint(0, n, (e,), generator=gen)
    indptr = torch.zeros(n + 1, dtype=torch.int64)
    indptr[1:] = torch.bincount(dst, minlength=n).cumsum(0)
    return indptr, src, gen


def eligible(indptr, src, v, t, node_ts=None, edge_ts=None, window=None):
    out = []
    for p in range(int(indptr[v]), int(indptr[v + 1])):
        ok = True
        for st in ((node_ts[src[p]] if node_ts is not None else None), (edge_ts[p] if edge_ts is not None else None)):
            if st is None:
                continue
            ok &= bool(st < t) and (window is None or bool(st >= t - window))
        if ok:
            out.append(p)
    return out


@pytest.mark.parametrize("use_node,use_edge,use_window", [(True, False, False), (False, True, False), (True, True, True)])
@pytest.mark.parametrize("fanout", [-1, 2, 5])
def test_homogeneous_respects_time_and_fanout(use_node, use_edge, use_window, fanout):
    indptr, src, gen = random_csc()
    n, e = indptr.numel() - 1, src.numel()
    node_ts = torch.randint(0, 100, (n,), generator=gen)
    edge_ts = torch.randint(0, 100, (e,), generator=gen)
    g = gb.fused_csc_sampling_graph(indptr, src, node_attributes={"nt": node_ts}, edge_attributes={"et": edge_ts})
    seeds = torch.tensor([3, 7, 7, 11, 29])                    # a duplicate seed with two different timestamps
    ts = torch.tensor([50, 20, 90, 0, 70])
    win = torch.tensor([40, 15, 30, 5, 60]) if use_window else None
    sub = g.temporal_sample_neighbors(seeds, ts, [fanout], seeds_pre_time_window=win,
                                      node_timestamp_attr_name="nt" if use_node else None,
                                      edge_timestamp_attr_name="et" if use_edge else None)
    ptr, idx, eid = sub.sampled_csc.indptr, sub.sampled_csc.indices, sub.original_edge_ids
    assert ptr.numel() == seeds.numel() + 1 and ptr[-1] == idx.numel() == eid.numel()
    for i, (v, t) in enumerate(zip(seeds.tolist(), ts.tolist())):
        allowed = eligible(indptr, src, v, t, node_ts if use_node else None, edge_ts if use_edge else None,
                           None if win is None else int(win[i]))
        got = eid[ptr[i]:ptr[i + 1]].tolist()
        assert set(got) <= set(allowed) and len(set(got)) == len(got)
        assert len(got) == (len(allowed) if fanout < 0 else min(fanout, len(allowed)))
        assert idx[ptr[i]:ptr[i + 1]].tolist() == src[got].tolist()


def test_with_replacement_and_weights():
    indptr, src, gen = random_csc(seed=1)
    e = src.numel()
    edge_ts = torch.randint(0, 100, (e,), generator=gen)
    w = torch.rand(e, generator=gen)
    w[::3] = 0.0
    g = gb.fused_csc_sampling_graph(indptr, src, edge_attributes={"et": edge_ts, "w": w})
    seeds = torch.arange(10)
    ts = torch.full((10,), 60)
    sub = g.temporal_sample_neighbors(seeds, ts, [4], replace=True, probs_name="w", edge_timestamp_attr_name="et")
    ptr, eid = sub.sampled_csc.indptr, sub.original_edge_ids
    for i in range(10):
        allowed = [p for p in eligible(indptr, src, i, 60, None, edge_ts) if w[p] > 0]
        got = eid[ptr[i]:ptr[i + 1]].tolist()
        assert len(got) == (4 if allowed else 0) and set(got) <= set(allowed)
    # without replacement zero-weight edges are never picked either
    sub = g.temporal_sample_neighbors(seeds, ts, [50], probs_name="w", edge_timestamp_attr_name="et")
    assert (w[sub.original_edge_ids] > 0).all()


def test_weighted_selection_frequencies():
    torch.manual_seed(0)
    owner = torch.zeros(4, dtype=torch.int64)
    pos = torch.arange(4)
    w = torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64)
    hits = torch.zeros(4)
    for _ in range(4000):
        sel, cnt = pick_per_owner(owner, pos, torch.ones(4, dtype=torch.bool), 1, 1, False, w)
        hits[sel] += 1
    assert torch.allclose(hits / hits.sum(), w.float() / 10, atol=0.03)


def hetero_graph(seed=2):
    gen = torch.Generator().manual_seed(seed)
    n_user, n_item = 8, 6                      # homogenised ids: users 0..7, items 8..13
    indptr, src, _ = random_csc(n_user + n_item, 160, seed)
    dst = torch.repeat_interleave(torch.arange(n_user + n_item), indptr[1:] - indptr[:-1])
    ntype = lambda x: (x >= n_user).long()
    # edge types: 0 user->user, 1 user->item, 2 item->user, 3 item->item
    etype = ntype(src) * 2 + ntype(dst)
    names = {"user:follows:user": 0, "user:buys:item": 1, "item:sold-to:user": 2, "item:similar:item": 3}
    return indptr, src, dst, etype, names, n_user, n_item, gen


def test_heterogeneous_types_times_and_per_type_fanouts():
    indptr, src, dst, etype, names, n_user, n_item, gen = hetero_graph()
    node_ts = torch.randint(0, 50, (n_user + n_item,), generator=gen)
    g = gb.fused_csc_sampling_graph(indptr, src, node_type_offset=torch.tensor([0, n_user, n_user + n_item]),
                                    type_per_edge=etype, node_type_to_id={"user": 0, "item": 1}, edge_type_to_id=names,
                                    node_attributes={"nt": node_ts})
    seeds = {"user": torch.tensor([1, 4]), "item": torch.tensor([0, 5, 5])}
    ts = {"user": torch.tensor([30, 45]), "item": torch.tensor([10, 49, 25])}
    fan = [2, 0, 3, -1]
    sub = g.temporal_sample_neighbors(seeds, ts, fan, node_timestamp_attr_name="nt")
    assert "user:buys:item" not in sub.sampled_csc                      # fan-out 0 relations are skipped
    for et, tid in names.items():
        if fan[tid] == 0:
            continue
        s_t, _, d_t = et.split(":")
        csc, eid = sub.sampled_csc[et], sub.original_edge_ids[et]
        off_s, off_d = (0 if s_t == "user" else n_user), (0 if d_t == "user" else n_user)
        assert csc.indptr.numel() == seeds[d_t].numel() + 1
        for i, (v, t) in enumerate(zip(seeds[d_t].tolist(), ts[d_t].tolist())):
            allowed = [p for p in eligible(indptr, src, v + off_d, t, node_ts) if etype[p] == tid]
            got = eid[csc.indptr[i]:csc.indptr[i + 1]].tolist()
            assert set(got) <= set(allowed) and len(got) == (len(allowed) if fan[tid] < 0 else min(fan[tid], len(allowed)))
            assert (csc.indices[csc.indptr[i]:csc.indptr[i + 1]] + off_s).tolist() == src[got].tolist()
    with pytest.raises(Exception):
        g.temporal_sample_neighbors(torch.tensor([1]), torch.tensor([3]), [2], node_timestamp_attr_name="nt")


def test_layer_variant_shares_variates_across_seeds():
    # two seeds with identical in-neighbourhoods and timestamps must pick identical sources in the layer variant
    src = torch.arange(2, 22).repeat(2)
    indptr = torch.tensor([0, 20, 40] + [40] * 20)
    edge_ts = torch.zeros(40, dtype=torch.int64)
    g = gb.fused_csc_sampling_graph(indptr, src, edge_attributes={"et": edge_ts})
    seeds, ts = torch.tensor([0, 1]), torch.tensor([5, 5])
    sub = g.temporal_sample_layer_neighbors(seeds, ts, [6], edge_timestamp_attr_name="et", random_seed=123)
    ptr, idx = sub.sampled_csc.indptr, sub.sampled_csc.indices
    assert sorted(idx[ptr[0]:ptr[1]].tolist()) == sorted(idx[ptr[1]:ptr[2]].tolist()) and ptr[1] == 6
    again = g.temporal_sample_layer_neighbors(seeds, ts, [6], edge_timestamp_attr_name="et", random_seed=123)
    assert torch.equal(again.sampled_csc.indices, idx)
    other = g.temporal_sample_layer_neighbors(seeds, ts, [6], edge_timestamp_attr_name="et", random_seed=124)
    assert not torch.equal(other.sampled_csc.indices.sort().values, idx.sort().values)
    u = _hash_uniform(torch.arange(20000), 7)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean().item() - 0.5) < 0.01
