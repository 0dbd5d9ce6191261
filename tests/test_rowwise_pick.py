"""Host row-wise pickers (csrc/cpu/rowwise_pick.cc): per-edge-type fan-outs, tag-biased sampling, top-k. Structure
checks are exact; distribution checks compare empirical frequencies against the closed form."""
import torch

import dgl_b200 as dgl
from dgl_b200.sampling import sample_etype_neighbors, sample_neighbors_biased, select_topk
from dgl_b200.transforms import sort_csc_by_tag, sort_csr_by_tag


def _two_rel(seed=0):
    torch.manual_seed(seed)
    hg = dgl.heterograph({("a", "r0", "a"): (torch.randint(0, 50, (400,)), torch.randint(0, 50, (400,))),
                          ("a", "r1", "a"): (torch.randint(0, 50, (300,)), torch.randint(0, 50, (300,)))})
    return dgl.to_homogeneous(hg), [0, 400, 700]


def test_etype_fanouts_exact_counts_and_edge_identity():
    g, off = _two_rel()
    u, v = g.edges()
    indeg = [torch.bincount(v[off[t]:off[t + 1]], minlength=50) for t in range(2)]
    fr = sample_etype_neighbors(g, torch.arange(50), off, [2, 3])
    eid = fr.edata[dgl.EID]
    assert eid.unique().numel() == fr.num_edges()
    assert torch.equal(u[eid], fr.edges()[0]) and torch.equal(v[eid], fr.edges()[1])
    et = (eid >= 400).long()
    for t, f in ((0, 2), (1, 3)):
        assert torch.equal(torch.bincount(fr.edges()[1][et == t], minlength=50), indeg[t].clamp(max=f))
    fr = sample_etype_neighbors(g, torch.arange(50), off, [2, 3], replace=True)
    assert fr.num_edges() == int(((indeg[0] > 0) * 2 + (indeg[1] > 0) * 3).sum())
    fr = sample_etype_neighbors(g, torch.arange(50), off, [-1, 0])
    assert fr.num_edges() == 400 and bool((fr.edata[dgl.EID] < 400).all())


def test_etype_probabilities_and_out_direction():
    g, off = _two_rel(1)
    g.edata["p"] = torch.zeros(700)
    g.edata["p"][::2] = 1.0
    fr = sample_etype_neighbors(g, torch.arange(50), off, [5, 5], prob="p")
    assert bool((fr.edata[dgl.EID] % 2 == 0).all())
    fr = sample_etype_neighbors(g, torch.arange(10), off, [1, 1], edge_dir="out")
    u, v = g.edges()
    assert set(fr.edges()[0].tolist()) <= set(range(10))
    assert torch.equal(u[fr.edata[dgl.EID]], fr.edges()[0])
    # weighted draw inside one type follows the weights
    g2 = dgl.graph((torch.tensor([1, 2, 3]), torch.tensor([0, 0, 0])), num_nodes=4)
    g2.edata["p"] = torch.tensor([1.0, 2.0, 7.0])
    hits = torch.zeros(3)
    for _ in range(3000):
        fr = sample_etype_neighbors(g2, torch.tensor([0]), [0, 3], [1], prob="p", replace=True)
        hits[fr.edata[dgl.EID][0]] += 1
    assert torch.allclose(hits / 3000, torch.tensor([0.1, 0.2, 0.7]), atol=0.04)


def test_biased_sampling_follows_tag_bias():
    torch.manual_seed(0)
    n = 40
    src = torch.arange(1, n)
    g = dgl.graph((src, torch.zeros(n - 1, dtype=torch.int64)), num_nodes=n)
    tag = torch.zeros(n, dtype=torch.int64)
    tag[1:11] = 0
    tag[11:31] = 1
    tag[31:] = 2
    gs = sort_csc_by_tag(g, tag)
    bias = [1.0, 0.0, 4.0]
    hits = torch.zeros(3)
    for _ in range(2000):
        fr = sample_neighbors_biased(gs, torch.tensor([0]), 2, bias, replace=True)
        for s in fr.edges()[0].tolist():
            hits[tag[s]] += 1
    # weights: tag0 = 1 * 10, tag2 = 4 * 9
    assert hits[1] == 0
    assert abs(float(hits[0] / hits.sum()) - 10 / 46) < 0.04
    fr = sample_neighbors_biased(gs, torch.tensor([0]), 100, bias)
    picked = sorted(fr.edges()[0].tolist())
    assert picked == list(range(1, 11)) + list(range(31, 40))  # every positive-bias neighbour exactly once
    fr = sample_neighbors_biased(gs, torch.tensor([0]), 5, bias)
    assert fr.num_edges() == 5 and len(set(fr.edges()[0].tolist())) == 5
    assert all(tag[s] != 1 for s in fr.edges()[0].tolist())
    go = sort_csr_by_tag(dgl.graph((torch.zeros(n - 1, dtype=torch.int64), src), num_nodes=n), tag)
    fr = sample_neighbors_biased(go, torch.tensor([0]), 3, bias, edge_dir="out")
    assert fr.num_edges() == 3 and set(fr.edges()[0].tolist()) == {0}


def test_topk_matches_sorted_oracle_with_stable_ties():
    torch.manual_seed(3)
    g = dgl.rand_graph(60, 900)
    for dt in (torch.float32, torch.float64, torch.int64):
        w = torch.randint(0, 6, (900,)).to(dt)  # many ties
        g.edata["w"] = w
        for asc in (False, True):
            for direction in ("in", "out"):
                fr = select_topk(g, 4, "w", edge_dir=direction, ascending=asc)
                u, v = g.edges()
                rows = v if direction == "in" else u
                got = {}
                for e, r in zip(fr.edata[dgl.EID].tolist(), (fr.edges()[1] if direction == "in" else fr.edges()[0]).tolist()):
                    got.setdefault(r, []).append(e)
                for r in range(60):
                    es = torch.nonzero(rows == r).squeeze(1)
                    if es.numel() == 0:
                        assert r not in got
                        continue
                    ws = w[es].double()
                    want = sorted(ws.tolist(), reverse=not asc)[:4]
                    have = sorted(w[torch.tensor(got[r])].double().tolist(), reverse=not asc)
                    assert have == want
    fr = select_topk(g, -1, "w")
    assert fr.num_edges() == 900
