"""GPU sampling kernels: ELL neighbour sampler invariants (uniform, without replacement)."""
import pytest
import torch

import dgl_b200 as dgl

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _csc(n, m, seed=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (m,), generator=g)
    dst = (torch.rand(m, generator=g) ** 3 * n).long().clamp(max=n - 1)  # skewed in-degrees incl. hubs and isolated nodes
    order = torch.argsort(dst, stable=True)
    src, dst = src[order], dst[order]
    indptr = torch.zeros(n + 1, dtype=torch.int64)
    indptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n), 0)
    return indptr, src


@pytest.mark.parametrize("idtype", [torch.int32, torch.int64])
@pytest.mark.parametrize("fanout", [3, 8, 10, 16, 25, 32, 40])
def test_sample_ell_invariants(idtype, fanout):
    from dgl_b200.ops.fused import sample_ell
    n, m = 3000, 60000
    indptr, indices = _csc(n, m, seed=fanout)
    ip, ix = indptr.to(DEV, idtype), indices.to(DEV, idtype)
    seeds = torch.randint(0, n, (2000,), generator=torch.Generator().manual_seed(1)).to(DEV, idtype)
    seeds[5] = -1  # empty slot
    n_dev = torch.tensor([1900], device=DEV, dtype=torch.int64)
    nbr, cnt, eid, total = sample_ell(ip, ix, seeds, fanout, seed=42, n_dev=n_dev, return_eids=True)
    torch.cuda.synchronize()
    nbr, cnt, eid, seeds_c = nbr.cpu().long(), cnt.cpu().long(), eid.cpu().long(), seeds.cpu().long()
    assert int(total.item()) == int(cnt[:1900].sum())
    assert (cnt[1900:] == 0).all() and (nbr[1900:] == -1).all()
    deg = indptr[1:] - indptr[:-1]
    for i in range(1900):
        v = int(seeds_c[i])
        if v < 0:
            assert cnt[i] == 0 and (nbr[i] == -1).all()
            continue
        d = int(deg[v])
        c = int(cnt[i])
        assert c == min(d, fanout)
        e = eid[i, :c]
        assert len(set(e.tolist())) == c  # without replacement
        assert ((e >= indptr[v]) & (e < indptr[v + 1])).all()
        assert torch.equal(indices[e], nbr[i, :c])
        assert (nbr[i, c:] == -1).all()


def test_sample_ell_uniform():
    """Each in-edge of a hub must be picked with probability fanout/deg."""
    from dgl_b200.ops.fused import sample_ell
    deg, fanout, reps = 50, 10, 20000
    indptr = torch.tensor([0, deg], dtype=torch.int32, device=DEV)
    indices = torch.arange(deg, dtype=torch.int32, device=DEV)
    seeds = torch.zeros(reps, dtype=torch.int32, device=DEV)
    nbr, cnt, _, _ = sample_ell(indptr, indices, seeds, fanout, seed=7)
    hist = torch.bincount(nbr.flatten().long(), minlength=deg).float() / reps
    assert (cnt == fanout).all()
    assert (hist - fanout / deg).abs().max() < 0.02, hist
    # a different step counter gives a different draw
    step = torch.tensor([3], dtype=torch.int64, device=DEV)
    nbr2, _, _, _ = sample_ell(indptr, indices, seeds, fanout, seed=7, step_dev=step)
    assert not torch.equal(nbr, nbr2)


@pytest.mark.parametrize("first_cap,nf,fan,universe", [(1000, 900, 10, 5000), (64, 64, 3, 50000), (5000, 4321, 7, 30000)])
def test_frontier_compact(first_cap, nf, fan, universe):
    from dgl_b200.ops.fused import FrontierCompactor
    g = torch.Generator().manual_seed(3)
    first = torch.randperm(universe, generator=g)[:first_cap].to(DEV, torch.int32)
    nbr = torch.randint(0, universe, (first_cap, fan), generator=g).to(DEV, torch.int32)
    nbr[torch.rand(first_cap, fan, generator=g).to(DEV) < 0.2] = -1  # padding inside rows
    n_dev = torch.tensor([nf], device=DEV, dtype=torch.int64)
    out_cap = first_cap * (1 + fan)
    fc = FrontierCompactor(first_cap, fan, out_cap, DEV)
    for rep in range(3):  # the table must come back clean: repeated calls give consistent results
        nodes, loc, count, ovf = fc.compact(first, n_dev, nbr)
        torch.cuda.synchronize()
        c = int(count.item())
        live = nbr[:nf]
        expect = torch.unique(torch.cat([first[:nf].long(), live[live >= 0].long()]))
        assert c == expect.numel() and int(ovf.item()) == 0
        assert torch.equal(nodes[:nf], first[:nf])                       # dst nodes are a prefix
        assert torch.equal(torch.sort(nodes[:c].long()).values, expect)   # exactly the unique set
        assert (nodes[c:] == -1).all()
        valid = live >= 0
        assert torch.equal(nodes[loc[:nf][valid].long()], live[valid])   # relabel is consistent
        assert (loc[:nf][~valid] == -1).all() and (loc[nf:] == -1).all()
        assert (fc.keys == -1).all() and (fc.vals == -2).all() and int(fc.counter.item()) == 0
    # capacity overflow is flagged and clamps the count
    fc2 = FrontierCompactor(first_cap, fan, nf + 5, DEV)
    nodes, loc, count, ovf = fc2.compact(first, n_dev, nbr)
    assert int(count.item()) == nf + 5 and int(ovf.item()) == 1


@pytest.mark.parametrize("fanout", [-1, 200, 300])
def test_weighted_sampling_above_128_picks(fanout):
    """Weighted sampling without replacement on a hub with > 128 positive-probability in-edges:
    every output slot must be a real, distinct, positive-probability in-edge - no uninitialised ids."""
    import dgl_b200 as dgl
    torch.manual_seed(0)
    n, hub_deg = 2000, 700
    src = torch.cat([torch.randint(1, n, (hub_deg,)), torch.randint(0, n, (5000,))])
    dst = torch.cat([torch.zeros(hub_deg, dtype=torch.int64), torch.randint(1, n, (5000,))])
    g = dgl.graph((src, dst), num_nodes=n)
    p = torch.rand(g.num_edges())
    p[torch.rand(g.num_edges()) < 0.3] = 0.0          # a third of the edges can never be picked
    g.edata["p"] = p
    gg = g.to(DEV)
    seeds = torch.tensor([0, 5, 17], device=DEV)
    sg = dgl.sampling.sample_neighbors(gg, seeds, fanout, prob="p")
    u, v = sg.edges()
    eid = sg.edata[dgl.EID]
    su, sv = gg.edges()
    assert torch.equal(su[eid], u) and torch.equal(sv[eid], v)          # edge ids point at real edges
    assert torch.unique(eid).numel() == eid.numel()                      # without replacement
    assert bool((gg.edata["p"][eid] > 0).all())
    nz_hub = int(((dst == 0) & (p > 0)).sum())
    want = nz_hub if fanout < 0 else min(fanout, nz_hub)
    assert int((v == 0).sum()) == want
    if fanout < 0:                                                       # everything eligible is returned
        cpu = dgl.sampling.sample_neighbors(g, seeds.cpu(), -1, prob="p")
        assert sorted(cpu.edata[dgl.EID].tolist()) == sorted(eid.cpu().tolist())


@pytest.mark.parametrize("idtype", [torch.int32, torch.int64])
def test_fused_block_sampler_matches_block_contract(idtype):
    """One-sync multi-layer sampler: blocks obey the MFG contract (dst prefix, NID/EID maps, real in-edges, fan-out)."""
    import dgl_b200 as dgl
    from dgl_b200.dataloading import NeighborSampler
    torch.manual_seed(0)
    n, m = 5000, 120000
    u, v = torch.randint(0, n, (m,)), torch.randint(0, n, (m,))
    g = dgl.graph((u, v), num_nodes=n, idtype=idtype).to(DEV)
    sampler = NeighborSampler([7, 5, 3])
    seeds = torch.randperm(n)[:300].to(idtype).to(DEV)
    for _ in range(2):                                   # second call reuses the cached workspace
        inp, out, blocks = sampler.sample_blocks(g, seeds)
        assert sampler._fast is not None and torch.equal(out, seeds)
        gu, gv = g.edges()
        deg = g.in_degrees()
        cur = seeds
        for blk, fan in zip(reversed(blocks), [3, 5, 7]):
            nd, ns = blk.num_dst_nodes(), blk.num_src_nodes()
            src_ids, dst_ids = blk.srcdata[dgl.NID], blk.dstdata[dgl.NID]
            assert torch.equal(dst_ids, cur) and torch.equal(src_ids[:nd], dst_ids)
            assert torch.unique(src_ids).numel() == ns
            bu, bv = blk.edges()
            eid = blk.edata[dgl.EID]
            assert torch.equal(gu[eid.long()], src_ids[bu.long()]) and torch.equal(gv[eid.long()], dst_ids[bv.long()])
            assert torch.unique(eid).numel() == eid.numel()                     # without replacement
            assert torch.equal(blk.in_degrees().long(), deg[dst_ids.long()].clamp(max=fan).long())
            nbr, cnt, f = blk._ell_cache
            assert f == fan and torch.equal(cnt.long(), blk.in_degrees().long())
            cur = src_ids
        assert torch.equal(inp, cur)
    # the fused SAGE layer consumes the attached ELL form; result equals the g-SpMM formulation
    from dgl_b200.nn import fused_sage
    conv = dgl.nn.SAGEConv(16, 8, "mean").to(DEV)
    x = torch.randn(blocks[0].num_src_nodes(), 16, device=DEV)
    y_fused = conv(blocks[0], x)
    fused_sage.ENABLED[0] = False
    try:
        y_ref = conv(blocks[0], x)
    finally:
        fused_sage.ENABLED[0] = True
    assert torch.allclose(y_fused, y_ref, atol=1e-4)


# ----------------------------------------------------------------------------------------------- LABOR-i on the device
def _labor_oracle(indptr, indices, eids, seeds, fanout, prob, iters):
    """c, pi by the defining equations, plain Python / fp64 on the host."""
    indptr, indices = indptr.cpu().tolist(), indices.cpu().tolist()
    eids = None if eids is None else eids.cpu().tolist()
    a = (lambda q: 1.0) if prob is None else (lambda q: float(prob[eids[q] if eids is not None else q]))
    pi = {}
    rows = [list(range(indptr[s], indptr[s + 1])) for s in seeds.cpu().tolist()]
    for r in rows:
        for q in r:
            pi[indices[q]] = 1.0

    def solve():
        cs = []
        for r in rows:
            w = [pi[indices[q]] * a(q) for q in r]
            live = sum(1 for x in w if x > 0)
            if live <= fanout:
                cs.append(float("inf"))
                continue
            lo, hi = 0.0, 1.0
            while sum(min(1.0, hi * x) for x in w) < fanout:
                hi *= 2
            for _ in range(100):
                mid = (lo + hi) / 2
                if sum(min(1.0, mid * x) for x in w) >= fanout:
                    hi = mid
                else:
                    lo = mid
            cs.append(hi)
        return cs

    c = solve()
    for _ in range(iters):
        nxt = {t: 0.0 for t in pi}
        for r, cs in zip(rows, c):
            for q in r:
                t = indices[q]
                nxt[t] = max(nxt[t], min(cs, 1e30) * pi[t] * a(q))
        pi = {t: max(v, 1e-12) for t, v in nxt.items()}
        c = solve()
    return c, pi, rows


@pytest.mark.parametrize("iters", [0, 1, 3])
@pytest.mark.parametrize("weighted", [False, True])
def test_labor_importance_kernels_match_oracle(iters, weighted, idtype):
    from dgl_b200._graph_kernels import labor_sample
    torch.manual_seed(0)
    n = 300
    g = dgl.rand_graph(n, 6000, idtype=idtype)
    indptr, indices, eids = g._graphs[0].csc()
    prob = None
    if weighted:
        prob = torch.rand(g.num_edges()) + 0.05
        prob[::7] = 0.0
    seeds = torch.randperm(n)[:64].to(idtype)
    fanout = 5
    d = lambda t: None if t is None else t.to(DEV)
    src, pos, eid, w, (c, touched, pi) = labor_sample(d(indptr), d(indices), d(eids), d(seeds), fanout, d(prob), iters, 11, 12, 0.0,
                                                      num_src=n, return_state=True)
    oc, opi, rows = _labor_oracle(indptr, indices, eids, seeds, fanout, prob, iters)
    c = c.cpu()
    for i, (a, b) in enumerate(zip(c.tolist(), oc)):
        assert (a > 1e37) == (b == float("inf")) and (b == float("inf") or abs(a - b) <= 2e-3 * b), (i, a, b)
    got_pi = dict(zip(touched.cpu().tolist(), pi.cpu().tolist()))
    assert set(got_pi) == set(opi)
    for t, v in opi.items():
        assert abs(got_pi[t] - v) <= 2e-3 * max(v, 1e-6), (t, got_pi[t], v)
    # sampled edges are in-edges of their seed, weights are Hajek-normalised, zero-probability edges never appear
    ind_l, eid_l = indices.tolist(), (None if eids is None else eids.tolist())
    pos, src, eid, w = pos.cpu().long(), src.cpu(), eid.cpu(), w.cpu()
    for i, r in enumerate(rows):
        sel = (pos == i).nonzero().squeeze(1)
        allowed = {(eid_l[q] if eid_l is not None else q): ind_l[q] for q in r}
        for j in sel.tolist():
            assert int(eid[j]) in allowed and allowed[int(eid[j])] == int(src[j])
            if prob is not None:
                assert prob[int(eid[j])] > 0
        if sel.numel():
            assert abs(float(w[sel].sum()) - 1.0) < 1e-4
    # a second call sees neutral cached importances again (same result for the same seeds)
    again = labor_sample(d(indptr), d(indices), d(eids), d(seeds), fanout, d(prob), iters, 11, 12, 0.0, num_src=n)
    assert torch.equal(again[0].cpu(), src) and torch.allclose(again[3].cpu(), w)


def test_labor_shared_variates_expected_size_and_continuous_seed():
    from dgl_b200._graph_kernels import labor_sample
    torch.manual_seed(1)
    n = 2000
    g = dgl.rand_graph(n, 80000, idtype=torch.int32).to(DEV)
    indptr, indices, eids = g._graphs[0].csc()
    seeds = torch.arange(512, dtype=torch.int32, device=DEV)
    fanout = 10
    cache = {}
    sizes, uniq = [], {}
    for it in (0, 2):
        tot, un = 0, 0
        for s in range(20):
            src, pos, eid, w = labor_sample(indptr, indices, eids, seeds, fanout, None, it, 100 + s, 0, 0.0, num_src=n, cache=cache)
            tot += src.numel()
            un += torch.unique(src).numel()
        sizes.append(tot / 20 / 512)
        uniq[it] = un / 20
    assert all(abs(s - fanout) < 0.5 for s in sizes), sizes        # E[#edges per seed] = fanout
    assert uniq[2] < uniq[0]                                         # importance iterations shrink the union of sources
    # one variate per source vertex: a source is kept by ALL or NONE of the seeds whose threshold it passes
    src, pos, eid, w = labor_sample(indptr, indices, eids, seeds, fanout, None, 0, 5, 0, 0.0, num_src=n, cache=cache)
    deg = (indptr[1:] - indptr[:-1])[seeds.long()].float()
    thr = (fanout / deg).clamp(max=1.0)
    kept_thr_min = torch.full((n,), 2.0, device=DEV).scatter_reduce(0, src.long(), thr[pos.long()], "amin")
    # every (seed, t) in-edge with threshold >= the smallest threshold at which t was kept must be kept too
    from dgl_b200._graph_kernels import _expand_rows
    q = _expand_rows(indptr, seeds)
    owner = torch.repeat_interleave(torch.arange(512, device=DEV), (indptr[1:] - indptr[:-1])[seeds.long()].long())
    t_all = indices[q].long()
    must = thr[owner] >= kept_thr_min[t_all]
    kept_pairs = set(zip(pos.tolist(), src.tolist()))
    all_pairs = list(zip(owner[must].tolist(), t_all[must].tolist()))
    assert all(p in kept_pairs for p in all_pairs)
    # continuous seed: mix = 0 / 1 reproduce the two end-point streams, in between the overlap decays smoothly
    def keys(mix, s1=7, s2=8):
        r = labor_sample(indptr, indices, eids, seeds, fanout, None, 0, s1, s2, mix, num_src=n, cache=cache)
        return set(zip(r[1].tolist(), r[0].tolist()))
    a, b = keys(0.0), keys(1.0)
    assert a == keys(0.0, 7, 999) and b == keys(0.0, 8, 0)
    jac = [len(a & keys(m)) / len(a | keys(m)) for m in (0.1, 0.5, 0.9)]
    assert jac[0] > jac[1] > jac[2] > len(a & b) / len(a | b) - 0.02


def test_sample_labors_uses_device_kernels_for_importance_sampling():
    g = dgl.rand_graph(500, 20000).to(DEV)
    g.edata["p"] = torch.rand(g.num_edges(), device=DEV)
    sg, ws = dgl.sampling.sample_labors(g, torch.arange(100, device=DEV), 5, prob="p", importance_sampling=2, random_seed=3)
    assert sg.num_edges() == ws[0].numel() > 0
    u, v = sg.edges()
    assert set(v.tolist()) <= set(range(100)) and g.has_edges_between(u, v).all()
    sg2, ws2 = dgl.sampling.sample_labors(g, torch.arange(100, device=DEV), 5, prob="p", importance_sampling=2, random_seed=3)
    assert torch.equal(sg.edata[dgl.EID], sg2.edata[dgl.EID])
