"""GPU sampling kernels: ELL neighbour sampler invariants (uniform, without replacement)."""
import pytest
import torch

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
