"""Index datasets, legacy collators and parent-storage stripping (reference dataloading/dataloader.py:166-620,
tests/python/pytorch/dataloading/test_dataloader.py::test_ddp_tensorized_dataset / test_node_dataloader)."""
import os
import pickle

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dgl_b200 as dgl
from dgl_b200 import dataloading as DL


def test_tensorized_dataset_batches_and_shuffle():
    ds = DL.TensorizedDataset(torch.arange(10), 4, False, True)
    assert len(ds) == 3
    first = torch.cat(list(ds))
    assert first.tolist() == list(range(10))                    # no shuffle() call yet: identity order
    torch.manual_seed(0)
    ds.shuffle()
    batches = list(ds)
    assert [b.numel() for b in batches] == [4, 4, 2] and sorted(torch.cat(batches).tolist()) == list(range(10))
    assert torch.cat(batches).tolist() != list(range(10))
    dropped = DL.TensorizedDataset(torch.arange(10), 4, True, False)
    assert len(dropped) == 2 and [b.tolist() for b in dropped] == [[0, 1, 2, 3], [4, 5, 6, 7]]
    mask = torch.tensor([True, False, True, True])
    assert torch.cat(list(DL.TensorizedDataset(mask, 8, False, False))).tolist() == [0, 2, 3]
    d = DL.TensorizedDataset({"a": torch.arange(3), "b": torch.arange(5)}, 4, False, False)
    out = list(d)
    assert out[0]["a"].tolist() == [0, 1, 2] and out[0]["b"].tolist() == [0] and out[1]["b"].tolist() == [1, 2, 3, 4]
    shared = DL.create_tensorized_dataset(torch.arange(6), 2, False, False, 0, True, True)
    assert isinstance(shared, DL.TensorizedDataset) and shared._perm.is_shared()


def _ddp_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {}
    for drop_last in (False, True):
        ds = DL.create_tensorized_dataset(torch.arange(11) * 10, 2, drop_last, True, 7, True, False)
        epochs = []
        for ep in range(2):
            ds.set_epoch(ep)
            ds.shuffle()
            epochs.append(torch.cat(list(ds)).tolist())
        res[drop_last] = (epochs, len(ds))
    torch.save(res, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_ddp_tensorized_dataset_shards(tmp_path):
    port = 29500 + os.getpid() % 500
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(2))
    # no drop_last: ceil(11/2)=6 per rank, the union covers everything with one wrap-around duplicate
    (e0, n0), (e1, _) = r0[False], r1[False]
    assert n0 == 3
    for ep in range(2):
        assert len(e0[ep]) == 6 and len(e1[ep]) == 6
        assert set(e0[ep]) | set(e1[ep]) == set(range(0, 110, 10))
    assert e0[0] != e0[1]                                             # reshuffled per epoch
    # drop_last: floor(11/2)=5 per rank, disjoint
    (d0, n), (d1, _) = r0[True], r1[True]
    assert n == 2 and len(d0[0]) == 4 and not set(d0[0]) & set(d1[0])   # 5 samples, batch 2, last short batch dropped


def test_legacy_collators_and_exclusion():
    g = dgl.rand_graph(100, 500)
    g.ndata["h"] = torch.randn(100, 4)
    nc = DL.NodeCollator(g, torch.arange(20), DL.NeighborSampler([3, 3]))
    loader = torch.utils.data.DataLoader(nc.dataset, collate_fn=nc.collate, batch_size=5)
    inp, out, blocks = next(iter(loader))
    assert out.tolist() == [0, 1, 2, 3, 4] and len(blocks) == 2 and blocks[-1].num_dst_nodes() == 5
    ec = DL.EdgeCollator(g, torch.arange(50), DL.NeighborSampler([3]), negative_sampler=DL.negative_sampler.Uniform(2))
    inp, pos, neg, blocks = ec.collate(ec.dataset[:4])
    assert pos.num_edges() == 4 and neg.num_edges() == 8 and len(blocks) == 1
    hg = dgl.heterograph({("u", "r", "v"): ([0, 1], [1, 0]), ("v", "s", "u"): ([0], [1])})
    hc = DL.NodeCollator(hg, {"u": torch.tensor([0, 1]), "v": torch.tensor([1])}, DL.MultiLayerFullNeighborSampler(1))
    assert len(hc.dataset) == 3
    _, out, _ = hc.collate(hc.dataset)
    assert out["u"].tolist() == [0, 1] and out["v"].tolist() == [1]
    rev = torch.arange(500).flip(0)
    assert DL.find_exclude_eids(g, "reverse_id", torch.tensor([1, 2]), reverse_eid_map=rev).tolist() == [1, 2, 498, 497]
    assert DL.find_exclude_eids(g, None, torch.tensor([1])) is None
    assert DL.find_exclude_eids(g, "self", torch.tensor([1]), output_device="cpu").tolist() == [1]
    ex = DL.find_exclude_eids(hg, "reverse_types", {"r": torch.tensor([0])}, reverse_etype_map={"r": "s", "s": "r"})
    assert ex[("u", "r", "v")].tolist() == [0] and ex[("v", "s", "u")].tolist() == [0]


def test_parent_storage_columns_are_stripped_for_ipc():
    g = dgl.rand_graph(2000, 8000)
    g.ndata["h"] = torch.randn(2000, 64)
    g.edata["w"] = torch.randn(8000, 16)
    sg = g.subgraph(torch.tensor([1, 2, 3]))
    DL.remove_parent_storage_columns(sg, g)
    assert len(pickle.dumps(sg)) < 20000                       # not 2000*64*4 bytes
    back = pickle.loads(pickle.dumps(sg))
    DL.restore_parent_storage_columns(back, g)
    assert torch.equal(back.ndata["h"], g.ndata["h"][1:4])
    assert torch.equal(back.edata["w"], g.edata["w"][back.edata[dgl.EID].long()])
    blocks = DL.NeighborSampler([2]).sample(g, torch.tensor([5, 6]))[2]
    DL.remove_parent_storage_columns(blocks, g)
    DL.restore_parent_storage_columns(blocks, g)


def test_collate_and_worker_wrappers():
    g = dgl.rand_graph(50, 200)
    cw = DL.CollateWrapper(DL.NeighborSampler([2]).sample, g, False, torch.device("cpu"))
    _, out, blocks = cw(torch.tensor([1, 2]))
    assert out.tolist() == [1, 2] and len(blocks) == 1
    seen = []
    DL.WorkerInitWrapper(seen.append)(3)
    DL.WorkerInitWrapper(None)(0)
    assert seen == [3]
