"""Every reference module path that has no same-named module here is synthesised by the alias import hook
(dgl_b200/_pathshim.py + _alias_table.py); all of their names must resolve. Also covers the last batch of small helpers those shims point at."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dgl_b200 as dgl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shims():
    from dgl_b200 import _pathshim
    yield from _pathshim.alias_modules()


def test_every_shim_name_resolves():
    mods = sorted(_shims())
    assert len(mods) > 100
    for mod in mods:
        m = importlib.import_module(mod)
        for name in m.__all__:
            assert getattr(m, name) is not None, (mod, name)
        with pytest.raises(AttributeError):
            getattr(m, "definitely_not_here")


def test_representative_reference_imports():
    from dgl_b200.data.citation_graph import CoraGraphDataset, load_cora  # noqa: F401
    from dgl_b200.distgnn.partition.libra_partition import libra_partition  # noqa: F401
    from dgl_b200.distributed.graph_services import SamplingRequest, sample_neighbors  # noqa: F401
    from dgl_b200.distributed.nn.pytorch.sparse_emb import DistEmbedding  # noqa: F401
    from dgl_b200.distributed.optim.pytorch.utils import alltoallv_cpu  # noqa: F401
    from dgl_b200.graphbolt.dataloader import DataLoader  # noqa: F401
    from dgl_b200.graphbolt.impl.gpu_feature_cache import GPUFeatureCache  # noqa: F401
    from dgl_b200.nn.pytorch.conv.sageconv import SAGEConv  # noqa: F401
    from dgl_b200.nn.pytorch.explain.subgraphx import MCTSNode, SubgraphX  # noqa: F401
    from dgl_b200.optim.pytorch.sparse_optim import SparseAdam  # noqa: F401
    from dgl_b200.sampling.node2vec_randomwalk import node2vec_random_walk  # noqa: F401
    assert SAGEConv is dgl.nn.SAGEConv and SparseAdam is dgl.optim.SparseAdam


def test_superpixel_graph_construction():
    from dgl_b200.data.superpixel import SuperPixelDataset, compute_adjacency_matrix_images, compute_edges_list, sigma
    rng = np.random.RandomState(0)
    coord, feat = rng.rand(20, 2), rng.rand(20, 3)
    A = compute_adjacency_matrix_images(coord, feat)
    assert A.shape == (20, 20) and np.allclose(A, A.T) and np.all(np.diag(A) == 0) and A.max() <= 1.0
    assert sigma(np.ones((3, 3)), kth=8).tolist() == [[1.0 + 1e-8]] * 3
    nb, val = compute_edges_list(A, kth=9)
    assert nb.shape == (20, 8) and val.shape == (20, 8)
    for i in range(20):                                    # the 8 strongest affinities of every row
        assert set(nb[i]) == set(np.argsort(A[i])[-8:]) and i not in nb[i]
    nb, val = compute_edges_list(A[:5, :5], kth=9)         # tiny graph -> fully connected without self loops
    assert nb.shape == (5, 4) and all(i not in nb[i] for i in range(5))
    assert compute_adjacency_matrix_images(coord, feat, use_feat=False).shape == (20, 20)
    ds = SuperPixelDataset(name="MNIST", scale=0.0005)
    assert type(ds).__name__ == "MNISTSuperPixelDataset" and len(ds) > 0
    with pytest.raises(dgl.DGLError):
        SuperPixelDataset(name="SVHN")


def test_distgnn_tools(tmp_path):
    from scipy.io import mmwrite
    import scipy.sparse as sp
    from dgl_b200 import distgnn
    g = dgl.rand_graph(200, 1200)
    parts, cnt = distgnn.libra_partition(3, g, str(tmp_path / "parts"))
    assert len(parts) == 3 and sum(p.num_edges() for p in parts) == 1200
    rep = distgnn.rep_per_node(str(tmp_path / "parts"), 3)
    assert sum(rep.values()) == int((cnt.clamp(min=1) - 1).sum())
    assert all(0 < v < 3 for v in rep.values())
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with pytest.raises(dgl.DGLError):
            distgnn.download_proteins()
        mmwrite("proteins.mtx", sp.random(50, 50, density=0.1, format="coo", random_state=0))
        pg = distgnn.proteins_mtx2dgl(train_size=30, test_size=10, val_size=5)
        assert pg.num_nodes() == 50 and int(pg.ndata["train_mask"].sum()) == 30 and int(pg.ndata["val_mask"].sum()) == 5
        distgnn.save(pg, "prot")
        back = distgnn.load_proteins("prot")
        assert back.num_edges() == pg.num_edges() and "label" in back.ndata
    finally:
        os.chdir(cwd)


def _a2a_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dgl_b200.distributed.optim import alltoall, alltoallv
    outs = [torch.empty(2) for _ in range(world)]
    alltoall(rank, world, outs, [torch.full((2,), float(rank * 10 + d)) for d in range(world)])
    vouts = [None] * world
    alltoallv(rank, world, vouts, [torch.arange(rank + d + 1, dtype=torch.float32).view(-1, 1) + rank for d in range(world)])
    torch.save((outs, vouts), os.path.join(out_dir, f"a2a{rank}.pt"))
    dist.destroy_process_group()


def test_alltoall_helpers_on_gloo(tmp_path):
    world = 3
    mp.spawn(_a2a_worker, args=(world, 29700 + os.getpid() % 200, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        outs, vouts = torch.load(os.path.join(tmp_path, f"a2a{r}.pt"))
        for s in range(world):
            assert outs[s].tolist() == [float(s * 10 + r)] * 2                  # what rank s addressed to rank r
            assert vouts[s].shape == (s + r + 1, 1) and vouts[s][0, 0] == s


def test_service_request_classes_and_extra_metadata():
    from dgl_b200 import distributed as dd
    from dgl_b200.distributed import rpc_api
    assert issubclass(dd.SamplingRequest, rpc_api.Request) and issubclass(dd.SubgraphResponse, rpc_api.Response)
    req = dd.EdgesRequest(torch.tensor([1]), 7)
    assert req.order_id == 7 and dd.InDegreeRequest(torch.tensor([0]), 1).n.tolist() == [0]
    with pytest.raises(NotImplementedError):
        dd.Collator().collate([1])
    assert dd.optim.initializer((2, 3), torch.float32).eq(0).all()
    from dgl_b200.graphbolt.impl.ondisk_metadata import ExtraMetaData
    m = ExtraMetaData(a=1, extra_fields={"b": 2})
    assert m["extra_fields"] == {"b": 2} and m.a == 1
