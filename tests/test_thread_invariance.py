"""Host kernels must give bit-identical results whatever the OpenMP thread count: they are written without atomics or
thread-dependent work splitting, so any difference between 1 and N threads is a race or a scheduling dependence.
(ThreadSanitizer cannot judge OpenMP regions here — libgomp is not instrumented — so this is the race check for them.)"""
import pytest
import torch

import dgl_b200 as dgl
import dgl_b200.sparse as dglsp
from dgl_b200 import utils as U
from dgl_b200.transforms.functional import kdtree_knn, rcmk_perm


def _with_threads(n, fn):
    old = U.get_num_threads()
    U.set_num_threads(n)
    try:
        return fn()
    finally:
        U.set_num_threads(old)


def _cases():
    torch.manual_seed(0)
    n = 4000
    key = torch.unique(torch.randint(0, n * n, (40000,)))
    A = dglsp.from_coo(key // n, key % n, torch.randn(key.shape[0]), (n, n))
    x = torch.randn(20000, 3)
    pos = torch.randn(16, 3000, 3)
    big = torch.randn(1, 50000, 3)
    g = dgl.rand_graph(20000, 160000)
    g.ndata["h"] = torch.randn(20000, 16)
    w = torch.rand(g.num_edges(), 1)
    p = torch.rand(5000)

    def spgemm():
        C = dglsp.spspmm(A, A)
        return torch.cat([C.val, C.row.float(), C.col.float()])

    def seeded(fn):
        def run():
            dgl.seed(11)
            return fn()
        return run

    return {
        "spgemm": spgemm,
        "sparse_add": lambda: (A + A.t()).val,
        "kdtree": lambda: kdtree_knn(x, x[:5000], 8)[0],
        "fps_many_clouds": lambda: dgl.geometry.farthest_point_sampler(pos, 64, start_idx=0),
        "fps_one_cloud": lambda: dgl.geometry.farthest_point_sampler(big, 32, start_idx=0),
        "choice_uniform": seeded(lambda: dgl.random.choice(1000, 100000)),
        "choice_weighted": seeded(lambda: dgl.random.choice(5000, 50000, prob=p)),
        "rcm": lambda: torch.as_tensor(rcmk_perm(g)),
        "gspmm_copy_u_sum": lambda: dgl.ops.copy_u_sum(g, g.ndata["h"]),
        "gspmm_u_mul_e_max": lambda: dgl.ops.u_mul_e_max(g, g.ndata["h"], w),
        "gsddmm_u_dot_v": lambda: dgl.ops.u_dot_v(g, g.ndata["h"], g.ndata["h"]),
        "edge_softmax": lambda: dgl.ops.edge_softmax(g, w),
        "neighbor_sampling": seeded(lambda: torch.stack(dgl.sampling.sample_neighbors(g, torch.arange(5000), 5).edges())),
        "to_block": lambda: dgl.to_block(dgl.sampling.sample_neighbors(g, torch.arange(3000), -1), torch.arange(3000)).srcdata[dgl.NID],
    }


@pytest.mark.parametrize("name", sorted(_cases()))
def test_result_is_independent_of_thread_count(name):
    fn = _cases()[name]
    one = _with_threads(1, fn)
    many = _with_threads(8, fn)
    again = _with_threads(3, fn)
    assert torch.equal(one, many), name
    assert torch.equal(one, again), name
