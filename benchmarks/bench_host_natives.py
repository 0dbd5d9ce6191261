"""Host-side native kernels vs the library / Python alternatives they replace (CPU only, wall clock, best of 3).

    python benchmarks/bench_host_natives.py --out profiles/host_natives.json

Rows: SpGEMM (csrc/cpu/structure.cc) vs torch.sparse.mm; kd-tree kNN (spatial.cc) vs chunked cdist+topk; BFS / RCM
(traversal.cc) vs scipy.sparse.csgraph; farthest point sampling vs the vectorised torch loop; random.choice without
replacement vs numpy."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import torch
from scipy.sparse import csgraph

import dgl_b200 as dgl
import dgl_b200.sparse as dglsp
from dgl_b200.transforms import functional as TF


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.manual_seed(0)
    rows = []

    def add(name, ours, theirs, what, **extra):
        rows.append(dict(op=name, ours_ms=round(ours, 2), baseline_ms=round(theirs, 2), baseline=what,
                         speedup=round(theirs / ours, 2), **extra))
        print(rows[-1])

    # SpGEMM A @ A, 100k x 100k, 2M nnz
    n, nnz = 100_000, 2_000_000
    key = torch.unique(torch.randint(0, n * n, (nnz,)))
    A = dglsp.from_coo(key // n, key % n, torch.randn(key.shape[0]), (n, n))
    ta = dglsp.to_torch_sparse_coo(A).coalesce()
    add("spgemm_100k_2Mnnz", best(lambda: dglsp.spspmm(A, A)), best(lambda: torch.sparse.mm(ta, ta), 1), "torch.sparse.mm",
        out_nnz=int(dglsp.spspmm(A, A).nnz))

    # kNN, 200k points in 3-D, k = 16
    x = torch.randn(200_000, 3)
    add("knn_200k_3d_k16", best(lambda: TF.kdtree_knn(x, x, 16)),
        best(lambda: torch.cat([torch.topk(torch.cdist(x[s:s + 4096], x), 16, largest=False)[1] for s in range(0, 40960, 4096)]), 1)
        * (200_000 / 40960), "chunked cdist+topk (extrapolated from 40960 queries)")

    # traversal / reordering on a 1M-node random graph
    g = dgl.rand_graph(1_000_000, 8_000_000)
    u, v = [t.numpy() for t in g.edges()]
    m = sp.coo_matrix((np.ones(len(u), dtype=np.int8), (u, v)), shape=(g.num_nodes(),) * 2).tocsr()
    g.create_formats_()
    add("bfs_1M_8M", best(lambda: dgl.bfs_nodes_generator(g, 0)),
        best(lambda: csgraph.breadth_first_order(m, 0, return_predecessors=False)), "scipy breadth_first_order")
    def scipy_rcm():
        mm = sp.coo_matrix((np.ones(len(u), dtype=np.int8), (u, v)), shape=(g.num_nodes(),) * 2).tocsr()
        return csgraph.reverse_cuthill_mckee((mm + mm.T).tocsr(), True)

    add("rcm_1M_8M_from_edge_list", best(lambda: TF.rcmk_perm(g), 1), best(scipy_rcm, 1),
        "scipy: coo->csr, A+A^T, reverse_cuthill_mckee (both sides start from the edge list)")

    # farthest point sampling, 8 clouds x 20k points -> 512
    pos = torch.randn(8, 20_000, 3)

    def torch_fps():
        B, N, _ = pos.shape
        dist = torch.full((B, N), float("inf"))
        cur = torch.zeros(B, dtype=torch.long)
        ar = torch.arange(B)
        for _ in range(512):
            dist = torch.minimum(dist, ((pos - pos[ar, cur].unsqueeze(1)) ** 2).sum(-1))
            cur = dist.argmax(1)

    add("fps_8x20k_512", best(lambda: dgl.geometry.farthest_point_sampler(pos, 512, start_idx=0)), best(torch_fps, 1),
        "vectorised torch loop")

    # choice without replacement: 1M draws from 100M
    add("choice_1M_of_100M", best(lambda: dgl.random.choice(100_000_000, 1_000_000, replace=False)),
        best(lambda: np.random.default_rng(0).choice(100_000_000, 1_000_000, replace=False), 1), "numpy Generator.choice")

    meta = {"threads": torch.get_num_threads(), "cpu": os.uname().machine, "rows": rows}
    if a.out:
        with open(a.out, "w") as f:
            json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
