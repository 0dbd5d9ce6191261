"""Shared-memory visit counter + top-k (csrc/cuda/visit_count.cu) against a dense bincount oracle, and the PinSAGE
sampler built on it against the walk it was given."""
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200._native import loader
from dgl_b200.base import id_code

pytestmark = pytest.mark.gpu


def _topk(visits, K):
    S, M = visits.shape
    node = torch.empty((S, K), dtype=visits.dtype, device="cuda")
    cnt = torch.empty((S, K), dtype=torch.int32, device="cuda")
    loader.cuda_lib().fn("dglb200_cuda_visit_topk", "ipliippp")(
        id_code(visits.dtype), visits.data_ptr(), S, M, K, node.data_ptr(), cnt.data_ptr(), loader.stream_ptr())
    torch.cuda.synchronize()
    return node.cpu(), cnt.cpu()


@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
@pytest.mark.parametrize("M,K,n", [(40, 5, 30), (700, 10, 200), (3000, 16, 100000), (8000, 3, 50)])
def test_topk_matches_bincount_oracle(idt, M, K, n):
    torch.manual_seed(M + K)
    S = 300
    visits = torch.randint(0, n, (S, M), dtype=idt)
    visits[torch.rand(S, M) < 0.2] = -1
    visits[0] = -1                      # a seed without any visit
    visits[1, :] = 7                    # a single distinct node
    node, cnt = _topk(visits.cuda(), K)
    for s in range(S):
        row = visits[s][visits[s] >= 0].long()
        if row.numel() == 0:
            assert cnt[s].tolist() == [0] * K and node[s].tolist() == [-1] * K
            continue
        uniq, c = torch.unique(row, return_counts=True)
        order = sorted(range(uniq.numel()), key=lambda i: (-int(c[i]), int(uniq[i])))[:K]
        want_nodes = [int(uniq[i]) for i in order]
        want_cnt = [int(c[i]) for i in order]
        k = len(order)
        assert node[s, :k].tolist() == want_nodes and cnt[s, :k].tolist() == want_cnt
        assert cnt[s, k:].tolist() == [0] * (K - k)


def test_pinsage_sampler_uses_the_device_counter():
    torch.manual_seed(0)
    g = dgl.heterograph({("item", "by", "user"): (torch.randint(0, 200, (3000,)), torch.randint(0, 80, (3000,))),
                         ("user", "buys", "item"): (torch.randint(0, 80, (3000,)), torch.randint(0, 200, (3000,)))}).to("cuda")
    sampler = dgl.sampling.PinSAGESampler(g, "item", "user", 3, 0.5, 20, 4)
    seeds = torch.arange(0, 200, 7, device="cuda")
    frontier = sampler(seeds)
    src, dst = frontier.edges()
    w = frontier.edata["weights"]
    assert frontier.num_nodes() == 200 and w.dtype == torch.int64 and bool((w > 0).all())
    assert set(dst.tolist()) <= set(seeds.tolist())
    deg = torch.bincount(dst, minlength=200)
    assert int(deg.max()) <= 4
    # inside a seed the neighbours are listed by decreasing visit count
    for s in seeds.tolist():
        ws = w[dst == s].tolist()
        assert ws == sorted(ws, reverse=True) and sum(ws) <= 20 * 3
