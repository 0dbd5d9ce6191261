"""SAGEConv against dense linear algebra for every aggregator, both projection orders (in > out and in < out), edge
weights, blocks, tuple inputs and empty graphs."""
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200.nn import SAGEConv


@pytest.mark.parametrize("agg", ["mean", "gcn", "pool"])
@pytest.mark.parametrize("dims", [(12, 5), (5, 12)])
@pytest.mark.parametrize("weighted", [False, True])
def test_matches_dense(agg, dims, weighted):
    torch.manual_seed(0)
    g = dgl.rand_graph(40, 200)
    u, v = [t.long() for t in g.edges()]
    i, o = dims
    ew = torch.rand(200) if weighted else None
    conv = SAGEConv(i, o, agg)
    x = torch.randn(40, i)
    A = torch.zeros(40, 40).index_put_((v, u), torch.ones(200), accumulate=True)
    W = A if ew is None else torch.zeros(40, 40).index_put_((v, u), ew, accumulate=True)
    deg = A.sum(1, keepdim=True)
    out = conv(g, x, edge_weight=ew)
    if agg == "mean":
        ref = conv.fc_self(x) + conv.fc_neigh((W @ x) / deg.clamp(min=1))
    elif agg == "gcn":
        ref = conv.fc_neigh((W @ x + x) / (deg + 1)) + conv.bias
    else:
        h = torch.relu(conv.fc_pool(x)).detach()
        m = torch.full((40, i), float("-inf"))
        for e in range(200):
            m[v[e]] = torch.maximum(m[v[e]], h[u[e]] * (1.0 if ew is None else ew[e]))
        m[torch.isinf(m)] = 0
        ref = conv.fc_self(x) + conv.fc_neigh(m)
    assert torch.allclose(out, ref, atol=1e-5)


@pytest.mark.parametrize("agg", ["mean", "gcn", "pool", "lstm"])
def test_blocks_tuple_inputs_and_empty_graphs(agg):
    torch.manual_seed(1)
    g = dgl.rand_graph(40, 200)
    blk = dgl.to_block(dgl.sampling.sample_neighbors(g, torch.arange(10), 4), torch.arange(10))
    xs = torch.randn(blk.num_src_nodes(), 6)
    conv = SAGEConv(6, 4, agg)
    a, b = conv(blk, xs), conv(blk, (xs, xs[:10]))
    assert a.shape == (10, 4) and torch.allclose(a, b)
    a.sum().backward()
    assert conv.fc_neigh.weight.grad is not None
    empty = dgl.graph(([], []), num_nodes=3)
    assert SAGEConv(4, 2, agg)(empty, torch.randn(3, 4)).shape == (3, 2)


def test_options_and_errors():
    g = dgl.rand_graph(10, 30)
    with pytest.raises(dgl.DGLError):
        SAGEConv(4, 2, "median")
    with pytest.raises(dgl.DGLError):
        SAGEConv(4, 2, "mean")(g, torch.randn(10, 4), edge_weight=torch.rand(5))
    conv = SAGEConv(4, 2, "gcn", bias=False)
    assert conv.bias is None and not hasattr(conv, "fc_self")
    normed = SAGEConv(4, 3, "mean", norm=torch.nn.LayerNorm(3), activation=torch.relu, feat_drop=0.5)
    normed.eval()
    assert normed(g, torch.randn(10, 4)).shape == (10, 3)
    assert set(SAGEConv(4, 2, "pool").state_dict()) == {"fc_pool.weight", "fc_pool.bias", "fc_neigh.weight", "fc_self.weight", "fc_self.bias"}
