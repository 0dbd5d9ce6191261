"""``dgl.sparse`` and GraphBolt in five minutes (reference notebooks/sparse/quickstart.ipynb, notebooks/graphbolt/
walkthrough.ipynb)."""
import torch

import dgl_b200 as dgl
import dgl_b200.graphbolt as gb
import dgl_b200.sparse as dglsp

# ---- sparse matrices: a GCN layer as X' = D^-1/2 (A + I) D^-1/2 X W
i = torch.tensor([[0, 1, 2, 2, 3], [1, 2, 0, 3, 0]])
A = dglsp.spmatrix(i, shape=(4, 4))
A_hat = A + dglsp.identity((4, 4))
D_inv_sqrt = dglsp.diag(A_hat.sum(1)) ** -0.5
X = torch.randn(4, 3)
H = D_inv_sqrt @ A_hat @ D_inv_sqrt @ X
print(H.shape, dglsp.softmax(A_hat).val.shape)

# ---- GraphBolt: item set -> sampler -> feature fetch -> (copy to GPU) as a datapipe
g = dgl.rand_graph(1000, 8000)
graph = gb.from_dglgraph(g)
features = gb.BasicFeatureStore({("node", None, "feat"): gb.TorchBasedFeature(torch.randn(1000, 16))})
items = gb.ItemSet((torch.arange(200), torch.randint(0, 3, (200,))), names=("seeds", "labels"))
dp = gb.ItemSampler(items, batch_size=50, shuffle=True).sample_neighbor(graph, [5, 5]).fetch_feature(features, ["feat"])
for mb in gb.DataLoader(dp):
    blocks = mb.blocks
    print(mb.seeds.shape, [b.num_edges() for b in blocks], mb.node_features["feat"].shape)
    break
