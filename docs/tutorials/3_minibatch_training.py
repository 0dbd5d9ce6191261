"""Stochastic training with neighbour sampling (reference tutorials/large/L1_large_node_classification.py and
notebooks/stochastic_training): the generic pipeline - ``NeighborSampler`` + ``DataLoader`` produce blocks (message-flow
graphs), every layer consumes one block. On a B200 the same model can be handed to ``FusedSAGETrainer`` which replays the
whole step (sampling, compaction, tcgen05 layers, loss, backward, Adam) as one CUDA graph."""
import torch
import torch.nn.functional as F

import dgl_b200 as dgl
from dgl_b200.dataloading import DataLoader, NeighborSampler
from dgl_b200.models import GraphSAGE

device = "cuda" if torch.cuda.is_available() else "cpu"
g = dgl.data.rmat_graph(5000, 60000, seed=0).to(device)
x = torch.randn(g.num_nodes(), 32, device=device)
y = torch.randint(0, 5, (g.num_nodes(),), device=device)
train_ids = torch.arange(2000, device=device, dtype=g.idtype)

loader = DataLoader(g, train_ids, NeighborSampler([10, 10]), batch_size=500, shuffle=True, device=device)
model = GraphSAGE(32, 64, 5, num_layers=2, dropout=0.0).to(device)
opt = torch.optim.Adam(model.parameters(), lr=5e-3)
for input_nodes, output_nodes, blocks in loader:
    logits = model(blocks, x[input_nodes.long()])
    loss = F.cross_entropy(logits, y[output_nodes.long()])
    opt.zero_grad()
    loss.backward()
    opt.step()
print("last mini-batch loss", float(loss))

if device == "cuda":
    from dgl_b200.models import FusedSAGETrainer
    tr = FusedSAGETrainer(g, x, y, 5, hidden=64, fanouts=[10, 10], batch_size=500)
    print("fused engine loss", tr.train_step(train_ids[:500].cpu()))
