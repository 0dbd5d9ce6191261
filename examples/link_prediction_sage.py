"""Link prediction with a GraphSAGE encoder, edge mini-batches and uniform negative sampling (reference
examples/pytorch/graphsage/link_pred.py): as_edge_prediction_sampler + NeighborSampler + dot-product scorer.
python examples/link_prediction_sage.py --device cuda"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgl_b200 as dgl
from dgl_b200.dataloading import DataLoader, NeighborSampler, as_edge_prediction_sampler, negative_sampler
from dgl_b200.models import LinkPredictor


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    g = dgl.to_bidirected(dgl.data.rmat_graph(20000, 200000, seed=0)).to(a.device)
    x = torch.randn(g.num_nodes(), 64, device=a.device)
    sampler = as_edge_prediction_sampler(NeighborSampler([10, 10]), negative_sampler=negative_sampler.Uniform(1))
    eids = torch.arange(g.num_edges(), device=a.device, dtype=g.idtype)
    dl = DataLoader(g, eids, sampler, batch_size=2048, shuffle=True, device=a.device)
    model = LinkPredictor(64, 128).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    step = 0
    for inp, pos, neg, blocks in dl:
        h = model.encode(blocks, x[inp.long()])
        loss = model.loss(model.score(pos, h), model.score(neg, h))
        opt.zero_grad(); loss.backward(); opt.step()
        if step % 10 == 0:
            print(f"step {step:3d} loss {loss.item():.4f}")
        step += 1
        if step >= a.steps:
            break
    return loss.item()


if __name__ == "__main__":
    main()
