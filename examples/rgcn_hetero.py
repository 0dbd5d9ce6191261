"""R-GCN node classification on an ogbn-mag-shaped random heterograph (BASELINE config 4: RelGraphConv typed g-SpMM).
Usage: python examples/rgcn_hetero.py --device cuda --scale 1.0"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.models import RGCN

# ogbn-mag: node types and (src, rel, dst, #edges) incl. reverse relations used by the reference example
MAG_NODES = {"paper": 736389, "author": 1134649, "institution": 8740, "field_of_study": 59965}
MAG_RELS = [("author", "affiliated_with", "institution", 1043998), ("author", "writes", "paper", 7145660),
            ("paper", "cites", "paper", 5416271), ("paper", "has_topic", "field_of_study", 7505078)]


def mag_shaped(scale=1.0, device="cpu", seed=0):
    gen = torch.Generator().manual_seed(seed)
    nn = {k: max(8, int(v * scale)) for k, v in MAG_NODES.items()}
    data = {}
    for s, r, d, e in MAG_RELS:
        e = max(16, int(e * scale))
        u = torch.randint(0, nn[s], (e,), generator=gen)
        v = torch.randint(0, nn[d], (e,), generator=gen)
        data[(s, r, d)] = (u, v)
        data[(d, "rev_" + r, s)] = (v, u)
    return dgl.heterograph(data, num_nodes_dict=nn, idtype=torch.int32, device=device), nn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--scale", type=float, default=0.01)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--hidden", type=int, default=64)
    args = ap.parse_args()
    hg, nn_ = mag_shaped(args.scale, args.device)
    g = dgl.to_homogeneous(hg)
    etypes = g.edata[dgl.ETYPE]
    ntypes = g.ndata[dgl.NTYPE]
    num_rels = len(hg.canonical_etypes)
    C = 349  # venues
    paper_t = hg.ntypes.index("paper")
    target = (ntypes == paper_t).nonzero().squeeze(1)
    labels = torch.randint(0, C, (target.numel(),), device=g.device)
    model = RGCN(g.num_nodes(), args.hidden, C, num_rels, num_bases=2).to(args.device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    for epoch in range(args.epochs):
        if args.device != "cpu":
            torch.cuda.synchronize()
        t0 = time.time()
        logits = model(g, etypes)
        loss = F.cross_entropy(logits[target], labels)
        opt.zero_grad(); loss.backward(); opt.step()
        if args.device != "cpu":
            torch.cuda.synchronize()
        print(f"epoch {epoch:3d} loss {loss.item():.4f} time {1e3 * (time.time() - t0):.1f} ms")
    return loss.item()


if __name__ == "__main__":
    main()
