"""Heterogeneous graph transformer on a synthetic author-paper-venue graph (reference examples/pytorch/hgt/): the
heterograph is homogenised, node / edge type ids select the per-type projections inside ``HGTConv``.
Task: predict the field of each paper; only authorship structure carries the signal.
python examples/hgt_hetero.py --device cuda"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import dgl_b200 as dgl
from dgl_b200.models import HGT


def make_graph(n_paper=400, n_author=120, fields=4, seed=0):
    gen = torch.Generator().manual_seed(seed)
    field_p = torch.randint(0, fields, (n_paper,), generator=gen)
    field_a = torch.randint(0, fields, (n_author,), generator=gen)
    src, dst = [], []
    for p in range(n_paper):                       # 3 authors per paper, 85% from the paper's own field
        same = torch.nonzero(field_a == field_p[p]).view(-1)
        for _ in range(3):
            pool = same if float(torch.rand(1, generator=gen)) < 0.85 and len(same) else torch.arange(n_author)
            src.append(int(pool[int(torch.randint(0, len(pool), (1,), generator=gen))])); dst.append(p)
    a, p = torch.tensor(src), torch.tensor(dst)
    g = dgl.heterograph({("author", "writes", "paper"): (a, p), ("paper", "written_by", "author"): (p, a)},
                        {"author": n_author, "paper": n_paper})
    g.nodes["author"].data["x"] = F.one_hot(field_a, fields).float() + 0.5 * torch.randn(n_author, fields, generator=gen)
    g.nodes["paper"].data["x"] = torch.zeros(n_paper, fields)            # papers have no features of their own
    g.nodes["paper"].data["y"] = field_p
    return g, fields


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=40)
    a = ap.parse_args()
    torch.manual_seed(0)
    hg, fields = make_graph()
    g = dgl.to_homogeneous(hg, ndata=["x"]).to(a.device)
    ntype, etype = g.ndata[dgl.NTYPE], g.edata[dgl.ETYPE]
    paper_tid = hg.ntypes.index("paper")
    papers = torch.nonzero(ntype == paper_tid).view(-1)
    y = hg.nodes["paper"].data["y"].to(a.device)
    train = torch.rand(papers.shape[0], device=a.device) < 0.4
    model = HGT(fields, 32, fields, len(hg.ntypes), len(hg.etypes), heads=4, num_layers=2).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    for ep in range(a.epochs):
        out = model(g, g.ndata["x"], ntype, etype)[papers]
        loss = F.cross_entropy(out[train], y[train])
        opt.zero_grad(); loss.backward(); opt.step()
    acc = float((out.argmax(1) == y)[~train].float().mean())
    print(f"loss {loss.item():.4f} held-out paper accuracy {acc:.3f}")
    return acc


if __name__ == "__main__":
    main()
