"""DeepWalk node embeddings on a two-community graph, evaluated with a linear probe (reference
examples/pytorch/deepwalk/, nn/pytorch/network_emb.py). Random walks come from the native walk kernel; the skip-gram
objective with negative sampling is ``dgl.nn.DeepWalk``; embeddings train with SparseAdam.
python examples/deepwalk_embedding.py"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgl_b200 as dgl
from dgl_b200.nn import DeepWalk


def two_communities(n=200, p_in=0.08, p_out=0.004, seed=0):
    gen = torch.Generator().manual_seed(seed)
    comm = (torch.arange(n) >= n // 2).long()
    prob = torch.where(comm[:, None] == comm[None, :], torch.tensor(p_in), torch.tensor(p_out))
    adj = torch.triu(torch.rand(n, n, generator=gen) < prob, 1)
    u, v = adj.nonzero(as_tuple=True)
    return dgl.graph((torch.cat([u, v]), torch.cat([v, u])), num_nodes=n), comm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=8)
    ap.add_argument("--dim", type=int, default=16)
    a = ap.parse_args()
    torch.manual_seed(0)
    dgl.seed(0)
    g, comm = two_communities()
    model = DeepWalk(g, emb_dim=a.dim, walk_length=20, window_size=4, negative_size=3)
    loader = torch.utils.data.DataLoader(torch.arange(g.num_nodes()), batch_size=64, shuffle=True, collate_fn=model.sample)
    opt = torch.optim.SparseAdam(model.parameters(), lr=0.02)
    for ep in range(a.epochs):
        for walks in loader:
            loss = model(walks)
            opt.zero_grad(); loss.backward(); opt.step()
        print(f"epoch {ep} loss {loss.item():.4f}")
    emb = model.node_embed.weight.detach()
    # linear probe: least-squares fit on half of the nodes, accuracy on the other half
    perm = torch.randperm(g.num_nodes())
    tr, te = perm[: len(perm) // 2], perm[len(perm) // 2:]
    X = torch.cat([emb, torch.ones(emb.shape[0], 1)], 1)
    w = torch.linalg.lstsq(X[tr], (comm[tr].float() * 2 - 1).unsqueeze(1)).solution
    acc = float(((X[te] @ w).view(-1) > 0).long().eq(comm[te]).float().mean())
    print(f"linear-probe accuracy {acc:.3f}")
    return acc


if __name__ == "__main__":
    main()
