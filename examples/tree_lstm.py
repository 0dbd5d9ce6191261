"""Child-sum Tree-LSTM on synthetic expression trees (reference examples/pytorch/tree_lstm/): batched trees are
propagated leaf-to-root with ``dgl.prop_nodes_topo`` (native Kahn levels -> one fused message-passing call per level).
Task: leaves carry +1/-1 tokens, the label of the tree is the sign of their sum.
python examples/tree_lstm.py --device cuda"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F
import dgl_b200 as dgl


def random_tree(gen, n_leaves):
    """Binary tree built bottom-up; edges point child -> parent. Returns (graph, leaf mask, token ids)."""
    nodes = list(range(n_leaves))
    src, dst, nxt = [], [], n_leaves
    while len(nodes) > 1:
        i = int(torch.randint(0, len(nodes) - 1, (1,), generator=gen))
        a, b = nodes.pop(i), nodes.pop(i)
        src += [a, b]; dst += [nxt, nxt]
        nodes.insert(i, nxt); nxt += 1
    g = dgl.graph((torch.tensor(src), torch.tensor(dst)), num_nodes=nxt)
    tok = torch.full((nxt,), 2)
    tok[:n_leaves] = torch.randint(0, 2, (n_leaves,), generator=gen)
    g.ndata["tok"] = tok
    g.ndata["leaf"] = torch.arange(nxt) < n_leaves
    label = int((tok[:n_leaves] * 2 - 1).sum() > 0)
    return g, label


class ChildSumTreeLSTM(nn.Module):
    def __init__(self, vocab, h):
        super().__init__()
        self.emb = nn.Embedding(vocab, h)
        self.W_iou = nn.Linear(h, 3 * h)
        self.U_iou = nn.Linear(h, 3 * h, bias=False)
        self.U_f = nn.Linear(h, h)
        self.W_f = nn.Linear(h, h)
        self.out = nn.Linear(h, 2)
        self.h = h

    def message(self, edges):
        return {"h": edges.src["h"], "c": edges.src["c"]}

    def reduce(self, nodes):
        h_child, c_child = nodes.mailbox["h"], nodes.mailbox["c"]                 # [N, children, H]
        f = torch.sigmoid(self.U_f(h_child) + self.W_f(nodes.data["x"]).unsqueeze(1))
        return {"iou": nodes.data["iou_x"] + self.U_iou(h_child.sum(1)), "c_in": (f * c_child).sum(1)}

    def apply(self, nodes):
        i, o, u = torch.chunk(nodes.data["iou"], 3, 1)
        c = torch.sigmoid(i) * torch.tanh(u) + nodes.data["c_in"]
        return {"h": torch.sigmoid(o) * torch.tanh(c), "c": c}

    def forward(self, bg):
        x = self.emb(bg.ndata["tok"])
        n = bg.num_nodes()
        bg.ndata.update({"x": x, "iou_x": self.W_iou(x), "iou": self.W_iou(x), "c_in": torch.zeros(n, self.h, device=x.device),
                         "h": torch.zeros(n, self.h, device=x.device), "c": torch.zeros(n, self.h, device=x.device)})
        # leaves first (no incoming messages), then level by level up to the roots
        bg.apply_nodes(self.apply, torch.nonzero(bg.ndata["leaf"]).view(-1))
        dgl.prop_nodes_topo(bg, self.message, self.reduce, apply_node_func=self.apply)
        roots = torch.nonzero(bg.out_degrees() == 0).view(-1)
        return self.out(bg.ndata["h"][roots])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--epochs", type=int, default=12)
    ap.add_argument("--trees", type=int, default=256)
    a = ap.parse_args()
    gen = torch.Generator().manual_seed(0)
    data = [random_tree(gen, int(torch.randint(3, 9, (1,), generator=gen)) * 2 - 1) for _ in range(a.trees)]
    torch.manual_seed(0)
    model = ChildSumTreeLSTM(3, 32).to(a.device)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    acc = 0.0
    for ep in range(a.epochs):
        correct = 0
        for s in range(0, len(data), 32):
            gs, ys = zip(*data[s:s + 32])
            bg, y = dgl.batch(list(gs)).to(a.device), torch.tensor(ys, device=a.device)
            out = model(bg)
            loss = F.cross_entropy(out, y)
            opt.zero_grad(); loss.backward(); opt.step()
            correct += int((out.argmax(1) == y).sum())
        acc = correct / len(data)
        print(f"epoch {ep:2d} loss {loss.item():.4f} train acc {acc:.3f}")
    return acc


if __name__ == "__main__":
    main()
