#!/usr/bin/env python
"""Convert graph files written by the reference (``dgl.save_graphs`` -> libdgl binary format) into this framework's
tensor-only archive (``dgl_b200.save_graphs``).

    python tools/convert_dgl_bin.py in.bin out.pt [--ref /path/to/site-packages-with-dgl]

The reference's reader does the parsing (it must be importable; ``--ref`` defaults to ``baseline/_ref`` of this
repository); structure, node / edge features, batch information and labels are carried over."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def convert_graph(rg, dgl_b200):
    """Reference DGLGraph -> dgl_b200.DGLGraph (same type names, edge order, features)."""
    data = {}
    for cet in rg.canonical_etypes:
        u, v = rg.edges(etype=cet, order="eid")
        data[cet] = (u, v)
    num_nodes = {nt: rg.num_nodes(nt) for nt in rg.ntypes}
    if len(rg.canonical_etypes) == 1 and len(rg.ntypes) == 1:
        g = dgl_b200.graph(data[rg.canonical_etypes[0]], num_nodes=rg.num_nodes(), idtype=rg.idtype)
        for k, v in rg.ndata.items():
            g.ndata[k] = v
        for k, v in rg.edata.items():
            g.edata[k] = v
    else:
        g = dgl_b200.heterograph(data, num_nodes_dict=num_nodes, idtype=rg.idtype)
        for nt in rg.ntypes:
            for k, v in rg.nodes[nt].data.items():
                g.nodes[nt].data[k] = v
        for cet in rg.canonical_etypes:
            for k, v in rg.edges[cet].data.items():
                g.edges[cet].data[k] = v
    if rg.batch_size > 1:
        for nt in rg.ntypes:
            g.set_batch_num_nodes(rg.batch_num_nodes(nt) if len(rg.ntypes) == 1 else {t: rg.batch_num_nodes(t) for t in rg.ntypes})
            break
        for cet in rg.canonical_etypes:
            g.set_batch_num_edges(rg.batch_num_edges(cet) if len(rg.canonical_etypes) == 1
                                  else {c: rg.batch_num_edges(c) for c in rg.canonical_etypes})
            break
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--ref", default=os.path.join(ROOT, "baseline", "_ref"))
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    import dgl_b200
    sys.path.insert(0, args.ref)
    os.environ.setdefault("DGLBACKEND", "pytorch")
    try:
        import dgl as ref
    except Exception as e:  # noqa: BLE001
        raise SystemExit(f"the reference build is needed to parse its binary format ({type(e).__name__}: {e})")
    graphs, labels = ref.load_graphs(args.src)
    out = [convert_graph(g, dgl_b200) for g in graphs]
    dgl_b200.save_graphs(args.dst, out, {k: torch.as_tensor(v) for k, v in labels.items()})
    print(f"converted {len(out)} graph(s), {len(labels)} label tensor(s): {args.dst}")


if __name__ == "__main__":
    main()
