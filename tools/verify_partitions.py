#!/usr/bin/env python
"""Check a set of DistDGL partitions against the chunked source graph
(parity: /root/reference/tools/verify_partitions.py + verification_utils.py): every node/edge owned exactly once,
edges map back to source edges, node features equal the source rows."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def verify(orig_dir, part_config):
    from chunk_graph import load_chunked
    import dgl_b200 as dgl
    from dgl_b200 import distributed as dd
    g, _ = load_chunked(orig_dir)
    if not g.is_homogeneous:
        g = dgl.to_homogeneous(g)
    with open(part_config) as f:
        num_parts = json.load(f)["num_parts"]
    u, v = g.edges()
    seen_n = torch.zeros(g.num_nodes(), dtype=torch.int64)
    seen_e = torch.zeros(g.num_edges(), dtype=torch.int64)
    import verification_utils as vu
    schema = vu.load_part_schema(part_config)
    for p in range(num_parts):
        pg, nf, ef, gpb, name, nts, ets = dd.load_partition(part_config, p)
        vu.verify_partition_data_types(pg, p)
        vu.verify_partition_formats(pg, None, p)
        vu.verify_metadata_counts(schema, pg, p)
        inner = pg.ndata["inner_node"].bool()
        seen_n[pg.ndata["orig_id"][inner]] += 1
        ie = pg.edata["inner_edge"].bool()
        oe = pg.edata["orig_id"]
        seen_e[oe[ie]] += 1
        pu, pv = pg.edges()
        assert torch.equal(u[oe], pg.ndata["orig_id"][pu]) and torch.equal(v[oe], pg.ndata["orig_id"][pv]), \
            f"partition {p}: edge endpoints do not match the source graph"
        for k, t in nf.items():
            feat = k.split("/", 1)[1]
            if feat in g.ndata:
                assert torch.equal(t, g.ndata[feat][pg.ndata["orig_id"][inner]]), f"partition {p}: node feature {feat}"
    assert bool((seen_n == 1).all()), "every node must be an inner node of exactly one partition"
    assert bool((seen_e == 1).all()), "every edge must be an inner edge of exactly one partition"
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--orig-dataset-dir", required=True)
    ap.add_argument("--part-graph-dir", required=True, help="partition config json")
    a = ap.parse_args()
    verify(a.orig_dataset_dir, a.part_graph_dir)
    print("partitions verified")
