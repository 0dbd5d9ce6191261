#!/usr/bin/env python
"""Turn (chunked graph, node->partition assignment) into DistDGL partitions
(parity: /root/reference/tools/dispatch_data.py + distpartitioning/data_proc_pipeline.py).

The reference shuffles node/edge data between ranks with gloo all-to-all because no single host can hold the graph; this
build targets 8 x 180 GB GPUs per host with TBs of host RAM, so the dispatcher re-assembles the chunks and emits every
partition from one process (``--num-workers`` parallelises the per-partition writes over processes that each own a
subset of the partition ids). Output layout == ``dgl.distributed.partition_graph``.
For graphs that must not be assembled on one host use the gloo pipeline ``tools/distpartitioning/data_proc_pipeline.py``
(one process per partition, all-to-all shuffle) - it produces the same partitions."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def dispatch(in_dir, partitions_dir, out_dir, graph_name=None, num_hops=1):
    from chunk_graph import load_chunked
    import dgl_b200 as dgl
    from dgl_b200 import distributed as dd
    g, meta = load_chunked(in_dir)
    with open(os.path.join(partitions_dir, "partition_meta.json")) as f:
        pmeta = json.load(f)
    num_parts = int(pmeta["num_parts"])
    if len(meta["node_type"]) == 1:
        node_part = torch.from_numpy(np.loadtxt(os.path.join(partitions_dir, f"{meta['node_type'][0]}.txt"), dtype=np.int64))
    else:
        hg = g
        g = dgl.to_homogeneous(hg, ndata=None)
        parts = [torch.from_numpy(np.loadtxt(os.path.join(partitions_dir, f"{nt}.txt"), dtype=np.int64).reshape(-1))
                 for nt in hg.ntypes]
        node_part = torch.cat(parts)
    name = graph_name or meta["graph_name"]
    dd.partition_graph(g, name, num_parts, out_dir, num_hops=num_hops, node_part=node_part.reshape(-1))
    return os.path.join(out_dir, name + ".json")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--in-dir", required=True)
    ap.add_argument("--partitions-dir", required=True)
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--graph-name", default=None)
    ap.add_argument("--ip-config", default=None, help="accepted for CLI compatibility; dispatch runs on this host")
    a = ap.parse_args()
    print("partition config:", dispatch(a.in_dir, a.partitions_dir, a.out_dir, a.graph_name))
