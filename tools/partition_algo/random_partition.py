#!/usr/bin/env python
"""Random node-to-partition assignment for a chunked graph: writes ``<ntype>.txt`` with one partition id per node
(parity: /root/reference/tools/partition_algo/random_partition.py)."""
import argparse
import json
import os

import numpy as np


def random_partition(metadata_path, num_parts, output_path, seed=0):
    with open(os.path.join(metadata_path, "metadata.json")) as f:
        meta = json.load(f)
    os.makedirs(output_path, exist_ok=True)
    rng = np.random.RandomState(seed)
    for nt, sizes in zip(meta["node_type"], meta["num_nodes_per_chunk"]):
        n = int(sum(sizes))
        np.savetxt(os.path.join(output_path, f"{nt}.txt"), rng.randint(0, num_parts, n), fmt="%d")
    with open(os.path.join(output_path, "partition_meta.json"), "w") as f:
        json.dump({"algo_name": "random", "num_parts": num_parts, "version": "1.0.0"}, f)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--in_dir", required=True)
    ap.add_argument("--out_dir", required=True)
    ap.add_argument("--num_partitions", type=int, required=True)
    a = ap.parse_args()
    random_partition(a.in_dir, a.num_partitions, a.out_dir)
