"""Small helpers of the pipeline (reference tools/distpartitioning/utils.py)."""
import json
import os

import numpy as np


def read_json(path):
    with open(path) as f:
        return json.load(f)


def write_json(obj, path):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump(obj, f, indent=2)


def get_idranges(names, counts_per_chunk):
    """{type: [[start, end] per chunk]} and {type: [global start, global end]} in the homogeneous id space."""
    per_chunk, glob, off = {}, {}, 0
    for name, counts in zip(names, counts_per_chunk):
        starts = np.concatenate([[0], np.cumsum(counts)])
        per_chunk[name] = [[int(starts[i]), int(starts[i + 1])] for i in range(len(counts))]
        glob[name] = [off, off + int(starts[-1])]
        off += int(starts[-1])
    return per_chunk, glob


def get_node_types(meta):
    return {nt: i for i, nt in enumerate(meta["node_type"])}, {i: nt for i, nt in enumerate(meta["node_type"])}


def get_edge_types(meta):
    return {et: i for i, et in enumerate(meta["edge_type"])}, {i: et for i, et in enumerate(meta["edge_type"])}


def memory_snapshot(tag, rank):
    try:
        import psutil
        rss = psutil.Process().memory_info().rss / 2 ** 20
        return f"[rank {rank}] {tag}: rss {rss:.0f} MiB"
    except Exception:  # noqa: BLE001
        return f"[rank {rank}] {tag}"
