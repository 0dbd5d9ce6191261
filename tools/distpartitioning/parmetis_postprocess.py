#!/usr/bin/env python
"""ParMETIS output -> per-node-type partition assignment files (reference tools/distpartitioning/parmetis_postprocess.py).

ParMETIS (``pm_dglpart``) writes lines ``<global node id> <partition id>``; the pipeline wants ``<ntype>.txt`` with one
partition id per type-local node id plus ``partition_meta.json`` ({"num_parts", "algo_name", "version"})."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tools.distpartitioning import constants as C  # noqa: E402
from tools.distpartitioning.dataset_utils import type_offsets  # noqa: E402
from tools.distpartitioning.utils import read_json  # noqa: E402


def post_process(in_dir, parmetis_output_file, partitions_dir):
    meta = read_json(os.path.join(in_dir, "metadata.json"))
    noff, _ = type_offsets(meta)
    arr = np.loadtxt(parmetis_output_file, dtype=np.int64).reshape(-1, 2)
    part = np.full(int(noff[-1]), -1, dtype=np.int64)
    part[arr[:, 0]] = arr[:, 1]
    if (part < 0).any():
        raise ValueError("the ParMETIS output does not cover every node")
    os.makedirs(partitions_dir, exist_ok=True)
    for ti, nt in enumerate(meta[C.STR_NODE_TYPE]):
        np.savetxt(os.path.join(partitions_dir, f"{nt}.txt"), part[noff[ti]: noff[ti + 1]], fmt="%d")
    with open(os.path.join(partitions_dir, "partition_meta.json"), "w") as f:
        json.dump({"num_parts": int(part.max()) + 1, "algo_name": "parmetis", "version": "1.0.0"}, f)
    return partitions_dir


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--in-dir", "--schema_path", dest="in_dir", required=True)
    ap.add_argument("--parmetis-output-file", "--parmetis_output_file", dest="pfile", required=True)
    ap.add_argument("--partitions-dir", "--partitions_dir", dest="pdir", required=True)
    a = ap.parse_args()
    print("assignment files written to", post_process(a.in_dir, a.pfile, a.pdir))
