#!/usr/bin/env python
"""Chunked graph -> ParMETIS input files (reference tools/distpartitioning/parmetis_preprocess.py).

Writes, for every chunk index i: ``<out>/parmetis_nfiles/node-i.txt`` with one line per node
``<node type id> <weight...> <type-local id>`` and ``<out>/parmetis_efiles/edge-i.txt`` with ``<global src> <global dst>``
(homogeneous ids), plus ``parmetis_nfiles.txt`` / ``parmetis_efiles.txt`` listing them with their global id ranges and
``<graph_name>_stats.txt`` (num_nodes num_edges num_node_weights) - the inputs of ``pm_dglpart``."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tools.distpartitioning import constants as C  # noqa: E402
from tools.distpartitioning.array_readwriter import get_array_parser  # noqa: E402
from tools.distpartitioning.dataset_utils import type_offsets  # noqa: E402
from tools.distpartitioning.utils import read_json  # noqa: E402


def gen_parmetis_input(in_dir, out_dir):
    meta = read_json(os.path.join(in_dir, "metadata.json"))
    noff, _ = type_offsets(meta)
    ntid = {t: i for i, t in enumerate(meta[C.STR_NODE_TYPE])}
    T = len(meta[C.STR_NODE_TYPE])
    os.makedirs(os.path.join(out_dir, "parmetis_nfiles"), exist_ok=True)
    os.makedirs(os.path.join(out_dir, "parmetis_efiles"), exist_ok=True)
    nfiles, efiles = [], []
    num_chunks = len(meta[C.STR_NUM_NODES_PER_CHUNK][0])
    for c in range(num_chunks):
        lines, lo_g = [], None
        for ti, sizes in enumerate(meta[C.STR_NUM_NODES_PER_CHUNK]):
            start = int(np.sum(sizes[:c]))
            ids = np.arange(start, start + sizes[c], dtype=np.int64)
            onehot = np.zeros((ids.shape[0], T), dtype=np.int64)
            onehot[:, ti] = 1                                        # one balance constraint per node type
            lines.append(np.concatenate([np.full((ids.shape[0], 1), ti), onehot, ids[:, None]], 1))
        arr = np.concatenate(lines)
        path = os.path.join(out_dir, "parmetis_nfiles", f"node-{c}.txt")
        np.savetxt(path, arr, fmt="%d")
        nfiles.append(path)
    for c in range(num_chunks):
        rows = []
        for ei, et in enumerate(meta[C.STR_EDGE_TYPE]):
            s, _, d = et.split(":")
            spec = meta[C.STR_EDGES][et]
            if c < len(spec[C.STR_DATA]):
                e = np.asarray(get_array_parser(spec[C.STR_FORMAT][C.STR_NAME]).read(os.path.join(in_dir, spec[C.STR_DATA][c])))
                rows.append(np.stack([e[:, 0] + noff[ntid[s]], e[:, 1] + noff[ntid[d]]], 1))
        arr = np.concatenate(rows) if rows else np.zeros((0, 2), dtype=np.int64)
        path = os.path.join(out_dir, "parmetis_efiles", f"edge-{c}.txt")
        np.savetxt(path, arr, fmt="%d")
        efiles.append(path)
    with open(os.path.join(out_dir, "parmetis_nfiles.txt"), "w") as f:
        f.write("\n".join(nfiles) + "\n")
    with open(os.path.join(out_dir, "parmetis_efiles.txt"), "w") as f:
        f.write("\n".join(efiles) + "\n")
    _, eoff = type_offsets(meta)
    with open(os.path.join(out_dir, f"{meta[C.STR_GRAPH_NAME]}_stats.txt"), "w") as f:
        f.write(f"{int(noff[-1])} {int(eoff[-1])} {T}\n")
    return nfiles, efiles


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--in-dir", "--input_dir", dest="in_dir", required=True)
    ap.add_argument("--out-dir", "--output_dir", dest="out_dir", required=True)
    a = ap.parse_args()
    n, e = gen_parmetis_input(a.in_dir, a.out_dir)
    print(f"wrote {len(n)} node files and {len(e)} edge files under {a.out_dir}")
