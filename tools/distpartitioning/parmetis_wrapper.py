#!/usr/bin/env python
"""preprocess -> pm_dglpart (under mpirun) -> postprocess (reference tools/distpartitioning/parmetis_wrapper.py).
ParMETIS itself is an external binary; when ``pm_dglpart`` is not on PATH the wrapper falls back to the in-tree
multilevel / LDG partitioner (dgl_b200.partition) so that the rest of the pipeline can still be driven end to end."""
import argparse
import os
import shutil
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tools.distpartitioning.parmetis_postprocess import post_process  # noqa: E402
from tools.distpartitioning.parmetis_preprocess import gen_parmetis_input  # noqa: E402
from tools.distpartitioning.utils import read_json  # noqa: E402


def run(in_dir, work_dir, partitions_dir, num_parts, hostfile=None):
    gen_parmetis_input(in_dir, work_dir)
    meta = read_json(os.path.join(in_dir, "metadata.json"))
    name = meta["graph_name"]
    out_file = os.path.join(work_dir, f"{name}_part.{num_parts}")
    if shutil.which("pm_dglpart"):
        cmd = ["mpirun", "-np", str(num_parts)] + (["--hostfile", hostfile] if hostfile else []) + ["pm_dglpart", name, "1"]
        subprocess.check_call(cmd, cwd=work_dir)
    else:
        import torch
        from tools.chunk_graph import load_chunked
        import dgl_b200 as dgl
        g, _ = load_chunked(in_dir)
        hg = g if g.is_homogeneous else dgl.to_homogeneous(g)
        part = dgl.partition.metis_partition_assignment(hg, num_parts)
        np.savetxt(out_file, np.stack([np.arange(hg.num_nodes()), part.numpy()], 1), fmt="%d")
    return post_process(in_dir, out_file, partitions_dir)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--in-dir", required=True)
    ap.add_argument("--work-dir", required=True)
    ap.add_argument("--partitions-dir", required=True)
    ap.add_argument("--num-parts", type=int, required=True)
    ap.add_argument("--hostfile", default=None)
    a = ap.parse_args()
    print("assignment:", run(a.in_dir, a.work_dir, a.partitions_dir, a.num_parts, a.hostfile))
