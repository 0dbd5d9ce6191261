#!/usr/bin/env python
"""Entry point of the distributed partitioning pipeline (reference tools/distpartitioning/data_proc_pipeline.py):

    python -m torch.distributed.run --nproc-per-node <num_parts> --master-addr 127.0.0.1 \\
        tools/distpartitioning/data_proc_pipeline.py --in-dir chunked --partitions-dir parts --out-dir out
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--in-dir", required=True, help="chunked graph (tools/chunk_graph.py)")
    ap.add_argument("--partitions-dir", required=True, help="<ntype>.txt assignment files + partition_meta.json")
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--graph-name", default=None)
    a = ap.parse_args()
    dist.init_process_group("gloo")
    from tools.distpartitioning.data_shuffle import gen_dist_partitions
    path = gen_dist_partitions(dist.get_rank(), dist.get_world_size(), a.in_dir, a.partitions_dir, a.out_dir, a.graph_name)
    if dist.get_rank() == 0:
        print("partition config:", path)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
