"""Distributed partitioning pipeline (parity: /root/reference/tools/distpartitioning/): turns a CHUNKED graph plus a
node -> partition assignment (ParMETIS / random_partition output) into DistDGL partitions with one process per
partition, exchanging nodes / edges / features with gloo all-to-all - no process ever holds the whole graph.

    python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 \\
        tools/distpartitioning/data_proc_pipeline.py --in-dir chunked --partitions-dir parts --out-dir out

Modules: constants, array_readwriter (numpy / csv / parquet), gloo_wrapper (all-to-all-v on gloo), dist_lookup (global id
-> partition id / shuffled id service), globalids (contiguous id assignment), data_shuffle (the pipeline),
convert_partition (partition assembly + writer), parmetis_preprocess / parmetis_postprocess / parmetis_wrapper, utils."""
