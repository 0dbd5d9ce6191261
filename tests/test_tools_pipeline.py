"""chunk_graph -> random_partition -> dispatch_data -> verify_partitions (reference tools/ pipeline) on a toy graph."""
import os
import sys

import torch

import dgl_b200 as dgl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tools", "partition_algo"))


def test_chunk_partition_dispatch_verify(tmp_path):
    from chunk_graph import chunk_graph, load_chunked
    from dispatch_data import dispatch
    from random_partition import random_partition
    from verify_partitions import verify
    g = dgl.rand_graph(500, 4000, idtype=torch.int64)
    g.ndata["feat"] = torch.randn(500, 6)
    g.ndata["label"] = torch.randint(0, 4, (500,))
    g.edata["w"] = torch.rand(4000)
    chunked = str(tmp_path / "chunked")
    meta = chunk_graph(g, "toy", num_chunks=3, output_path=chunked)
    assert meta["num_nodes_per_chunk"] == [[167, 167, 166]] and len(meta["edges"]["_N:_E:_N"]["data"]) == 3
    g2, _ = load_chunked(chunked)
    assert torch.equal(g2.ndata["feat"], g.ndata["feat"]) and torch.equal(g2.edges()[0], g.edges()[0])
    parts = str(tmp_path / "parts")
    random_partition(chunked, 3, parts)
    cfg = dispatch(chunked, parts, str(tmp_path / "out"))
    assert verify(chunked, cfg)
    # the result is consumable by the distributed runtime
    from dgl_b200 import distributed as dd
    os.environ["DGL_DIST_MODE"] = "standalone"
    dd.initialize(part_config=cfg)
    dg = dd.DistGraph("toy")
    assert dg.num_nodes() == 500 and dg.num_edges() == 4000
    dd.exit_client()
