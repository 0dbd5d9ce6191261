"""Assemble and write one partition in the layout of ``dgl.distributed.partition_graph`` (reference
tools/distpartitioning/convert_partition.py): inner nodes first, 1-hop halo sources after them, all edges are inner
edges (destination owned by this partition)."""
import json
import os

import torch


def build_partition(part_id, inner_new_ids, inner_orig_ids, edges_new, edge_new_ids, edge_orig_ids, node_ranges,
                    halo_orig_lookup):
    """edges_new: [m, 2] (new src, new dst). ``halo_orig_lookup(new_ids) -> orig ids`` resolves halo nodes (collective)."""
    start = int(node_ranges[part_id][0])
    n_inner = inner_new_ids.shape[0]
    src, dst = edges_new[:, 0], edges_new[:, 1]
    is_inner_src = (src >= start) & (src < start + n_inner)
    halo = torch.unique(src[~is_inner_src])
    halo_orig = halo_orig_lookup(halo)
    nodes = torch.cat([inner_new_ids, halo])
    ends = torch.tensor([r[1] for r in node_ranges], dtype=torch.int64)
    part_of = torch.searchsorted(ends, nodes, right=True)
    # local ids: inner nodes are contiguous, halo nodes through a sorted lookup
    if halo.numel():
        pos = torch.searchsorted(halo, src).clamp(max=halo.numel() - 1)     # only meaningful for halo sources
        local_src = torch.where(is_inner_src, src - start, n_inner + pos)
    else:
        local_src = src - start
    local_dst = dst - start
    order = torch.argsort(edge_new_ids)
    return {"src": local_src[order], "dst": local_dst[order], "num_nodes": int(nodes.shape[0]),
            "ndata": {"_ID": nodes, "inner_node": (part_of == part_id).to(torch.uint8), "part_id": part_of,
                      "orig_id": torch.cat([inner_orig_ids, halo_orig])},
            "edata": {"_ID": edge_new_ids[order], "inner_edge": torch.ones(order.shape[0], dtype=torch.uint8),
                      "orig_id": edge_orig_ids[order]}}, order


def write_partition(out_dir, part_id, graph_dict, node_feats, edge_feats):
    pdir = os.path.join(out_dir, f"part{part_id}")
    os.makedirs(pdir, exist_ok=True)
    # same artifact names and binary containers as the reference pipeline (tools/distpartitioning/convert_partition.py)
    import dgl_b200
    from dgl_b200.data import dgl_binary
    pg = dgl_b200.graph((graph_dict["src"], graph_dict["dst"]), num_nodes=graph_dict["num_nodes"])
    for k, v in graph_dict["ndata"].items():
        pg.ndata[k] = v
    for k, v in graph_dict["edata"].items():
        pg.edata[k] = v
    dgl_binary.write_graphs(os.path.join(pdir, "graph.dgl"), [pg])
    dgl_binary.write_tensors(os.path.join(pdir, "node_feat.dgl"), node_feats)
    dgl_binary.write_tensors(os.path.join(pdir, "edge_feat.dgl"), edge_feats)
    return {"node_feats": f"part{part_id}/node_feat.dgl", "edge_feats": f"part{part_id}/edge_feat.dgl",
            "part_graph": f"part{part_id}/graph.dgl"}


def write_metadata(out_dir, graph_name, num_nodes, num_edges, node_ranges, edge_ranges, part_files):
    meta = {"graph_name": graph_name, "num_nodes": int(num_nodes), "num_edges": int(num_edges), "part_method": "given",
            "num_parts": len(node_ranges), "halo_hops": 1, "node_map": {"_N": node_ranges},
            "edge_map": {"_N:_E:_N": edge_ranges}, "ntypes": {"_N": 0}, "etypes": {"_N:_E:_N": 0}}
    for p, files in enumerate(part_files):
        meta[f"part-{p}"] = files
    path = os.path.join(out_dir, f"{graph_name}.json")
    with open(path, "w") as f:
        json.dump(meta, f, indent=2)
    return path
