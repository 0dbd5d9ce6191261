"""The distributed shuffle (reference tools/distpartitioning/data_shuffle.py ``gen_dist_partitions``), one process per
partition on gloo:

  1. every rank loads the partition ids of ITS slice of the global node-id space (DistLookupService);
  2. edges: a rank reads the chunks assigned to it, asks the lookup service for the partition of every destination and
     sends each edge (with its feature rows) to the partition that owns the destination (all-to-all-v);
  3. nodes: the range owner sends every node id (and its feature rows) to the node's partition;
  4. partitions number their nodes / edges contiguously (globalids), publish orig -> shuffled ids, translate the edge
     end points through the lookup service;
  5. every rank assembles its partition (inner nodes + 1-hop halo) and writes it; rank 0 writes the metadata JSON.
"""
import json
import os

import numpy as np
import torch
import torch.distributed as dist

from . import constants as C
from .convert_partition import build_partition, write_metadata, write_partition
from .dataset_utils import read_my_edges, read_rows, type_offsets
from .dist_lookup import DistLookupService
from .globalids import assign_shuffle_global_ids
from .gloo_wrapper import alltoallv_cpu, gather_metadata_json


def _load_assignment_slice(partitions_dir, meta, lo, hi):
    """Partition ids of global node ids [lo, hi): the assignment is stored as one text file per node type."""
    noff, _ = type_offsets(meta)
    parts = []
    for ti, nt in enumerate(meta[C.STR_NODE_TYPE]):
        a, b = max(lo, int(noff[ti])), min(hi, int(noff[ti + 1]))
        if a < b:
            ids = np.loadtxt(os.path.join(partitions_dir, f"{nt}.txt"), dtype=np.int64).reshape(-1)
            parts.append(torch.from_numpy(ids[a - int(noff[ti]): b - int(noff[ti])].copy()))
    return torch.cat(parts) if parts else torch.empty(0, dtype=torch.int64)


def _split_by(dest, world, *tensors):
    order = torch.argsort(dest, stable=True)
    counts = torch.bincount(dest, minlength=world).tolist()
    return [list(torch.split(t[order], counts)) for t in tensors]


def gen_dist_partitions(rank, world, in_dir, partitions_dir, out_dir, graph_name=None):
    with open(os.path.join(in_dir, "metadata.json")) as f:
        meta = json.load(f)
    with open(os.path.join(partitions_dir, "partition_meta.json")) as f:
        pmeta = json.load(f)
    assert int(pmeta["num_parts"]) == world, "run one process per partition"
    noff, eoff = type_offsets(meta)
    N, E = int(noff[-1]), int(eoff[-1])
    hetero = len(meta[C.STR_NODE_TYPE]) > 1 or len(meta[C.STR_EDGE_TYPE]) > 1
    lo, hi = N * rank // world, N * (rank + 1) // world
    lookup = DistLookupService(rank, world, N, _load_assignment_slice(partitions_dir, meta, lo, hi))

    # ---- 2. edges -> partition of the destination ------------------------------------------------------------
    edges, eranges = read_my_edges(in_dir, meta, rank, world)
    dest = lookup.get_partition_ids(edges[:, 1])
    (sp,) = _split_by(dest, world, edges)
    my_edges = torch.cat(alltoallv_cpu(rank, world, sp))
    # only features defined for EVERY edge type are shuffled (the decision must not depend on the rank: collectives)
    efeat_names = sorted(k for k in {k for feats in meta[C.STR_EDGE_DATA].values() for k in feats}
                         if all(k in meta[C.STR_EDGE_DATA].get(et, {}) for et in meta[C.STR_EDGE_TYPE]))
    my_efeats = {}
    for k in efeat_names:
        rows = []
        for et, a, b in eranges:
            ei = meta[C.STR_EDGE_TYPE].index(et)
            rows.append(read_rows(in_dir, meta[C.STR_EDGE_DATA][et][k], meta[C.STR_NUM_EDGES_PER_CHUNK][ei], a, b))
        local = torch.cat(rows) if rows else None
        tail = dist_obj_tail(local, rank, world)
        local = local if local is not None else torch.empty((0,) + tail[0], dtype=tail[1])
        (sf,) = _split_by(dest, world, local)
        my_efeats[k] = torch.cat(alltoallv_cpu(rank, world, sf))

    # ---- 3. nodes (ids + features) -> their partition -----------------------------------------------------------
    gids = torch.arange(lo, hi, dtype=torch.int64)
    (sn,) = _split_by(lookup.part_ids, world, gids)
    my_nodes = torch.cat(alltoallv_cpu(rank, world, sn))
    nfeat_names = sorted(k for k in {k for feats in meta[C.STR_NODE_DATA].values() for k in feats}
                         if all(k in meta[C.STR_NODE_DATA].get(nt, {}) for nt in meta[C.STR_NODE_TYPE]))
    my_nfeats = {}
    for k in nfeat_names:
        rows = []
        for ti, nt in enumerate(meta[C.STR_NODE_TYPE]):
            a, b = max(lo, int(noff[ti])), min(hi, int(noff[ti + 1]))
            if a < b:
                rows.append(read_rows(in_dir, meta[C.STR_NODE_DATA][nt][k], meta[C.STR_NUM_NODES_PER_CHUNK][ti],
                                      a - int(noff[ti]), b - int(noff[ti])))
        local = torch.cat(rows) if rows else None
        tail = dist_obj_tail(local, rank, world)
        local = local if local is not None else torch.empty((0,) + tail[0], dtype=tail[1])
        (sf,) = _split_by(lookup.part_ids, world, local)
        my_nfeats[k] = torch.cat(alltoallv_cpu(rank, world, sf))

    # ---- 4. shuffled ids ----------------------------------------------------------------------------------------
    norder = torch.argsort(my_nodes)
    my_nodes = my_nodes[norder]
    my_nfeats = {k: v[norder] for k, v in my_nfeats.items()}
    new_nids, node_ranges = assign_shuffle_global_ids(rank, world, my_nodes.shape[0])
    lookup.set_shuffle_nids(my_nodes, new_nids)
    eorder = torch.argsort(my_edges[:, 3])
    my_edges = my_edges[eorder]
    my_efeats = {k: v[eorder] for k, v in my_efeats.items()}
    new_eids, edge_ranges = assign_shuffle_global_ids(rank, world, my_edges.shape[0])
    new_src = lookup.get_shuffle_nids(my_edges[:, 0])
    new_dst = lookup.get_shuffle_nids(my_edges[:, 1])

    # ---- 5. assemble + write --------------------------------------------------------------------------------------
    def halo_orig(new_ids):
        """COLLECTIVE: original ids of shuffled ids, answered by the partitions that own them."""
        ends = torch.tensor([r[1] for r in node_ranges], dtype=torch.int64)
        owner = torch.searchsorted(ends, new_ids, right=True)
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=world).tolist()
        asked = alltoallv_cpu(rank, world, list(torch.split(new_ids[order], counts)))
        start = node_ranges[rank][0]
        back = alltoallv_cpu(rank, world, [my_nodes[q - start] for q in asked])
        out = torch.empty_like(new_ids)
        out[order] = torch.cat(back) if back else new_ids.new_empty(0)
        return out

    gdict, order = build_partition(rank, new_nids, my_nodes, torch.stack([new_src, new_dst], 1), new_eids, my_edges[:, 3],
                                   node_ranges, halo_orig)
    nf = {f"_N/{k}": v for k, v in my_nfeats.items()}
    ef = {f"_N:_E:_N/{k}": v[order] for k, v in my_efeats.items()}
    if hetero:
        ntype_of = torch.searchsorted(torch.from_numpy(noff[1:].copy()), my_nodes, right=True)
        nf["_N/_TYPE"] = ntype_of
        ef["_N:_E:_N/_TYPE"] = my_edges[:, 2][order]
    files = write_partition(out_dir, rank, gdict, nf, ef)
    all_files = gather_metadata_json(files, rank, world)
    path = None
    if rank == 0:
        path = write_metadata(out_dir, graph_name or meta[C.STR_GRAPH_NAME], N, E, node_ranges, edge_ranges, all_files)
    dist.barrier()
    return path


def dist_obj_tail(local, rank, world):
    """Agree on (trailing shape, dtype) of a feature across ranks (a rank may hold no rows of it)."""
    mine = None if local is None else (tuple(local.shape[1:]), str(local.dtype))
    out = [None] * world
    dist.all_gather_object(out, mine)
    shape, dt = next(o for o in out if o is not None)
    return tuple(shape), getattr(torch, dt.replace("torch.", ""))
