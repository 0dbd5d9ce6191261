"""Individual checks behind ``verify_partitions.py`` (reference: tools/verification_utils.py:24-293), each usable on its
own from a notebook or a test. All of them raise ``AssertionError`` with the offending partition / type in the message.

The checks work on what ``dgl_b200.distributed.load_partition`` returns, so they apply equally to partitions written by
this framework, by its tools/distpartitioning pipeline and by the reference (same file names and containers)."""
import json
import os

import numpy as np
import torch

NID, EID, NTYPE, ETYPE = "_ID", "_ID", "_TYPE", "_TYPE"
# dtypes the loaders and samplers rely on (reference distributed/partition.py RESERVED_FIELD_DTYPE)
RESERVED_NODE_FIELDS = {"inner_node": (torch.uint8, torch.bool), NID: (torch.int64,), "part_id": (torch.int64,),
                        NTYPE: (torch.int16, torch.int32, torch.int64)}
RESERVED_EDGE_FIELDS = {"inner_edge": (torch.uint8, torch.bool), EID: (torch.int64,), ETYPE: (torch.int16, torch.int32, torch.int64)}


def read_file(fname, ftype=None):
    """Array from ``.npy`` / ``.txt`` / ``.csv`` / ``.parquet`` (type from the extension when ``ftype`` is None)."""
    ftype = ftype or os.path.splitext(fname)[1].lstrip(".")
    if ftype in ("numpy", "npy"):
        return np.load(fname)
    if ftype in ("txt", "csv"):
        return np.loadtxt(fname, delimiter="," if ftype == "csv" else None, ndmin=1)
    if ftype == "parquet":
        import pyarrow.parquet as pq
        t = pq.read_table(fname)
        return np.stack([c.to_numpy() for c in t.columns], 1).squeeze()
    raise ValueError(f"unknown file type {ftype!r} for {fname}")


def verify_partition_data_types(part_g, partid=None):
    for k, ok in RESERVED_NODE_FIELDS.items():
        if k in part_g.ndata:
            assert part_g.ndata[k].dtype in ok, f"partition {partid}: ndata[{k!r}] has dtype {part_g.ndata[k].dtype}"
    for k, ok in RESERVED_EDGE_FIELDS.items():
        if k in part_g.edata:
            assert part_g.edata[k].dtype in ok, f"partition {partid}: edata[{k!r}] has dtype {part_g.edata[k].dtype}"


def verify_partition_formats(part_g, formats=None, partid=None):
    """``formats``: comma separated list that must have been materialised (default: at least coo)."""
    created = part_g.formats()["created"]
    for f in (formats.split(",") if formats else ["coo"]):
        assert f.strip() in created, f"partition {partid}: sparse format {f!r} was not stored (created: {created})"


def _inner_mask(part_g, field, type_field, type_id):
    m = part_g.ndata[field].bool() if field == "inner_node" else part_g.edata[field].bool()
    data = part_g.ndata if field == "inner_node" else part_g.edata
    if type_id is not None and type_field in data:
        m = m & (data[type_field] == type_id)
    return m


def verify_metadata_counts(part_schema, part_g, partid):
    """Inner node / edge counts of a partition equal the widths of its ranges in the partition metadata."""
    for ntype, tid in part_schema["ntypes"].items():
        lo, hi = part_schema["node_map"][ntype][partid]
        single = len(part_schema["ntypes"]) == 1
        got = int(_inner_mask(part_g, "inner_node", NTYPE, None if single else tid).sum())
        assert got == hi - lo, f"partition {partid}: {got} inner nodes of type {ntype}, metadata says {hi - lo}"
    for etype, tid in part_schema["etypes"].items():
        lo, hi = part_schema["edge_map"][etype][partid]
        single = len(part_schema["etypes"]) == 1
        got = int(_inner_mask(part_g, "inner_edge", ETYPE, None if single else tid).sum())
        assert got == hi - lo, f"partition {partid}: {got} inner edges of type {etype}, metadata says {hi - lo}"


def verify_graph_feats(g, gpb, part_g, node_feats, edge_feats, orig_nids=None, orig_eids=None, partid=None):
    """Feature rows stored with a partition equal the rows of the source graph ``g``. ``orig_nids`` / ``orig_eids`` map
    the partitioner's new per-type ids back to source ids (dict type -> tensor); identity when None."""
    inner_n = part_g.ndata[NID][part_g.ndata["inner_node"].bool()]
    tids, per_type = gpb.map_to_per_ntype(inner_n)
    for t, ntype in enumerate(g.ntypes):
        ids = per_type[tids == t]
        assert bool((gpb.nid2partid(ids, ntype) == gpb.partid).all()), f"partition {partid}: foreign inner node of {ntype}"
        src_rows = ids if orig_nids is None else torch.as_tensor(orig_nids[ntype])[ids]
        local = gpb.nid2localnid(ids, gpb.partid, ntype)
        for name, val in g.nodes[ntype].data.items():
            key = f"{ntype}/{name}"
            if key not in node_feats:
                continue
            assert torch.equal(node_feats[key][local], val[src_rows]), f"partition {partid}: node feature {key} differs"
    inner_e = part_g.edata[EID][part_g.edata["inner_edge"].bool()]
    tids, per_type = gpb.map_to_per_etype(inner_e)
    for t, cet in enumerate(g.canonical_etypes):
        ids = per_type[tids == t]
        ekey = ":".join(cet)
        src_rows = ids if orig_eids is None else torch.as_tensor(orig_eids[ekey] if ekey in orig_eids else orig_eids[cet])[ids]
        local = gpb.eid2localeid(ids, gpb.partid, cet)
        for name, val in g.edges[cet].data.items():
            key = f"{ekey}/{name}"
            if key not in edge_feats:
                continue
            assert torch.equal(edge_feats[key][local], val[src_rows]), f"partition {partid}: edge feature {key} differs"


def get_node_partids(partitions_dir, graph_schema):
    """Per node type, the partition id of every node as assigned by the partitioner's ``<ntype>.txt`` files (one id per
    line, node order); the type's node count comes from the chunked-graph metadata."""
    out = {}
    counts = graph_schema.get("num_nodes_per_type") or graph_schema.get("num_nodes_per_chunk")
    for i, ntype in enumerate(graph_schema["node_type"]):
        ids = np.loadtxt(os.path.join(partitions_dir, f"{ntype}.txt"), dtype=np.int64, ndmin=1)
        if counts is not None:
            want = int(np.sum(counts[i]))
            assert ids.shape[0] == want, f"{ntype}.txt lists {ids.shape[0]} nodes, the graph has {want}"
        out[ntype] = torch.from_numpy(ids)
    return out


def verify_node_partitionids(node_partids, part_g, g, gpb, orig_nids=None, partid=None):
    """Every inner node of a partition was assigned to that partition by the partitioner's assignment files."""
    inner_n = part_g.ndata[NID][part_g.ndata["inner_node"].bool()]
    tids, per_type = gpb.map_to_per_ntype(inner_n)
    for t, ntype in enumerate(g.ntypes):
        ids = per_type[tids == t]
        src = ids if orig_nids is None else torch.as_tensor(orig_nids[ntype])[ids]
        got = node_partids[ntype][src]
        assert bool((got == gpb.partid).all()), f"partition {partid}: {int((got != gpb.partid).sum())} nodes of type " \
                                               f"{ntype} were assigned elsewhere by the partitioner"


def read_orig_ids(out_dir, fname, num_parts):
    """Concatenate the per-partition ``orig_nids`` / ``orig_eids`` dictionaries the pipeline writes next to each partition
    (``<out_dir>/part<i>/<fname>``, tensor dictionaries keyed by type) into one mapping per type."""
    from dgl_b200.distributed.partition import _load_feat_file
    merged = {}
    for p in range(num_parts):
        path = os.path.join(out_dir, f"part{p}", fname)
        if not os.path.exists(path):
            continue
        for k, v in _load_feat_file(path).items():
            merged.setdefault(k, []).append(torch.as_tensor(v))
    return {k: torch.cat(v) for k, v in merged.items()}


def load_part_schema(part_config):
    with open(part_config) as f:
        return json.load(f)
