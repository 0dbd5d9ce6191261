"""Graph partitioning for distributed training: partition artifacts + partition book.

Parity: /root/reference/python/dgl/distributed/partition.py (partition_graph :817-1400, load_partition :286,
load_partition_book, inner_node/inner_edge masks :37-45) and graph_partition_book.py (RangePartitionBook).
Artifacts: ``<out>/<name>.json`` + ``<out>/part<i>/{graph.dgl,node_feat.dgl,edge_feat.dgl}``. Node (and edge) ids are
reshuffled so every partition owns one contiguous global-id range per type; ``orig_id`` keeps the mapping."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from ..base import EID, NID, DGLError
from ..convert import graph as make_graph
from ..data import dgl_binary as _binary
from ..partition import metis_partition_assignment

RESERVED_FIELD_DTYPE = {"inner_node": torch.uint8, "inner_edge": torch.uint8}


class RangePartitionBook:
    """Global id -> (partition, local id) for range-partitioned node / edge id spaces."""

    def __init__(self, part_id, num_parts, node_map, edge_map, ntypes, etypes):
        self._partid = part_id
        self._num_partitions = num_parts
        self._ntypes = dict(ntypes)
        self._etypes = dict(etypes)
        # maps: type -> [num_parts, 2] ranges
        self._node_map = {k: torch.as_tensor(v).long() for k, v in node_map.items()}
        self._edge_map = {k: torch.as_tensor(v).long() for k, v in edge_map.items()}

    @property
    def partid(self):
        return self._partid

    def num_partitions(self):
        return self._num_partitions

    @property
    def ntypes(self):
        return list(self._ntypes.keys())

    @property
    def etypes(self):
        return list(self._etypes.keys())

    canonical_etypes = etypes

    # ---- reference graph_partition_book.py RangePartitionBook extras ------------------------------------------
    @property
    def global_nid_dtype(self):
        return torch.int64

    @property
    def global_eid_dtype(self):
        return torch.int64

    def is_homogeneous(self):
        return len(self._ntypes) == 1 and len(self._etypes) == 1

    def to_canonical_etype(self, etype):
        return etype

    def shared_memory(self, graph_name):
        """Publish the book's range tables in POSIX shared memory under ``graph_name`` (server -> trainers)."""
        from ..utils.shared_mem import tensors_to_shared_memory
        tensors = {f"node~{k}": v for k, v in self._node_map.items()}
        tensors.update({f"edge~{k}": v for k, v in self._edge_map.items()})
        return tensors_to_shared_memory(graph_name + "_gpb", tensors,
                                        meta={"part_id": self._partid, "num_parts": self._num_partitions,
                                              "ntypes": self._ntypes, "etypes": self._etypes})

    def _homo_offsets(self, maps, order):
        """Start of every type's block in the homogeneous id space of every partition: ids are laid out partition by
        partition, type by type inside a partition (the layout ``partition_graph`` writes)."""
        types = sorted(order, key=order.get)
        P = self._num_partitions
        sizes = torch.stack([maps[t][:, 1] - maps[t][:, 0] for t in types], 1)            # [P, T]
        flat = sizes.reshape(-1)
        starts = (torch.cumsum(flat, 0) - flat).reshape(P, len(types))
        return types, sizes, starts

    def _to_homo(self, ids, typ, maps, order):
        types, sizes, starts = self._homo_offsets(maps, order)
        t = types.index(typ)
        ids = torch.as_tensor(ids).long()
        rng = maps[typ]
        part = torch.searchsorted(rng[:, 1].contiguous(), ids, right=True)
        return starts[part, t] + (ids - rng[part, 0])

    def _to_per_type(self, ids, maps, order):
        types, sizes, starts = self._homo_offsets(maps, order)
        ids = torch.as_tensor(ids).long()
        ends = (starts + sizes).reshape(-1)
        slot = torch.searchsorted(ends.contiguous(), ids, right=True)
        p, t = slot // len(types), slot % len(types)
        base = torch.stack([maps[x][:, 0] for x in types], 1)                              # [P, T]
        return t, base[p, t] + (ids - starts.reshape(-1)[slot])

    def map_to_homo_nid(self, ids, ntype):
        return self._to_homo(ids, ntype, self._node_map, self._ntypes)

    def map_to_homo_eid(self, ids, etype):
        return self._to_homo(ids, etype, self._edge_map, self._etypes)

    def map_to_per_ntype(self, ids):
        return self._to_per_type(ids, self._node_map, self._ntypes)

    def map_to_per_etype(self, ids):
        return self._to_per_type(ids, self._edge_map, self._etypes)

    @property
    def local_ntype_offset(self):
        types, sizes, _ = self._homo_offsets(self._node_map, self._ntypes)
        return [0] + torch.cumsum(sizes[self._partid], 0).tolist()

    @property
    def local_etype_offset(self):
        types, sizes, _ = self._homo_offsets(self._edge_map, self._etypes)
        return [0] + torch.cumsum(sizes[self._partid], 0).tolist()

    def _num(self, m):
        return int((m[:, 1] - m[:, 0]).sum())

    def num_nodes(self, ntype="_N"):
        return self._num(self._node_map[ntype])

    def num_edges(self, etype="_E"):
        return self._num(self._edge_map[self._ekey(etype)])

    def _ekey(self, etype):
        if etype in self._edge_map:
            return etype
        if isinstance(etype, tuple) and ":".join(etype) in self._edge_map:  # canonical tuple for "src:rel:dst" keys
            return ":".join(etype)
        for k in self._edge_map:
            if isinstance(k, tuple) and k[1] == etype or (isinstance(k, str) and k.split(":")[1:2] == [etype]):
                return k
        raise DGLError(f"unknown edge type {etype}")

    def metadata(self):
        return [{"machine_id": i, "num_nodes": {k: int(v[i, 1] - v[i, 0]) for k, v in self._node_map.items()},
                 "num_edges": {str(k): int(v[i, 1] - v[i, 0]) for k, v in self._edge_map.items()}}
                for i in range(self._num_partitions)]

    def nid2partid(self, nids, ntype="_N"):
        ends = self._node_map[ntype][:, 1].contiguous()
        return torch.searchsorted(ends.to(nids.device), nids.long(), right=True)

    def eid2partid(self, eids, etype="_E"):
        ends = self._edge_map[self._ekey(etype)][:, 1].contiguous()
        return torch.searchsorted(ends.to(eids.device), eids.long(), right=True)

    def partid2nids(self, partid, ntype="_N"):
        lo, hi = self._node_map[ntype][partid].tolist()
        return torch.arange(lo, hi)

    def partid2eids(self, partid, etype="_E"):
        lo, hi = self._edge_map[self._ekey(etype)][partid].tolist()
        return torch.arange(lo, hi)

    def nid2localnid(self, nids, partid, ntype="_N"):
        return nids - int(self._node_map[ntype][partid, 0])

    def eid2localeid(self, eids, partid, etype="_E"):
        return eids - int(self._edge_map[self._ekey(etype)][partid, 0])

    def to_dict(self):
        return {"node_map": {k: v.tolist() for k, v in self._node_map.items()},
                "edge_map": {str(k): v.tolist() for k, v in self._edge_map.items()}}


GraphPartitionBook = RangePartitionBook


def partition_graph(g, graph_name, num_parts, out_path, num_hops=1, part_method="metis", balance_ntypes=None,
                    balance_edges=False, return_mapping=False, num_trainers_per_machine=1, objtype="cut",
                    graph_formats=None, use_graphbolt=False, **kwargs):
    """Partition a (hetero)graph into ``num_parts`` parts with ``num_hops`` of halo nodes and save the artifacts.

    Heterographs follow the reference flow (python/dgl/distributed/partition.py:817-1400): the graph is homogenised with
    its NTYPE / ETYPE recorded, the homogeneous nodes are assigned to partitions, and ids are reshuffled so that every
    partition owns one contiguous range PER TYPE (the layout ``RangePartitionBook`` decodes: partition by partition, type
    by type inside a partition). Features are written under type-prefixed keys (``ntype/name``, ``src:rel:dst/name``).
    A partition holds its inner nodes, their ``num_hops``-hop in-neighbourhood as halo nodes, and exactly the in-edges of
    the inner nodes and of the halo nodes closer than ``num_hops`` (no halo-to-outermost-halo edges)."""
    from ..base import ETYPE, NTYPE
    hetero = not g.is_homogeneous
    g = g.cpu() if g.device.type != "cpu" else g
    if hetero:
        from ..convert import to_homogeneous
        hg = to_homogeneous(g)
        ntype_of, etype_of = hg.ndata[NTYPE].long(), hg.edata[ETYPE].long()
        type_nid, type_eid = hg.ndata[NID].long(), hg.edata[EID].long()       # per-type original ids
        ntypes, cets = list(g.ntypes), list(g.canonical_etypes)
    else:
        hg = g
        ntype_of = torch.zeros(g.num_nodes(), dtype=torch.int64)
        etype_of = torch.zeros(g.num_edges(), dtype=torch.int64)
        type_nid, type_eid = torch.arange(g.num_nodes()), torch.arange(g.num_edges())
        ntypes, cets = ["_N"], [("_N", "_E", "_N")]
    T, R = len(ntypes), len(cets)
    ekeys = [":".join(c) for c in cets]
    n = hg.num_nodes()
    if kwargs.get("node_part") is not None:  # precomputed assignment (tools/dispatch_data.py, ParMETIS output)
        node_part = torch.as_tensor(kwargs["node_part"]).long()
        if node_part.shape[0] != n or int(node_part.max()) >= num_parts:
            raise DGLError("node_part must assign every node to a partition id < num_parts")
        part_method = "given"
    elif num_parts == 1:
        node_part = torch.zeros(n, dtype=torch.int64)
    elif part_method == "metis":
        bal = balance_ntypes if balance_ntypes is not None or not hetero else None
        node_part = metis_partition_assignment(hg, num_parts, balance_ntypes=bal, balance_edges=balance_edges)
    elif part_method == "random":
        node_part = torch.randint(0, num_parts, (n,))
    else:
        raise DGLError("Unknown partitioning method: " + part_method)
    # ---- reshuffle: homogeneous ids ordered by (partition, type); edges by (partition of dst, type)
    order = torch.sort(node_part * T + ntype_of, stable=True)[1]
    new_nid = torch.empty_like(order)
    new_nid[order] = torch.arange(n)
    ncnt = torch.bincount(node_part * T + ntype_of, minlength=num_parts * T).reshape(num_parts, T)
    pstart = torch.cumsum(ncnt.sum(1), 0) - ncnt.sum(1)                        # homogeneous start of every partition
    u, v = hg.all_edges()
    nu, nv = new_nid[u.long()], new_nid[v.long()]
    epart = node_part[v.long()]                                                # an edge lives where its destination lives
    eorder = torch.sort(epart * R + etype_of, stable=True)[1]
    new_eid = torch.empty_like(eorder)
    new_eid[eorder] = torch.arange(eorder.shape[0])
    ecnt = torch.bincount(epart * R + etype_of, minlength=num_parts * R).reshape(num_parts, R)
    epstart = torch.cumsum(ecnt.sum(1), 0) - ecnt.sum(1)

    def per_type_ranges(cnt):
        start = torch.cumsum(cnt, 0) - cnt                                     # [P, types]: start inside each type's id space
        return [[[int(start[p, t]), int(start[p, t] + cnt[p, t])] for p in range(num_parts)] for t in range(cnt.shape[1])]

    node_map = dict(zip(ntypes, per_type_ranges(ncnt)))
    edge_map = dict(zip(ekeys, per_type_ranges(ecnt)))
    os.makedirs(out_path, exist_ok=True)
    meta = {"graph_name": graph_name, "num_nodes": n, "num_edges": hg.num_edges(), "part_method": part_method,
            "num_parts": num_parts, "halo_hops": num_hops, "node_map": node_map, "edge_map": edge_map,
            "ntypes": {t: i for i, t in enumerate(ntypes)}, "etypes": {k: i for i, k in enumerate(ekeys)}}
    part_of_new = node_part[order]  # partition of every new node id (monotone)
    ntype_new, etype_new_by_orig = ntype_of[order], etype_of
    gn = make_graph((nu, nv), num_nodes=n)
    for p in range(num_parts):
        inner = torch.arange(int(pstart[p]), int(pstart[p] + ncnt[p].sum()))
        seen = torch.zeros(n, dtype=torch.bool)
        seen[inner] = True
        dst_ok = seen.clone()                                                  # nodes whose in-edges are stored
        frontier, halo = inner, []
        for hop in range(num_hops):
            src, _ = gn.in_edges(frontier)
            new = torch.unique(src)
            new = new[~seen[new]]
            seen[new] = True
            halo.append(new)
            if hop < num_hops - 1:
                dst_ok[new] = True
            frontier = new
            if new.numel() == 0:
                break
        nodes = torch.cat([inner] + halo)  # inner nodes first => local ids of inner nodes = global - start
        local = torch.full((n,), -1, dtype=torch.int64)
        local[nodes] = torch.arange(nodes.shape[0])
        keep = dst_ok[nv] & seen[nu]
        inner_e = keep & (part_of_new[nv] == p)
        ke = torch.nonzero(keep).squeeze(1)
        # inner edges first, ordered by new edge id
        is_inner = inner_e[ke]
        key = (~is_inner).long() * (int(new_eid.max()) + 1 if new_eid.numel() else 1) + new_eid[ke]
        ke = ke[torch.argsort(key)]
        pg = make_graph((local[nu[ke]], local[nv[ke]]), num_nodes=nodes.shape[0])
        ndata = {NID: nodes, "inner_node": (part_of_new[nodes] == p).to(torch.uint8), "part_id": part_of_new[nodes],
                 "orig_id": order[nodes], NTYPE: ntype_new[nodes]}
        edata = {EID: new_eid[ke], "inner_edge": inner_e[ke].to(torch.uint8), "orig_id": ke, ETYPE: etype_new_by_orig[ke]}
        pdir = os.path.join(out_path, f"part{p}")
        os.makedirs(pdir, exist_ok=True)
        for k, v in ndata.items():
            pg.ndata[k] = v
        for k, v in edata.items():
            pg.edata[k] = v
        # the reference's artifact names and binary containers: partitions written by either framework load in the other,
        # and loading memory-maps the files instead of reading them
        _binary.write_graphs(os.path.join(pdir, "graph.dgl"), [pg])
        nfeat, efeat = {}, {}
        off = int(pstart[p])
        for t, nt in enumerate(ntypes):                                        # inner nodes of type t, in new-id order
            ids = type_nid[order[off: off + int(ncnt[p, t])]]
            off += int(ncnt[p, t])
            frame = g.nodes[nt].data if hetero else g.ndata
            for k, val in frame.items():
                nfeat[f"{nt}/{k}"] = val[ids]
        off = int(epstart[p])
        for r, cet in enumerate(cets):
            ids = type_eid[eorder[off: off + int(ecnt[p, r])]]
            off += int(ecnt[p, r])
            frame = g.edges[cet].data if hetero else g.edata
            for k, val in frame.items():
                efeat[f"{ekeys[r]}/{k}"] = val[ids]
        _binary.write_tensors(os.path.join(pdir, "node_feat.dgl"), nfeat)
        _binary.write_tensors(os.path.join(pdir, "edge_feat.dgl"), efeat)
        meta[f"part-{p}"] = {"node_feats": f"part{p}/node_feat.dgl", "edge_feats": f"part{p}/edge_feat.dgl",
                             "part_graph": f"part{p}/graph.dgl"}
    cfg = os.path.join(out_path, f"{graph_name}.json")
    with open(cfg, "w") as f:
        json.dump(meta, f, indent=2)
    if return_mapping:
        if not hetero:
            return order, eorder
        # per type: original per-type id of every new per-type id
        nmap = {nt: type_nid[order[ntype_of[order] == t]] for t, nt in enumerate(ntypes)}
        emap = {cet: type_eid[eorder[etype_of[eorder] == r]] for r, cet in enumerate(cets)}
        return nmap, emap
    return None


def _load_part_graph(path):
    """Partition structure file: the reference's binary container (graph.dgl) or the round-1 tensor archive (graph.pt)."""
    if _binary.is_dgl_binary(path):
        return _binary.read_graphs(path)[0][0]
    d = torch.load(path, weights_only=True)
    g = make_graph((d["src"], d["dst"]), num_nodes=d["num_nodes"])
    for k, v in d["ndata"].items():
        g.ndata[k] = v
    for k, v in d["edata"].items():
        g.edata[k] = v
    return g


def _load_feat_file(path):
    if _binary.is_dgl_binary(path):
        return _binary.read_tensors(path)
    return torch.load(path, weights_only=True)


def load_partition_book(part_config, part_id):
    with open(part_config) as f:
        meta = json.load(f)
    gpb = RangePartitionBook(part_id, meta["num_parts"], meta["node_map"], meta["edge_map"], meta["ntypes"], meta["etypes"])
    return gpb, meta["graph_name"], list(meta["ntypes"].keys()), list(meta["etypes"].keys())


def load_partition(part_config, part_id, load_feats=True, use_graphbolt=False):
    """Returns (graph, node_feats, edge_feats, gpb, graph_name, ntypes, etypes)."""
    with open(part_config) as f:
        meta = json.load(f)
    base = os.path.dirname(part_config)
    info = meta[f"part-{part_id}"]
    g = _load_part_graph(os.path.join(base, info["part_graph"]))
    nf = _load_feat_file(os.path.join(base, info["node_feats"])) if load_feats else {}
    ef = _load_feat_file(os.path.join(base, info["edge_feats"])) if load_feats else {}
    gpb, name, ntypes, etypes = load_partition_book(part_config, part_id)
    return g, nf, ef, gpb, name, ntypes, etypes


def load_partition_feats(part_config, part_id, load_nodes=True, load_edges=True):
    with open(part_config) as f:
        meta = json.load(f)
    base = os.path.dirname(part_config)
    info = meta[f"part-{part_id}"]
    nf = _load_feat_file(os.path.join(base, info["node_feats"])) if load_nodes else None
    ef = _load_feat_file(os.path.join(base, info["edge_feats"])) if load_edges else None
    return nf, ef


def gb_convert_single_dgl_partition(part_id, graph_formats, part_config, store_eids=True, store_inner_node=False,
                                    store_inner_edge=False):
    """Convert ONE partition written by ``partition_graph`` into a GraphBolt ``FusedCSCSamplingGraph`` saved next to it as
    ``fused_csc_sampling_graph.pt`` (reference distributed/partition.py:1844). Returns the path."""
    from ..base import EID, NID
    from ..graphbolt import fused_csc_sampling_graph as make_fused
    g, _, _, gpb, _, ntypes, etypes = load_partition(part_config, part_id, load_feats=False)
    indptr, indices, eids = g._graphs[0].csc()
    node_attrs, edge_attrs = {}, {}
    if NID in g.ndata:
        node_attrs[NID] = g.ndata[NID]
    if store_inner_node and "inner_node" in g.ndata:
        node_attrs["inner_node"] = g.ndata["inner_node"]
    perm = eids.long() if eids is not None else None
    if store_eids and EID in g.edata:
        edge_attrs[EID] = g.edata[EID] if perm is None else g.edata[EID][perm]
    if store_inner_edge and "inner_edge" in g.edata:
        edge_attrs["inner_edge"] = g.edata["inner_edge"] if perm is None else g.edata["inner_edge"][perm]
    fg = make_fused(indptr.long(), indices.long(), node_attributes=node_attrs or None, edge_attributes=edge_attrs or None)
    with open(part_config) as f:
        meta = json.load(f)
    base = os.path.dirname(part_config)
    rel = os.path.join(os.path.dirname(meta[f"part-{part_id}"]["part_graph"]), "fused_csc_sampling_graph.pt")
    path = os.path.join(base, rel)
    torch.save(fg, path)
    return rel


def dgl_partition_to_graphbolt(part_config, *, store_eids=True, store_inner_node=False, store_inner_edge=False,
                               graph_formats=None, n_jobs=1):
    """Convert every partition of ``part_config`` to GraphBolt and record the files in the config
    (``part_graph_graphbolt``), as the reference does (distributed/partition.py:1965)."""
    with open(part_config) as f:
        meta = json.load(f)
    for pid in range(meta["num_parts"]):
        rel = gb_convert_single_dgl_partition(pid, graph_formats, part_config, store_eids, store_inner_node, store_inner_edge)
        meta[f"part-{pid}"]["part_graph_graphbolt"] = rel
    with open(part_config, "w") as f:
        json.dump(meta, f, sort_keys=True, indent=4)


# ---- small public helpers of the partition pipeline (reference partition.py:91-107, :1588-1596, :1623-1720) -------------
def _narrowest_int(max_value, tensor):
    """``tensor`` in the smallest integer dtype that can index ``max_value`` items."""
    for dt in (torch.int32, torch.int64):
        if max_value <= torch.iinfo(dt).max:
            return tensor.to(dt)
    return tensor


def process_partitions(g, formats=None, sort_etypes=False):
    """Normalise a partition graph before it is written: reserved columns in their canonical dtypes and, if requested,
    CSR / CSC neighbour lists grouped by edge type (what the typed samplers of the servers expect)."""
    from ..base import ETYPE
    for name, dt in {**RESERVED_FIELD_DTYPE, "_TYPE": torch.int32}.items():
        for frame in (g.ndata, g.edata):
            if name in frame:
                frame[name] = frame[name].to(dt)
    if sort_etypes and formats:
        from ..transforms import sort_csc_by_tag, sort_csr_by_tag
        if "csr" in formats:
            g = sort_csr_by_tag(g, tag=g.edata[ETYPE], tag_type="edge")
        if "csc" in formats:
            g = sort_csc_by_tag(g, tag=g.edata[ETYPE], tag_type="edge")
    return g


def is_homogeneous(ntypes, etypes):
    return len(ntypes) == 1 and len(etypes) == 1


def init_type_per_edge(graph, gpb):
    """edge-type id of every edge of a partition graph, recovered from its global edge ids."""
    return gpb.map_to_per_etype(graph.edata[EID])[0]


def cast_various_to_minimum_dtype_gb(num_parts, indptr, indices, type_per_edge, etypes, ntypes, node_attributes,
                                     edge_attributes, part_meta=None, graph=None, edge_count=None, node_count=None,
                                     tot_edge_count=None, tot_node_count=None):
    """Shrink the integer arrays of a GraphBolt partition to the narrowest dtype that holds their value range
    (indptr: total edge count; indices / node ids: total node count; edge ids: total edge count; type ids: number of
    types; inner flags: uint8). Returns the same six objects in the same order."""
    if graph is not None:
        if part_meta is None:
            raise DGLError("part_meta is needed when the counts are taken from a graph")
        tot_edge_count, tot_node_count = graph.num_edges(), graph.num_nodes()
        node_count, edge_count = part_meta["num_nodes"], part_meta["num_edges"]
    else:                                            # counts not given: take them from the arrays themselves
        tot_edge_count = int(indptr[-1]) if tot_edge_count is None else tot_edge_count
        tot_node_count = int(indptr.shape[0]) - 1 if tot_node_count is None else tot_node_count
        edge_count = tot_edge_count if edge_count is None else edge_count
        node_count = tot_node_count if node_count is None else node_count
    indptr = _narrowest_int(tot_edge_count, indptr)
    indices = _narrowest_int(tot_node_count, indices)
    if type_per_edge is not None:
        type_per_edge = type_per_edge.to(torch.uint8 if len(etypes) <= 255 else torch.int16 if len(etypes) < 2 ** 15 else torch.int32)
    limits = {NID: node_count, EID: edge_count, "_TYPE": max(len(ntypes), len(etypes))}
    for attrs in (node_attributes, edge_attributes):
        for key in list(attrs or {}):
            if key in limits:
                attrs[key] = _narrowest_int(limits[key], attrs[key])
            elif key in RESERVED_FIELD_DTYPE:
                attrs[key] = attrs[key].to(RESERVED_FIELD_DTYPE[key])
    return indptr, indices, type_per_edge, node_attributes, edge_attributes
