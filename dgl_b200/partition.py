"""Graph partitioning utilities (parity: /root/reference/python/dgl/partition.py: metis_partition_assignment :404,
metis_partition :517, partition_graph_with_halo :185, NDArrayPartition :474, reshuffle_graph)."""
from __future__ import annotations

import torch

from . import random as dgl_random
from ._native import loader
from .base import EID, NID, DGLError, id_code
from .convert import graph as make_graph
from .subgraph import node_subgraph


def reorder_nodes(g, new_node_ids):
    """Graph whose node ``v`` of ``g`` is renamed ``new_node_ids[v]`` (a permutation of 0..N-1); edge order is kept and
    ``ndata['orig_id']`` maps new ids back (reference partition.py:22)."""
    new_node_ids = torch.as_tensor(new_node_ids, device=g.device).to(g.idtype)
    if new_node_ids.shape[0] != g.num_nodes():
        raise DGLError("The number of new node ids must match #nodes in the graph.")
    sorted_ids, idx = torch.sort(new_node_ids)
    if g.num_nodes() and not torch.equal(sorted_ids, torch.arange(g.num_nodes(), dtype=g.idtype, device=g.device)):
        raise DGLError("The new node IDs are incorrect.")
    u, v = g.edges()
    out = make_graph((new_node_ids[u.long()], new_node_ids[v.long()]), num_nodes=g.num_nodes(), idtype=g.idtype,
                     device=g.device)
    out.ndata["orig_id"] = idx.to(g.idtype)
    return out


def get_peak_mem():
    """Peak virtual memory of this process in GB (``VmPeak`` of /proc/self/status; 0.0 where unavailable)."""
    try:
        with open("/proc/self/status") as f:
            for line in f:
                if line.startswith("VmPeak"):
                    return int(line.split()[1]) / 1024 / 1024
    except OSError:
        pass
    return 0.0


def _sym_csr(g):
    """CSR of the symmetrised homogeneous view of g (CPU int64)."""
    from .convert import to_homogeneous
    hg = g if g.is_homogeneous else to_homogeneous(g)
    u, v = hg.all_edges()
    u, v = u.cpu().long(), v.cpu().long()
    n = hg.num_nodes()
    row = torch.cat([u, v])
    col = torch.cat([v, u])
    from .graph_index import coo_to_csr
    indptr, indices, _ = coo_to_csr(row, col, n, torch.int64)
    return hg, indptr.contiguous(), indices.contiguous(), n


def metis_partition_assignment(g, k, balance_ntypes=None, balance_edges=False, mode="k-way", objtype="cut",
                               refine_iters=8, imbalance=1.03, algorithm="multilevel"):
    """Assign every node to one of ``k`` parts minimising the edge cut under a balance constraint.
    The reference calls METIS; this build ships its own partitioners (csrc/cpu/partition.cc): ``algorithm="multilevel"``
    (default: heavy-edge-matching coarsening, LDG on the coarsest graph, k-way boundary refinement on the way back) or
    ``"ldg"`` (one streaming pass + label propagation: faster, weaker on graphs with community structure).
    Returns an int64 tensor of part ids."""
    if algorithm not in ("multilevel", "ldg"):
        raise DGLError("algorithm must be 'multilevel' or 'ldg'")
    if mode not in ("k-way", "recursive"):
        raise DGLError("mode must be 'k-way' or 'recursive'")
    hg, indptr, indices, n = _sym_csr(g)
    if k == 1:
        return torch.zeros(n, dtype=torch.int64)
    vwgt = None
    if balance_edges:
        vwgt = (indptr[1:] - indptr[:-1]).clamp(min=1).contiguous()
    part = torch.empty(n, dtype=torch.int64)
    loader.cpu_lib().fn("dglb200_cpu_ml_partition" if algorithm == "multilevel" else "dglb200_cpu_ldg_partition", "ipplipidlp")(
        id_code(torch.int64), indptr.data_ptr(), indices.data_ptr(), n, int(k), 0 if vwgt is None else vwgt.data_ptr(),
        int(refine_iters), float(imbalance), dgl_random.next_seed(), part.data_ptr())
    if balance_ntypes is not None:
        # rebalance per node type: move surplus nodes of over-represented types to the lightest parts
        nt = torch.as_tensor(balance_ntypes).long().cpu()
        for t in torch.unique(nt).tolist():
            idx = torch.nonzero(nt == t).squeeze(1)
            target = -(-idx.numel() // k)
            cnt = torch.bincount(part[idx], minlength=k)
            for p in torch.nonzero(cnt > target * 1.1).squeeze(1).tolist():
                extra = idx[part[idx] == p][target:]
                for v in extra.tolist():
                    q = int(torch.argmin(cnt))
                    if cnt[q] >= target:
                        break
                    part[v] = q
                    cnt[q] += 1
                    cnt[p] -= 1
    return part


def partition_graph_with_halo(g, node_part, extra_cached_hops, reshuffle=False):
    """Split ``g`` by ``node_part`` keeping ``extra_cached_hops`` hops of HALO nodes; every partition is a
    subgraph with ``ndata['inner_node']``, ``edata['inner_edge']``, original ids in NID/EID
    (reference partition.py:185-300)."""
    node_part = torch.as_tensor(node_part).long().to(g.device)
    k = int(node_part.max()) + 1
    u, v = g.all_edges()
    parts = {}
    orig_nids = orig_eids = None
    if reshuffle:
        order = torch.sort(node_part, stable=True)[1]
        new_id = torch.empty_like(order)
        new_id[order] = torch.arange(order.shape[0], device=order.device)
        g2 = make_graph((new_id[u.long()].to(g.idtype), new_id[v.long()].to(g.idtype)), num_nodes=g.num_nodes(),
                        idtype=g.idtype, device=g.device)
        g2.ndata["orig_id"] = order.to(g.idtype)
        g2.edata["orig_id"] = torch.arange(g.num_edges(), dtype=g.idtype, device=g.device)
        node_part = node_part[order]
        orig_nids, orig_eids = g2.ndata["orig_id"], g2.edata["orig_id"]
        g = g2
        u, v = g.all_edges()
    for p in range(k):
        inner = torch.nonzero(node_part == p).squeeze(1)
        frontier = inner
        allnodes = [inner]
        seen = torch.zeros(g.num_nodes(), dtype=torch.bool, device=g.device)
        seen[inner] = True
        for _ in range(extra_cached_hops):
            src, _ = g.in_edges(frontier.to(g.idtype))
            new = torch.unique(src.long())
            new = new[~seen[new]]
            seen[new] = True
            allnodes.append(new)
            frontier = new
            if new.numel() == 0:
                break
        nodes = torch.cat(allnodes)
        sub = node_subgraph(g, nodes.to(g.idtype))
        # keep only edges whose destination is an inner node or within the halo
        sub.ndata["inner_node"] = (node_part[sub.ndata[NID].long()] == p).to(torch.int32)
        su, sv = sub.all_edges()
        sub.edata["inner_edge"] = sub.ndata["inner_node"][sv.long()].clone()
        sub.ndata["part_id"] = node_part[sub.ndata[NID].long()]
        if reshuffle:
            sub.ndata["orig_id"] = orig_nids[sub.ndata[NID].long()]
            sub.edata["orig_id"] = orig_eids[sub.edata[EID].long()]
        parts[p] = sub
    if reshuffle:
        return parts, orig_nids, orig_eids
    return parts, None, None


def metis_partition(g, k, extra_cached_hops=0, reshuffle=False, balance_ntypes=None, balance_edges=False,
                    mode="k-way"):
    node_part = metis_partition_assignment(g, k, balance_ntypes, balance_edges, mode)
    parts, _, _ = partition_graph_with_halo(g, node_part, extra_cached_hops, reshuffle)
    return parts


def reshuffle_graph(g, node_part=None):
    """Relabel nodes so each partition owns a contiguous id range (reference partition.py:reshuffle_graph)."""
    if node_part is None:
        return g
    node_part = torch.as_tensor(node_part).long()
    order = torch.sort(node_part, stable=True)[1]
    new_id = torch.empty_like(order)
    new_id[order] = torch.arange(order.shape[0])
    u, v = g.all_edges()
    g2 = make_graph((new_id[u.long()].to(g.idtype), new_id[v.long()].to(g.idtype)), num_nodes=g.num_nodes(),
                    idtype=g.idtype, device=g.device)
    g2.ndata["orig_id"] = order.to(g.idtype)
    g2.edata["orig_id"] = torch.arange(g.num_edges(), dtype=g.idtype)
    for key, val in g.ndata.items():
        g2.ndata[key] = val[order]
    for key, val in g.edata.items():
        g2.edata[key] = val
    return g2


def libra_partition_assignment(g, k):
    """Vertex-cut (edge) partition: returns the part id of every edge (reference distgnn/partition/libra_partition.py)."""
    u, v = g.all_edges()
    u, v = u.cpu().long().contiguous(), v.cpu().long().contiguous()
    out = torch.empty(u.shape[0], dtype=torch.int64)
    loader.cpu_lib().fn("dglb200_cpu_libra_partition", "ippllip")(id_code(torch.int64), u.data_ptr(), v.data_ptr(),
                                                                  u.shape[0], g.num_nodes(), int(k), out.data_ptr())
    return out


class NDArrayPartition:
    """Partition of an index space across ``num_parts`` workers: 'remainder' (id % P) or 'range'
    (reference partition.py:474-560, src/partition/ndarray_partition.cc, cuda/partition_op.cu)."""

    def __init__(self, array_size, num_parts, mode="remainder", part_ranges=None):
        assert num_parts > 0, 'Invalid "num_parts", must be > 0.'
        self._array_size, self._num_parts, self._mode = array_size, num_parts, mode
        if mode == "remainder":
            assert part_ranges is None, "When using remainder-based partitioning, \"part_ranges\" should not be specified."
        elif mode == "range":
            if part_ranges is None:
                per = -(-array_size // num_parts)
                part_ranges = torch.arange(0, num_parts + 1) * per
                part_ranges[-1] = array_size
            part_ranges = torch.as_tensor(part_ranges).long()
            assert part_ranges[0] == 0 and part_ranges[-1] == array_size, 'part_ranges[0] must be 0, and part_ranges[-1] must be "array_size".'
            self._ranges = part_ranges
        else:
            raise DGLError(f"Unknown partition mode {mode}")

    def get(self):
        """The partition object itself (the reference returns its C++ handle here)."""
        return self

    def num_parts(self):
        return self._num_parts

    def array_size(self):
        return self._array_size

    def get_local_indices(self, part, ctx=None):
        if self._mode == "remainder":
            idx = torch.arange(part, self._array_size, self._num_parts)
        else:
            idx = torch.arange(int(self._ranges[part]), int(self._ranges[part + 1]))
        return idx if ctx is None else idx.to(ctx)

    def local_size(self, part):
        return int(self.get_local_indices(part).shape[0])

    def owner(self, idxs):
        if self._mode == "remainder":
            return idxs % self._num_parts
        return torch.searchsorted(self._ranges.to(idxs.device), idxs, right=True) - 1

    def map_to_local(self, idxs):
        if self._mode == "remainder":
            return idxs // self._num_parts
        return idxs - self._ranges.to(idxs.device)[self.owner(idxs)]

    def map_to_global(self, idxs, part_id):
        if self._mode == "remainder":
            return idxs * self._num_parts + part_id
        return idxs + int(self._ranges[part_id])

    def generate_permutation(self, idxs):
        """Permutation grouping ``idxs`` by owner + per-part counts (basis of sparse all-to-all).
        CUDA inputs: owner/count kernel + radix sort on ceil(log2 P) bits (csrc/cuda/graph_kernels.cu)."""
        if idxs.is_cuda and idxs.dtype in (torch.int32, torch.int64):
            from ._graph_kernels import partition_owner, partition_perm
            own, counts, _ = partition_owner(idxs.contiguous(), self._num_parts,
                                             self._ranges if self._mode == "range" else None)
            return partition_perm(own, self._num_parts), counts
        own = self.owner(idxs)
        perm = torch.sort(own, stable=True)[1]
        counts = torch.bincount(own, minlength=self._num_parts)
        return perm, counts
