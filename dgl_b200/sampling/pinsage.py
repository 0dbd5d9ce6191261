"""PinSAGE-style random-walk neighbor samplers (parity: /root/reference/python/dgl/sampling/pinsage.py;
visit counting: on the GPU one CTA per seed counts in a shared-memory hash table and extracts the top-k in place
(csrc/cuda/visit_count.cu, counterpart of src/graph/sampling/randomwalks/frequency_hashmap.cu); on the host a sort +
run-length pass)."""
from __future__ import annotations

import torch

from ..base import EID
from ..convert import graph as make_graph
from ..convert import heterograph as make_heterograph
from ..heterograph import _to_tensor
from .randomwalks import random_walk


class RandomWalkNeighborSampler:
    """For each seed run ``num_random_walks`` walks with restarts along ``metapath`` and keep the
    ``num_neighbors`` most visited end nodes; returns a graph seed <- neighbor with edge weights = visit counts."""

    def __init__(self, G, num_traversals, termination_prob, num_random_walks, num_neighbors, metapath=None,
                 weight_column="weights"):
        self.G = G
        self.weight_column = weight_column
        self.num_random_walks = num_random_walks
        self.num_neighbors = num_neighbors
        self.num_traversals = num_traversals
        if metapath is None:
            if len(G.ntypes) > 1 or len(G.etypes) > 1:
                raise ValueError("Metapath must be specified if the graph is homogeneous.")
            metapath = [G.canonical_etypes[0]]
        start_ntype = G.to_canonical_etype(metapath[0])[0]
        end_ntype = G.to_canonical_etype(metapath[-1])[-1]
        if start_ntype != end_ntype:
            raise ValueError("The metapath must start and end at the same node type.")
        self.ntype = start_ntype
        self.metapath_hops = len(metapath)
        self.metapath = metapath
        self.full_metapath = metapath * num_traversals
        restart_prob = torch.zeros(self.metapath_hops * num_traversals)
        restart_prob[self.metapath_hops:: self.metapath_hops] = termination_prob
        self.restart_prob = restart_prob
        self.device_counter = True  # GPU graphs: count visits in the shared-memory kernel (False = sort / run-length)

    def __call__(self, seed_nodes):
        G = self.G
        seed_nodes = _to_tensor(seed_nodes, G.idtype, G.device)
        rep = seed_nodes.repeat_interleave(self.num_random_walks)
        paths, _ = random_walk(G, rep, metapath=self.full_metapath, restart_prob=self.restart_prob.to(G.device))
        ends = paths[:, self.metapath_hops:: self.metapath_hops]
        n = G.num_nodes(self.ntype)
        M = self.num_random_walks * ends.shape[1]
        if ends.is_cuda and self.device_counter and 0 < M <= 8192 and n < (1 << 40):
            return self._device_topk(seed_nodes, ends.reshape(seed_nodes.shape[0], M).contiguous(), n)
        src = ends.reshape(-1)
        dst = rep.unsqueeze(1).expand_as(ends).reshape(-1)
        ok = src >= 0
        src, dst = src[ok], dst[ok]
        n = G.num_nodes(self.ntype)
        key = dst.long() * n + src.long()
        uniq, cnt = torch.unique(key, return_counts=True)
        d, s = uniq // n, uniq % n
        # top-k by count per dst
        order = torch.sort(-cnt, stable=True)[1]
        order = order[torch.sort(d[order], stable=True)[1]]
        dd = d[order]
        starts = torch.searchsorted(dd, dd)
        rank = torch.arange(dd.shape[0], device=dd.device) - starts
        keep = order[rank < self.num_neighbors]
        g = make_graph((s[keep].to(G.idtype), d[keep].to(G.idtype)), num_nodes=n, idtype=G.idtype, device=G.device)
        g.edata[self.weight_column] = cnt[keep]
        return g


    def _device_topk(self, seed_nodes, visits, n):
        """Top ``num_neighbors`` visited nodes per seed, counted and selected inside one kernel."""
        from .._native import loader
        from ..base import id_code
        G = self.G
        S, M = visits.shape
        K = int(self.num_neighbors)
        out_node = torch.empty((S, K), dtype=visits.dtype, device=visits.device)
        out_cnt = torch.empty((S, K), dtype=torch.int32, device=visits.device)
        loader.cuda_lib().fn("dglb200_cuda_visit_topk", "ipliippp")(
            id_code(visits.dtype), visits.data_ptr(), S, M, K, out_node.data_ptr(), out_cnt.data_ptr(), loader.stream_ptr())
        keep = out_cnt.reshape(-1) > 0
        src = out_node.reshape(-1)[keep]
        dst = seed_nodes.unsqueeze(1).expand(S, K).reshape(-1)[keep]
        g = make_graph((src.to(G.idtype), dst.to(G.idtype)), num_nodes=n, idtype=G.idtype, device=G.device)
        g.edata[self.weight_column] = out_cnt.reshape(-1)[keep].long()
        return g


class PinSAGESampler(RandomWalkNeighborSampler):
    """Item-item sampler on a bipartite user-item graph: walks item -> user -> item."""

    def __init__(self, G, ntype, other_type, num_traversals, termination_prob, num_random_walks, num_neighbors,
                 weight_column="weights"):
        metagraph = G.canonical_etypes
        fw = [c for c in metagraph if c[0] == ntype and c[2] == other_type][0]
        bw = [c for c in metagraph if c[0] == other_type and c[2] == ntype][0]
        super().__init__(G, num_traversals, termination_prob, num_random_walks, num_neighbors, metapath=[fw, bw],
                         weight_column=weight_column)
