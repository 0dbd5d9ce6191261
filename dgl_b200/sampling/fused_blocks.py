"""One-synchronisation multi-layer neighbour sampling into message-flow graphs.

The layer-by-layer formulation (``sample_neighbors`` -> ``to_block`` per layer, the reference's
python/dgl/dataloading/neighbor_sampler.py:152-206 over src/array/cuda/rowwise_sampling.cu + cuda_to_block.cu) reads
sizes back to the host after every kernel family: >= 2 host synchronisations per layer before the next layer can even be
sized. Here ALL layers are enqueued back to back on capacity-sized buffers with device-side counts - the same
``sample_ell`` (sub-warp Floyd sampling into a dense ELL block) and cooperative ``frontier_compact`` kernels the fused
training engine replays in its CUDA graph - and the host reads every count ONCE at the end (L + 1 node counts and L edge
counts in one copy). Blocks are then cut out of the capacity buffers as views: CSC arrays by one ``ell_to_csc`` launch per
layer, and the ELL form is attached to the block so ``SAGEConv`` can feed the tcgen05 fused layer without re-deriving it.

Requirements: homogeneous graph on the GPU, uniform sampling without replacement, fan-outs in [1, 64] (or -1 handled by
the caller's fallback), < 2^31 nodes and edges.
"""
from __future__ import annotations

import torch

from .._native import loader
from ..base import EID, NID, id_code
from ..graph_index import UnitGraph
from ..heterograph import DGLBlock
from ..ops.fused import FrontierCompactor


def _p(t):
    return 0 if t is None else t.data_ptr()


def _csc_int32(g):
    """int32 CSC of the graph's single relation (cached on the structure) + its edge-id map (None = identity)."""
    ug = g._graphs[0]
    cache = ug._degs
    hit = cache.get("csc_i32")
    if hit is None:
        indptr, indices, eids = ug.csc()
        hit = (indptr.to(torch.int32).contiguous(), indices.to(torch.int32).contiguous(), eids)
        cache["csc_i32"] = hit
    return hit


ENABLED = [__import__("os").environ.get("DGLB200_FUSED_SAMPLER", "1") != "0"]   # A/B switch


def usable(g, fanouts, prob, replace, edge_dir, exclude_eids):
    if not ENABLED[0]:
        return False
    if not (g.device.type == "cuda" and len(g.canonical_etypes) == 1 and len(g.ntypes) == 1):
        return False
    if prob is not None or replace or edge_dir != "in" or exclude_eids is not None:
        return False
    if not all(isinstance(f, int) and 1 <= f <= 64 for f in fanouts):
        return False
    return g.num_nodes() < 2 ** 31 and g.num_edges() < 2 ** 31


class _Workspace:
    """Capacity buffers of one (batch capacity, fan-outs) shape: worst case ``cap[l-1] = cap[l] * (1 + fanout_l)``."""

    def __init__(self, batch_cap, fanouts, device):
        L = len(fanouts)
        self.batch_cap, self.fanouts, self.L = batch_cap, list(fanouts), L
        caps = [0] * (L + 1)
        caps[L] = batch_cap
        for l in range(L, 0, -1):
            caps[l - 1] = caps[l] * (1 + fanouts[l - 1])
        self.caps = caps
        i32 = torch.int32
        self.seeds = torch.full((batch_cap,), -1, dtype=i32, device=device)
        self.n = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(L + 1)]
        self.nbr_g = [None] + [torch.empty((caps[l], fanouts[l - 1]), dtype=i32, device=device) for l in range(1, L + 1)]
        self.eid = [None] + [torch.empty((caps[l], fanouts[l - 1]), dtype=i32, device=device) for l in range(1, L + 1)]
        self.cnt = [None] + [torch.empty(caps[l], dtype=i32, device=device) for l in range(1, L + 1)]
        self.totals = [None] + [torch.zeros(16, dtype=torch.int64, device=device) for _ in range(1, L + 1)]
        self.compactors = [None] + [FrontierCompactor(caps[l], fanouts[l - 1], caps[l - 1], device) for l in range(1, L + 1)]
        self.counts_host = torch.zeros(2 * L + 1, dtype=torch.int64).pin_memory()


class FusedBlockSampler:
    """Callable ``(g, seeds) -> (input_nodes, output_nodes, blocks)`` for fixed fan-outs; workspaces are cached per batch
    capacity (rounded up to a power of two) and reused across calls."""

    def __init__(self, fanouts, seed=0):
        self.fanouts = [int(f) for f in fanouts]
        self.seed = int(seed)
        self._ws = {}
        self._step = 0

    def _workspace(self, n_seeds, device):
        cap = 64
        while cap < n_seeds:
            cap <<= 1
        key = (cap, device.index)
        ws = self._ws.get(key)
        if ws is None:
            ws = self._ws[key] = _Workspace(cap, self.fanouts, device)
        return ws

    def enqueue(self, ws, indptr, indices, uva=False, step_dev=None, compact_input=True):
        """Enqueue every layer on the workspace's capacity buffers (seeds in ``ws.seeds``, their count in ``ws.n[L]``)
        WITHOUT reading anything back: returns (S, nbr_local) lists indexed by layer. With ``step_dev`` (int64[1] on the
        device) the random stream is keyed by that counter instead of a host value, so the sequence is CUDA-graph
        capturable and still differs from replay to replay. ``compact_input=False`` skips the compaction of the input layer
        (``S[0]`` / ``nbr_local[1]`` stay None; ``ws.nbr_g[1]`` holds global ids)."""
        lib = loader.cuda_lib()
        L = len(self.fanouts)
        S = [None] * (L + 1)
        nbr_local = [None] * (L + 1)
        S[L] = ws.seeds
        it = id_code(torch.int32)
        for l in range(L, 0, -1):
            f = self.fanouts[l - 1]
            ws.totals[l].zero_()
            host_seed = self.seed * 7919 + l + (0 if step_dev is not None else 131 * self._step)
            lib.fn("dglb200_cuda_sample_ell", "ippiplpilpppppi" + "p")(
                it, _p(indptr), _p(indices), 1 if uva else 0, _p(S[l]), ws.caps[l], _p(ws.n[l]), f, host_seed, _p(step_dev),
                _p(ws.nbr_g[l]), _p(ws.eid[l]), _p(ws.cnt[l]), _p(ws.totals[l]), 1, loader.stream_ptr())
            if l == 1 and not compact_input:
                continue          # the input layer will read the feature table by global id: its frontier is never built
            S[l - 1], nbr_local[l], n_prev, _ = ws.compactors[l].compact(S[l], ws.n[l], ws.nbr_g[l])
            ws.n[l - 1] = n_prev
        return S, nbr_local

    def sample_layers(self, indptr, indices, seeds, uva=False):
        """Core pipeline on int32 CSC arrays (device, or host-pinned with ``uva``). Returns per layer l = 1..L a dict with
        ``dst`` / ``src`` (global ids), ``nbr`` (local ELL), ``cnt``, ``eid`` (CSC positions, ELL shaped), ``fanout``,
        ``n_dst`` / ``n_src`` / ``n_edges`` (host ints) - after exactly ONE host synchronisation."""
        dev = seeds.device
        L = len(self.fanouts)
        n_seeds = int(seeds.shape[0])
        ws = self._workspace(n_seeds, dev)
        ws.seeds[:n_seeds] = seeds.to(torch.int32)
        if n_seeds < ws.batch_cap:
            ws.seeds[n_seeds:] = -1
        ws.n[L].fill_(n_seeds)
        self._step += 1
        S, nbr_local = self.enqueue(ws, indptr, indices, uva)
        # ---- ONE host read: node counts of S_L .. S_0 and the edge count of every layer
        packed = torch.cat([ws.n[l].view(1) for l in range(L, -1, -1)] + [ws.totals[l].sum().view(1) for l in range(L, 0, -1)])
        ws.counts_host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        counts = ws.counts_host.tolist()
        n_nodes = {l: int(counts[L - l]) for l in range(L, -1, -1)}
        n_edges = {l: int(counts[L + 1 + (L - l)]) for l in range(L, 0, -1)}
        layers = []
        for l in range(1, L + 1):
            nd, ns = n_nodes[l], n_nodes[l - 1]
            dst = S[l][:nd].clone() if l == L else S[l][:nd]           # the seeds live in the reused workspace
            layers.append(dict(dst=dst, src=S[l - 1][:ns], nbr=nbr_local[l][:nd], cnt=ws.cnt[l][:nd].clone(), eid=ws.eid[l],
                               fanout=self.fanouts[l - 1], n_dst=nd, n_src=ns, n_edges=n_edges[l],
                               n_dst_dev=ws.n[l].clone()))
        return layers

    @staticmethod
    def layer_to_csc(layer, idtype, eid_map=None, want_eids=True):
        """(indptr, indices, eids) of one sampled layer in ``idtype`` (one cumsum + one ``ell_to_csc`` launch)."""
        dev = layer["cnt"].device
        nd, ne = layer["n_dst"], layer["n_edges"]
        ptr64 = torch.zeros(nd + 1, dtype=torch.int64, device=dev)
        torch.cumsum(layer["cnt"], 0, out=ptr64[1:])
        idx = torch.empty(ne, dtype=idtype, device=dev)
        eids = torch.empty(ne, dtype=idtype, device=dev) if want_eids else None
        loader.cuda_lib().fn("dglb200_ell_to_csc", "ipppplipppp")(
            id_code(idtype), _p(layer["nbr"]), _p(layer["eid"]), _p(layer["cnt"]), _p(ptr64), nd, layer["fanout"], _p(idx),
            _p(eids), _p(eid_map), loader.stream_ptr())
        ptr = ptr64.to(idtype)
        from .._sparse_ops import SPLIT_DEG
        if layer["fanout"] <= SPLIT_DEG:
            try:                           # every row has at most `fanout` entries: the balanced g-SpMM needs no work items
                ptr._dglb_split_meta = {"n_long": 0}
            except Exception:  # noqa: BLE001
                pass
        return ptr, idx, eids

    @staticmethod
    def layer_to_block(layer, idtype, canonical_etypes, ntypes, eid_map=None, want_eids=True):
        """DGLBlock of one sampled layer, built without any validation pass (sizes and ids are known to be consistent),
        carrying its ELL form and the device-side destination count for the fused SAGE layer."""
        ptr, idx, eids = FusedBlockSampler.layer_to_csc(layer, idtype, eid_map, want_eids)
        ug = UnitGraph.from_csc(layer["n_src"], layer["n_dst"], ptr, idx)
        block = DGLBlock([ug], None, list(canonical_etypes), [layer["n_src"], layer["n_dst"]], srctypes=list(ntypes),
                         dsttypes=list(ntypes))
        block.srcdata[NID] = layer["src"].to(idtype)
        block.dstdata[NID] = layer["dst"].to(idtype)
        if eids is not None:
            block.edata[EID] = eids
        block._ell_cache = (layer["nbr"], layer["cnt"], layer["fanout"])
        block._n_dst_dev = layer["n_dst_dev"]
        return block, ptr, idx, eids

    def __call__(self, g, seeds):
        indptr, indices, csc_eids = _csc_int32(g)
        blocks = [self.layer_to_block(layer, g.idtype, g.canonical_etypes, g.ntypes, csc_eids)[0]
                  for layer in self.sample_layers(indptr, indices, seeds)]
        return blocks[0].srcdata[NID], blocks[-1].dstdata[NID], blocks
