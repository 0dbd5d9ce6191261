"""MiniBatch and SampledSubgraph containers (parity: /root/reference/python/dgl/graphbolt/minibatch.py,
sampled_subgraph.py, impl/sampled_subgraph_impl.py)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Union

import torch

from .base import CSCFormatBase, etype_str_to_tuple


@dataclass
class SampledSubgraphImpl:
    sampled_csc: Union[CSCFormatBase, Dict[str, CSCFormatBase]] = None
    original_column_node_ids: Union[torch.Tensor, Dict[str, torch.Tensor]] = None
    original_row_node_ids: Union[torch.Tensor, Dict[str, torch.Tensor]] = None
    original_edge_ids: Union[torch.Tensor, Dict[str, torch.Tensor]] = None

    def to(self, device, non_blocking=False):
        def mv(x):
            if x is None:
                return None
            if isinstance(x, dict):
                return {k: mv(v) for k, v in x.items()}
            return x.to(device, non_blocking=non_blocking)
        return SampledSubgraphImpl(mv(self.sampled_csc), mv(self.original_column_node_ids), mv(self.original_row_node_ids),
                                   mv(self.original_edge_ids))

    def _tensors(self):
        for x in (self.sampled_csc, self.original_column_node_ids, self.original_row_node_ids, self.original_edge_ids):
            for v in (x.values() if isinstance(x, dict) else [x]):
                if v is None:
                    continue
                if hasattr(v, "indptr"):
                    yield v.indptr
                    yield v.indices
                else:
                    yield v

    def pin_memory(self):
        def pin(x):
            if x is None:
                return None
            if isinstance(x, dict):
                return {k: pin(v) for k, v in x.items()}
            if hasattr(x, "indptr"):
                return CSCFormatBase(indptr=x.indptr.pin_memory(), indices=x.indices.pin_memory())
            return x.pin_memory()
        return SampledSubgraphImpl(pin(self.sampled_csc), pin(self.original_column_node_ids), pin(self.original_row_node_ids),
                                   pin(self.original_edge_ids))

    def is_pinned(self):
        return all(t.is_pinned() for t in self._tensors())

    def to_pyg(self, x):
        """This layer in PyG convention (reference sampled_subgraph.py:286): ``x`` is the source-side feature tensor (or
        ``{ntype: tensor}`` for heterogeneous layers). Returns :class:`PyGLayerData` /
        :class:`PyGLayerHeteroData` with x = (source features, destination features = first ``num_dst`` rows),
        ``edge_index`` [2, E] and ``size`` = (num src, num dst). Works without torch_geometric installed."""
        from .base import etype_str_to_tuple
        from .compat import PyGLayerData, PyGLayerHeteroData

        def one(csc):
            deg = (csc.indptr[1:] - csc.indptr[:-1]).long()
            dst = torch.repeat_interleave(torch.arange(deg.shape[0], device=deg.device), deg)
            return torch.stack([csc.indices.long(), dst]), int(deg.shape[0])

        if not isinstance(self.sampled_csc, dict):
            edge_index, n_dst = one(self.sampled_csc)
            return PyGLayerData((x, x[:n_dst]), edge_index, (int(x.shape[0]), n_dst))
        x_dst, edge_index, sizes = {}, {}, {}
        for etype, csc in self.sampled_csc.items():
            ei, n_dst = one(csc)
            s_t, _, d_t = etype_str_to_tuple(etype)
            x_dst[d_t] = x[d_t][:n_dst]
            edge_index[etype] = ei
            sizes[etype] = (int(x[s_t].shape[0]), n_dst)
        return PyGLayerHeteroData((x, x_dst), edge_index, sizes)

    def exclude_edges(self, edges, assume_num_node_within_int32=True):
        """Remove the given (src, dst) pairs (original ids) from the sampled block."""
        from .utils import exclude_edges_from_subgraph
        return exclude_edges_from_subgraph(self, edges)


SampledSubgraph = SampledSubgraphImpl


@dataclass
class MiniBatch:
    """One training minibatch flowing through the datapipe."""
    labels: Union[torch.Tensor, Dict[str, torch.Tensor]] = None
    seeds: Union[torch.Tensor, Dict[str, torch.Tensor]] = None
    indexes: Union[torch.Tensor, Dict[str, torch.Tensor]] = None
    sampled_subgraphs: List[SampledSubgraphImpl] = None
    input_nodes: Union[torch.Tensor, Dict[str, torch.Tensor]] = None
    node_features: Union[Dict[str, torch.Tensor], Dict] = None
    edge_features: List = None
    compacted_seeds: Union[torch.Tensor, Dict[str, torch.Tensor]] = None
    _blocks: list = field(default=None, repr=False)

    def __repr__(self):
        parts = []
        for k in ("seeds", "labels", "input_nodes", "indexes", "compacted_seeds"):
            v = getattr(self, k)
            if v is not None:
                parts.append(f"{k}={v}")
        if self.sampled_subgraphs is not None:
            parts.append(f"sampled_subgraphs={len(self.sampled_subgraphs)} layers")
        if self.node_features is not None:
            parts.append(f"node_features={list(self.node_features.keys())}")
        return "MiniBatch(" + ",\n          ".join(parts) + ")"

    def node_ids(self):
        return self.input_nodes

    def num_layers(self):
        return 0 if self.sampled_subgraphs is None else len(self.sampled_subgraphs)

    def edge_ids(self, layer_id):
        return self.sampled_subgraphs[layer_id].original_edge_ids

    def set_node_features(self, node_features):
        self.node_features = node_features

    def set_edge_features(self, edge_features):
        self.edge_features = edge_features

    @property
    def blocks(self):
        """DGLBlocks built straight from the sampled CSC (no re-compaction)."""
        if not self.sampled_subgraphs:
            return None
        if self._blocks is None:
            from ..convert import create_block
            from ..base import NID, EID
            blocks = []
            for sg in self.sampled_subgraphs:
                ready = getattr(sg, "_block", None)
                if ready is not None:                         # built by the one-sync sampler together with the subgraph
                    blocks.append(ready)
                    continue
                rn, cn = sg.original_row_node_ids, sg.original_column_node_ids
                if isinstance(sg.sampled_csc, dict):
                    data, nsrc, ndst = {}, {}, {}
                    for et, csc in sg.sampled_csc.items():
                        s, e, d = etype_str_to_tuple(et)
                        data[(s, e, d)] = ("csc", (csc.indptr, csc.indices, []))
                    for t, v in rn.items():
                        nsrc[t] = int(v.shape[0])
                    for t, v in cn.items():
                        ndst[t] = int(v.shape[0])
                    for (s, e, d) in list(data.keys()):
                        nsrc.setdefault(s, 0); ndst.setdefault(d, 0)
                    b = create_block(data, nsrc, ndst)
                    for t, v in rn.items():
                        if t in b.srctypes:
                            b.srcnodes[t].data[NID] = v
                    for t, v in cn.items():
                        if t in b.dsttypes:
                            b.dstnodes[t].data[NID] = v
                    if sg.original_edge_ids is not None:
                        for et, v in sg.original_edge_ids.items():
                            b.edges[etype_str_to_tuple(et)].data[EID] = v
                else:
                    csc = sg.sampled_csc
                    b = create_block(("csc", (csc.indptr, csc.indices, [])), num_src_nodes=int(rn.shape[0]),
                                     num_dst_nodes=int(cn.shape[0]))
                    b.srcdata[NID] = rn
                    b.dstdata[NID] = cn
                    if sg.original_edge_ids is not None:
                        b.edata[EID] = sg.original_edge_ids
                    ell = getattr(sg, "_ell_cache", None)
                    if ell is not None:                       # sampled as an ELL block: keep that form for the fused layer
                        b._ell_cache = ell
                blocks.append(b)
            self._blocks = blocks
        return self._blocks

    def compute_blocks(self):
        """Build (and cache) the DGL blocks of this mini-batch (reference minibatch.py compute_blocks)."""
        return self.blocks

    def is_pinned(self):
        def ok(x):
            if x is None:
                return True
            if isinstance(x, dict):
                return all(ok(v) for v in x.values())
            if isinstance(x, (list, tuple)):
                return all(ok(v) for v in x)
            if torch.is_tensor(x):
                return x.is_pinned()
            if hasattr(x, "is_pinned"):
                return x.is_pinned()
            return True
        return all(ok(getattr(self, k)) for k in ("labels", "seeds", "indexes", "input_nodes", "node_features",
                                                  "edge_features", "compacted_seeds", "sampled_subgraphs"))

    def to_pyg_data(self):
        """PyG ``Data`` of the mini-batch (reference minibatch.py to_pyg_data): needs torch_geometric; the sampled layers
        are merged into the first layer's edge_index (homogeneous graphs)."""
        try:
            from torch_geometric.data import Data
        except ImportError as e:  # pragma: no cover - optional dependency
            raise ImportError("to_pyg_data requires torch_geometric") from e
        sg = self.sampled_subgraphs[0]
        feat = next(iter(self.node_features.values())) if self.node_features else None
        _, edge_index, _ = sg.to_pyg(feat if feat is not None else sg.original_row_node_ids)
        return Data(x=feat, edge_index=edge_index, y=self.labels, batch_size=None if self.seeds is None else len(self.seeds),
                    n_id=self.node_ids())

    def to(self, device, non_blocking=False):
        def mv(x):
            if x is None:
                return None
            if isinstance(x, dict):
                return {k: mv(v) for k, v in x.items()}
            if isinstance(x, (list, tuple)):
                return type(x)(mv(v) for v in x)
            if hasattr(x, "to"):
                return x.to(device, non_blocking=non_blocking) if torch.is_tensor(x) or hasattr(x, "sampled_csc") else x.to(device)
            return x
        out = MiniBatch()
        for k in ("labels", "seeds", "indexes", "sampled_subgraphs", "input_nodes", "node_features", "edge_features",
                  "compacted_seeds"):
            setattr(out, k, mv(getattr(self, k)))
        return out

    def pin_memory(self):
        def pin(x):
            if x is None:
                return None
            if isinstance(x, dict):
                return {k: pin(v) for k, v in x.items()}
            if isinstance(x, (list, tuple)):
                return type(x)(pin(v) for v in x)
            if torch.is_tensor(x):
                return x.pin_memory()
            return x
        out = MiniBatch()
        for k in ("labels", "seeds", "indexes", "input_nodes", "node_features", "edge_features", "compacted_seeds"):
            setattr(out, k, pin(getattr(self, k)))
        out.sampled_subgraphs = self.sampled_subgraphs
        return out
