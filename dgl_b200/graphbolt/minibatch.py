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
                blocks.append(b)
            self._blocks = blocks
        return self._blocks

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
