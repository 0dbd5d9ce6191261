"""Cluster-GCN sampler (parity: /root/reference/python/dgl/dataloading/cluster_gcn.py)."""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from ..base import DGLError
from .base import Sampler, set_edge_lazy_features, set_node_lazy_features


class ClusterGCNSampler(Sampler):
    def __init__(self, g, k, cache_path="cluster_gcn.pkl", balance_ntypes=None, balance_edges=False, mode="k-way",
                 prefetch_ndata=None, prefetch_edata=None, output_device=None):
        super().__init__()
        if os.path.exists(cache_path):
            try:
                with open(cache_path, "rb") as f:
                    self.partition_offset, self.partition_node_ids = pickle.load(f)
            except (EOFError, TypeError, ValueError):
                raise DGLError(f"The contents in the cache file {cache_path} is invalid. Please remove the cache "
                               "file or specify another path.")
            if len(self.partition_offset) != k + 1:
                raise DGLError(f"Number of partitions in the cache does not match the value of k. "
                               f"Please remove the cache file {cache_path} or specify another path.")
            if len(self.partition_node_ids) != g.num_nodes():
                raise DGLError(f"Number of nodes in the cache does not match the given graph. "
                               f"Please remove the cache file {cache_path} or specify another path.")
        else:
            from ..partition import metis_partition_assignment
            partition_ids = metis_partition_assignment(g, k, balance_ntypes=balance_ntypes, balance_edges=balance_edges,
                                                       mode=mode)
            partition_ids = partition_ids.cpu().numpy()
            partition_node_ids = np.argsort(partition_ids, kind="stable")
            partition_size = torch.from_numpy(np.bincount(partition_ids, minlength=k))
            partition_offset = torch.zeros(k + 1, dtype=torch.int64)
            partition_offset[1:] = torch.cumsum(partition_size, 0)
            self.partition_offset = partition_offset
            self.partition_node_ids = torch.from_numpy(partition_node_ids)
            with open(cache_path, "wb") as f:
                pickle.dump((self.partition_offset, self.partition_node_ids), f)
        self.prefetch_ndata = prefetch_ndata or []
        self.prefetch_edata = prefetch_edata or []
        self.output_device = output_device

    def sample(self, g, partition_ids):
        node_ids = torch.cat([self.partition_node_ids[self.partition_offset[i]: self.partition_offset[i + 1]]
                              for i in (partition_ids.tolist() if torch.is_tensor(partition_ids) else partition_ids)], 0)
        sg = g.subgraph(node_ids.to(g.device), relabel_nodes=True, output_device=self.output_device)
        set_node_lazy_features(sg, self.prefetch_ndata)
        set_edge_lazy_features(sg, self.prefetch_edata)
        return sg
