"""NodeEmbedding: sparse-gradient embedding table, optionally sharded across GPUs.

Parity: /root/reference/python/dgl/nn/pytorch/sparse_emb.py (NodeEmbedding :17, all_get/all_set_embedding :319-386).
Multi-GPU sharding uses the peer-memory table of :class:`dgl_b200.parallel.P2PComm`: lookups are direct remote
row gathers over NVLink and gradient pushes are remote ``red.add`` scatters (csrc/cuda/comm.cu), replacing the
reference's 3x all_to_all_single + blocking stream sync (python/dgl/cuda/nccl.py:72-183)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..base import DGLError

_STORE = {}


class NodeEmbedding:
    def __init__(self, num_embeddings, embedding_dim, name, init_func=None, device=None, partition=None, comm=None):
        self._num_embeddings = num_embeddings
        self._embedding_dim = embedding_dim
        self._name = name
        self._optm_state = None
        self._trace = []
        self._comm = comm
        self._partition = partition
        if dist.is_initialized():
            self._rank, self._world_size = dist.get_rank(), dist.get_world_size()
        else:
            self._rank, self._world_size = -1, 0
        if device is None:
            device = torch.device("cpu")
        self._device = torch.device(device)
        if comm is not None and comm.world_size > 1:
            # range partition: shard r owns rows [r*shard, (r+1)*shard)
            self._shard_rows = -(-num_embeddings // comm.world_size)
            dim_pad = (embedding_dim + 3) // 4 * 4
            self._dim_pad = dim_pad
            self._shards = comm.alloc(self._shard_rows * dim_pad * 4)
            self._tensor = self._shards.view(torch.float32)[: self._shard_rows * dim_pad].view(self._shard_rows, dim_pad)
            lo = comm.rank * self._shard_rows
            n_local = max(0, min(num_embeddings, lo + self._shard_rows) - lo)
            if init_func is not None and n_local > 0:
                self._tensor[:n_local, :embedding_dim] = init_func(
                    torch.empty((n_local, embedding_dim), device=self._tensor.device))
            comm.barrier(3)
        else:
            key = name
            if key not in _STORE:
                emb = torch.empty((num_embeddings, embedding_dim), dtype=torch.float32, device=self._device)
                if self._device.type == "cpu" and self._world_size > 1:
                    emb.share_memory_()
                if init_func is not None:
                    emb = init_func(emb)
                _STORE[key] = emb
            self._tensor = _STORE[key]
            self._shards = None

    def __call__(self, node_ids, device=torch.device("cpu")):
        if self._shards is not None:
            idx = node_ids.to(self._comm.device)
            out = torch.empty((idx.shape[0], self._dim_pad), dtype=torch.float32, device=self._comm.device)
            self._comm.gather(self._shards, self._shard_rows, idx.contiguous(), out)
            emb = out[:, : self._embedding_dim].to(device)
        else:
            emb = self._tensor[node_ids.to(self._tensor.device)].to(device)
        if torch.is_grad_enabled():
            emb.requires_grad_(True)
            emb.retain_grad()
            self._trace.append((node_ids.to(device), emb))
        return emb

    @property
    def store(self):
        return self._tensor

    @property
    def partition(self):
        return self._partition

    @property
    def rank(self):
        return self._rank

    @property
    def world_size(self):
        return self._world_size

    @property
    def name(self):
        return self._name

    @property
    def num_embeddings(self):
        return self._num_embeddings

    @property
    def embedding_dim(self):
        return self._embedding_dim

    def set_optm_state(self, state):
        self._optm_state = state

    @property
    def optm_state(self):
        return self._optm_state

    @property
    def trace(self):
        return self._trace

    def reset_trace(self):
        self._trace = []

    @property
    def weight(self):
        return self._tensor

    def all_set_embedding(self, values):
        if self._shards is not None:
            lo = self._comm.rank * self._shard_rows
            hi = min(self._num_embeddings, lo + self._shard_rows)
            if hi > lo:
                self._tensor[: hi - lo, : self._embedding_dim] = values[lo:hi].to(self._tensor.device)
            self._comm.barrier(3)
        else:
            self._tensor[:] = values.to(self._tensor.device)

    def all_get_embedding(self):
        if self._shards is not None:
            idx = torch.arange(self._num_embeddings, device=self._comm.device)
            out = torch.empty((idx.shape[0], self._dim_pad), dtype=torch.float32, device=self._comm.device)
            self._comm.gather(self._shards, self._shard_rows, idx, out)
            return out[:, : self._embedding_dim].cpu()
        return self._tensor.detach().cpu()
