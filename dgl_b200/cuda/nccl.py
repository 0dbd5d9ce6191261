"""Sparse all-to-all push/pull for sharded tensors (parity: /root/reference/python/dgl/cuda/nccl.py:7-189).
Default path: peer-memory gather / scatter-add kernels over NVLink (``comm=P2PComm``). The NCCL path (three
all_to_all_single calls, as in the reference) is kept as the baseline."""
from __future__ import annotations

import torch
import torch.distributed as dist


def sparse_all_to_all_push(idx, value, partition):
    """Send (idx, value) rows to their owners; returns the (local idx, values) this rank received."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return partition.map_to_local(idx), value
    world = dist.get_world_size()
    perm, send_counts = partition.generate_permutation(idx)
    idx_s, val_s = partition.map_to_local(idx[perm]), value[perm]
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    idx_r = idx_s.new_empty(sum(rc))
    val_r = val_s.new_empty((sum(rc),) + tuple(val_s.shape[1:]))
    dist.all_to_all_single(idx_r, idx_s, rc, sc)
    dist.all_to_all_single(val_r, val_s, rc, sc)
    return idx_r, val_r


def sparse_all_to_all_pull(req_idx, value, partition):
    """Fetch rows ``req_idx`` of the sharded tensor whose local shard is ``value``."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value[partition.map_to_local(req_idx)]
    perm, send_counts = partition.generate_permutation(req_idx)
    idx_s = partition.map_to_local(req_idx[perm])
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    idx_r = idx_s.new_empty(sum(rc))
    dist.all_to_all_single(idx_r, idx_s, rc, sc)
    resp = value[idx_r]
    out_s = resp.new_empty((sum(sc),) + tuple(resp.shape[1:]))
    dist.all_to_all_single(out_s, resp, sc, rc)
    out = torch.empty_like(out_s)
    out[perm] = out_s
    return out
