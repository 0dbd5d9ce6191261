"""All-to-all helpers on the gloo backend (reference tools/distpartitioning/gloo_wrapper.py). gloo has no all_to_all for
uneven splits, so sizes are exchanged with an all_gather and the payloads with pairwise isend / irecv."""
import json

import torch
import torch.distributed as dist


def allgather_sizes(send_sizes, world):
    """send_sizes[r] = rows this rank sends to r  ->  [world, world] matrix (row = sender)."""
    t = torch.tensor(send_sizes, dtype=torch.int64)
    out = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(out, t)
    return torch.stack(out)


def alltoallv_cpu(rank, world, tensors):
    """tensors[r] goes to rank r (any first-dim length, identical trailing dims and dtype on all ranks); returns the list
    of tensors received from every rank (index = sender)."""
    assert len(tensors) == world
    sizes = allgather_sizes([int(t.shape[0]) for t in tensors], world)
    tail, dtype = tuple(tensors[rank].shape[1:]), tensors[rank].dtype
    recv = [torch.empty((int(sizes[s, rank]),) + tail, dtype=dtype) for s in range(world)]
    recv[rank] = tensors[rank].clone()
    works = []
    for peer in range(world):
        if peer == rank:
            continue
        if tensors[peer].shape[0]:
            works.append(dist.isend(tensors[peer].contiguous(), peer))
        if recv[peer].shape[0]:
            works.append(dist.irecv(recv[peer], peer))
    for w in works:
        w.wait()
    return recv


def alltoall_cpu(rank, world, output_list, input_list):
    res = alltoallv_cpu(rank, world, input_list)
    for o, r in zip(output_list, res):
        o.copy_(r)


def allgather_ints(value, world):
    out = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(out, torch.tensor([int(value)], dtype=torch.int64))
    return [int(x) for x in out]


def gather_metadata_json(obj, rank, world):
    """Gather one JSON-serialisable object per rank on rank 0 (list indexed by rank; None elsewhere)."""
    out = [None] * world
    dist.all_gather_object(out, json.dumps(obj))
    return [json.loads(x) for x in out] if rank == 0 else None
