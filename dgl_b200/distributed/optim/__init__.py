"""Distributed sparse optimizers for ``DistEmbedding`` (parity: python/dgl/distributed/optim/pytorch/sparse_optim.py)."""
from ..dist_graph import DistSparseAdagrad as SparseAdagrad  # noqa: F401
from ..dist_graph import DistSparseAdam as SparseAdam  # noqa: F401
from ..dist_graph import DistSparseGradOptimizer  # noqa: F401


def initializer(shape, dtype):
    """Zero state tensor for a sparse optimizer (reference distributed/optim/pytorch/sparse_optim.py:451)."""
    import torch
    return torch.zeros(shape, dtype=dtype)


# ---- exchange helpers of distributed/optim/pytorch/utils.py --------------------------------------------------------
def alltoall_cpu(rank, world_size, output_tensor_list, input_tensor_list):
    """Equal-shape all-to-all on a CPU backend (gloo has no all_to_all): ``world_size`` scatters, rank i being the source
    of round i."""
    import torch.distributed as dist
    inputs = [t.cpu() for t in input_tensor_list]
    for src in range(world_size):
        dist.scatter(output_tensor_list[src], inputs if src == rank else [], src=src)


def alltoallv_cpu(rank, world_size, output_tensor_list, input_tensor_list):
    """Variable-shape all-to-all on a CPU backend: shapes are exchanged first (fixed-size header), then payloads move
    with non-blocking sends and blocking receives; ``output_tensor_list`` entries are replaced by the received tensors."""
    import torch
    import torch.distributed as dist
    MAXD = 8
    inputs = [t.cpu().contiguous() for t in input_tensor_list]
    heads = []
    for t in inputs:
        h = torch.full((MAXD + 1,), -1, dtype=torch.int64)
        h[0] = t.dim()
        h[1:1 + t.dim()] = torch.tensor(list(t.shape), dtype=torch.int64)
        heads.append(h)
    got = [torch.empty(MAXD + 1, dtype=torch.int64) for _ in range(world_size)]
    alltoall_cpu(rank, world_size, got, heads)
    pending = [dist.isend(inputs[dst], dst=dst) for dst in range(world_size) if dst != rank]
    for src in range(world_size):
        if src == rank:
            output_tensor_list[src] = inputs[src]
            continue
        nd = int(got[src][0])
        buf = torch.empty(tuple(got[src][1:1 + nd].tolist()), dtype=inputs[rank].dtype)
        dist.recv(buf, src=src)
        output_tensor_list[src] = buf
    for req in pending:
        req.wait()


def alltoall(rank, world_size, output_tensor_list, input_tensor_list):
    """Equal-shape all-to-all: NCCL's native collective on GPUs, the scatter fallback elsewhere."""
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        dist.all_to_all(output_tensor_list, input_tensor_list)
    else:
        alltoall_cpu(rank, world_size, output_tensor_list, input_tensor_list)


def alltoallv(rank, world_size, output_tensor_list, input_tensor_list):
    """Variable-shape all-to-all (NCCL handles uneven splits natively)."""
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        dist.all_to_all(output_tensor_list, input_tensor_list)
    else:
        alltoallv_cpu(rank, world_size, output_tensor_list, input_tensor_list)
