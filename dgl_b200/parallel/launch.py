"""Process-group bootstrap helpers (torchrun env: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def local_device():
    if torch.cuda.is_available():
        lr = int(os.environ.get("LOCAL_RANK", "0"))
        return torch.device("cuda", lr % torch.cuda.device_count())
    return torch.device("cpu")


def init_process_group_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = local_device()
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if dev.type == "cuda" else "gloo"
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = dev
        dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
    return rank, world, dev
