"""Multi-GPU runtime: one process per GPU, torch.distributed (NCCL) for bootstrap only; data movement on the
hot paths goes through peer-mapped symmetric memory and hand-written kernels (csrc/cuda/comm.cu)."""
from .p2p import HaloPlan, P2PComm, SymmetricBuffer  # noqa: F401
from .launch import init_process_group_from_env, local_device  # noqa: F401


def __getattr__(name):
    if name in ("HaloAggregate", "PartitionedSAGE"):
        from . import partitioned
        return getattr(partitioned, name)
    raise AttributeError(name)
