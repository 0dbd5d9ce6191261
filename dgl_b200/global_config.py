"""Run-time toggles (parity: /root/reference/python/dgl/global_config.py, docs/source/env_var.rst). libxsmm does not
exist in this build (CPU kernels are OpenMP); the switch is kept so scripts run unchanged. New knobs: kernel path and
communication path selectors."""
import os

_cfg = {"use_libxsmm": False, "deterministic": bool(int(os.environ.get("USE_DETERMINISTIC_ALG", "0"))),
        "comm_path": os.environ.get("DGLB200_COMM", "p2p"), "gemm_path": os.environ.get("DGLB200_GEMM", "auto")}


def use_libxsmm(flag):
    _cfg["use_libxsmm"] = bool(flag)


def is_libxsmm_enabled():
    return _cfg["use_libxsmm"]


def set_comm_path(path):
    """'p2p' (peer-memory fused kernels) or 'nccl' (baseline collectives)."""
    assert path in ("p2p", "nccl")
    _cfg["comm_path"] = path


def get_comm_path():
    return _cfg["comm_path"]


def set_gemm_path(path):
    """'auto' | 'tcgen05' (hand-written sm_100a kernel) | 'cublas'."""
    assert path in ("auto", "tcgen05", "cublas")
    _cfg["gemm_path"] = path


def get_gemm_path():
    return _cfg["gemm_path"]


def set_deterministic(flag):
    """Bit-reproducible sparse reductions (reference env ``USE_DETERMINISTIC_ALG``, src/array/cuda/spmm.cu:33-35):
    g-SpMM stops splitting long rows across work items (no floating-point atomics in the forward aggregation)."""
    _cfg["deterministic"] = bool(flag)


def is_deterministic():
    return _cfg["deterministic"]
