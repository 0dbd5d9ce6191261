"""ctypes loader for the two native libraries.

Calling convention of every exported symbol: ``int f(args...)`` returning 0 on
success; on failure the message is fetched with ``dglb200_last_error()``.
Argument kinds are described by a compact signature string:

    p = pointer (``tensor.data_ptr()`` or 0)   l = int64   i = int32
    f = float   d = double                      s = const char*

CUDA entry points take the stream handle (``torch.cuda.current_stream()
.cuda_stream``) as their last ``p`` argument so they enqueue on PyTorch's
current stream and are CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes
import os
import threading
from pathlib import Path

from . import build as _build
from ..utils import trace as _trace

# DGLB200_LIBDIR points the loader at an alternative build of the native libraries (e.g. the sanitizer build
# produced by script/run_sanitizers.sh --python)
_LIBDIR = Path(os.environ["DGLB200_LIBDIR"]) if os.environ.get("DGLB200_LIBDIR") else Path(__file__).resolve().parent.parent / "lib"
_KIND = {"p": ctypes.c_void_p, "l": ctypes.c_int64, "i": ctypes.c_int32,
         "f": ctypes.c_float, "d": ctypes.c_double, "s": ctypes.c_char_p}

_lock = threading.Lock()
_cpu = None
_cuda = None
_cuda_err = None


class NativeError(RuntimeError):
    pass


class _Lib:
    def __init__(self, path):
        self.path = str(path)
        self.cdll = ctypes.CDLL(self.path, mode=ctypes.RTLD_GLOBAL if False else ctypes.DEFAULT_MODE)
        self._cache = {}
        self._err = self.cdll.dglb200_last_error
        self._err.restype = ctypes.c_char_p
        self._err.argtypes = []

    def fn(self, name, sig):
        key = (name, sig)
        f = self._cache.get(key)
        if f is None:
            raw = getattr(self.cdll, name)
            raw.restype = ctypes.c_int32
            raw.argtypes = [_KIND[c] for c in sig]
            err = self._err

            is_cuda = "cuda" in self.path
            tstate = _trace._state

            def f(*args, _raw=raw, _name=name):
                rc = _trace.call(_name, _raw, args, is_cuda) if tstate["on"] else _raw(*args)
                if rc != 0:
                    msg = err()
                    raise NativeError(f"{_name}: {msg.decode() if msg else rc}")
                return rc

            self._cache[key] = f
        return f

    def raw(self, name, restype, argtypes):
        f = getattr(self.cdll, name)
        f.restype = restype
        f.argtypes = argtypes
        return f

    def has(self, name):
        return hasattr(self.cdll, name)


def _need_build(path: Path, kind: str = "") -> bool:
    """Missing library, or (cheap mtime test) a source newer than it; the build itself is incremental."""
    if not path.exists():
        return True
    if not kind:
        return False
    src = Path(__file__).resolve().parent.parent / "csrc"
    try:
        built = path.stat().st_mtime
        files = list((src / kind).glob("*")) + list(src.glob("*.h"))
        return any(f.stat().st_mtime > built for f in files)
    except OSError:
        return False


def cpu_lib() -> _Lib:
    global _cpu
    if _cpu is None:
        with _lock:
            if _cpu is None:
                p = _LIBDIR / "libdglb200_cpu.so"
                if _need_build(p, "cpu") or os.environ.get("DGLB200_REBUILD"):
                    _build.build_cpu(verbose=bool(os.environ.get("DGLB200_VERBOSE")))
                    try:
                        os.utime(p)  # nothing changed (hash-cached): stop the mtime test from firing again
                    except OSError:
                        pass
                _cpu = _Lib(p)
                from ..utils import cpu_quota
                cpu_quota.cap_native(_cpu)  # OpenMP teams sized by the cgroup CPU quota, not by the visible core count
                _guard_forked_children(_cpu)
    return _cpu


_FORK_GUARD = []


def _guard_forked_children(lib):
    """GNU OpenMP does not survive ``fork()``: a child inherits the parent's thread-pool bookkeeping without the
    threads, and its first parallel region with more than one thread waits for them forever. Every forked child
    (DataLoader workers, ``multiprocessing`` pools, user code) therefore starts with one-thread native teams - those
    run inline and never touch the pool. A child that wants parallel native kernels again calls
    ``dgl_b200.utils.cpu_quota.apply(force=True)`` after it has exec'ed or re-initialised OpenMP."""
    if _FORK_GUARD or not hasattr(os, "register_at_fork"):
        return
    try:
        setter = lib.fn("dglb200_cpu_set_num_threads", "i")
    except Exception:  # noqa: BLE001 - library without the setter: nothing to pin
        return
    _FORK_GUARD.append(setter)
    os.register_at_fork(after_in_child=lambda: setter(1))


def cuda_lib() -> _Lib:
    """Load the sm_100a kernel library. Fails loudly: there is no eager/PyTorch
    fallback for CUDA tensors when this library is absent."""
    global _cuda, _cuda_err
    if _cuda is None:
        with _lock:
            if _cuda is None:
                p = _LIBDIR / "libdglb200_cuda.so"
                if _need_build(p) or os.environ.get("DGLB200_REBUILD"):
                    _build.build_cuda(verbose=bool(os.environ.get("DGLB200_VERBOSE")))
                _cuda = _Lib(p)
    return _cuda


def cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """data pointer of a tensor or 0 for None."""
    return 0 if t is None else t.data_ptr()
