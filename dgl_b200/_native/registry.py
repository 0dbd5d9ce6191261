"""Global function registry: the user-visible half of the reference's packed-function FFI
(/root/reference/python/dgl/_ffi/function.py:164-305 ``register_func`` / ``get_global_func`` / ``list_global_func_names`` /
``extract_ext_funcs``; type codes /root/reference/python/dgl/_ffi/runtime_ctypes.py:20-45).

The framework's native code is a flat C ABI (``int dglb200_<name>(...)``, see loader.py) rather than type-erased packed
functions, so a "global function" is one of two things here:

* a Python callable registered under a name with :func:`register_func` (extension hook, same decorator form), or
* an exported native entry point, looked up by its symbol name in the host library and then the CUDA library and
  returned as a ``ctypes`` function with ``restype=int`` (the caller sets ``argtypes``; ``loader._Lib.fn`` is the typed
  route used internally).

``list_global_func_names`` enumerates both by reading the dynamic symbol tables of the two shared objects directly
(ELF64 ``.dynsym``), so it needs no external tool and does not load the CUDA library just to list it."""
from __future__ import annotations

import ctypes
import struct
from pathlib import Path

from ..base import DGLError

_PY_FUNCS = {}


class TypeCode:
    """Argument type codes of the reference ABI, kept for code that switches on them."""
    INT = 0
    UINT = 1
    FLOAT = 2
    HANDLE = 3
    NULL = 4
    DGL_TYPE = 5
    DGL_CONTEXT = 6
    ARRAY_HANDLE = 7
    OBJECT_HANDLE = 8
    MODULE_HANDLE = 9
    FUNC_HANDLE = 10
    STR = 11
    BYTES = 12
    NDARRAY_CONTAINER = 13
    EXT_BEGIN = 15


def register_func(func_name, f=None, override=False):
    """Register ``f`` under ``func_name``; usable as ``@register_func("name")`` or ``@register_func`` (uses ``__name__``)."""
    if callable(func_name):
        f, func_name = func_name, func_name.__name__
    if not isinstance(func_name, str):
        raise ValueError("expect string function name")

    def register(fn):
        if not override and func_name in _PY_FUNCS:
            raise DGLError(f"Global function {func_name} is already registered")
        _PY_FUNCS[func_name] = fn
        return fn

    return register(f) if f is not None else register


def _lib_paths():
    from . import loader
    return [loader._LIBDIR / "libdglb200_cpu.so", loader._LIBDIR / "libdglb200_cuda.so"]


def _elf_exports(path: Path, prefix="dglb200_"):
    """Defined global function symbols of an ELF64 shared object whose names start with ``prefix``."""
    try:
        blob = path.read_bytes()
    except OSError:
        return []
    if blob[:4] != b"\x7fELF" or blob[4] != 2 or blob[5] != 1:  # 64-bit little-endian only
        return []
    shoff, = struct.unpack_from("<Q", blob, 0x28)
    shentsize, shnum = struct.unpack_from("<HH", blob, 0x3A)
    sections = [struct.unpack_from("<IIQQQQIIQQ", blob, shoff + i * shentsize) for i in range(shnum)]
    names = []
    for sec in sections:
        if sec[1] != 11:  # SHT_DYNSYM
            continue
        _, _, _, _, off, size, link, _, _, entsize = sec
        stroff = sections[link][4]
        for k in range(size // (entsize or 24)):
            st_name, st_info, _other, st_shndx = struct.unpack_from("<IBBH", blob, off + k * 24)
            if st_shndx == 0 or (st_info & 0xF) != 2 or (st_info >> 4) not in (1, 2):  # defined FUNC, GLOBAL/WEAK
                continue
            end = blob.index(b"\0", stroff + st_name)
            name = blob[stroff + st_name:end].decode("ascii", "replace")
            if name.startswith(prefix):
                names.append(name)
    return names


def list_global_func_names():
    """Registered Python functions plus every exported native entry point (sorted, de-duplicated)."""
    out = set(_PY_FUNCS)
    for p in _lib_paths():
        out.update(_elf_exports(p))
    return sorted(out)


def get_global_func(name, allow_missing=False):
    """Python-registered function, else the native entry point of that symbol name, else ``None`` / DGLError."""
    fn = _PY_FUNCS.get(name)
    if fn is not None:
        return fn
    from . import loader
    for getter in (loader.cpu_lib, loader.cuda_lib):
        try:
            lib = getter()
        except Exception:  # noqa: BLE001 - the CUDA library is optional on CPU-only hosts
            continue
        if lib.has(name):
            raw = getattr(lib.cdll, name)
            raw.restype = ctypes.c_int32
            return raw
    if allow_missing:
        return None
    raise DGLError(f"Cannot find global function {name}")


def extract_ext_funcs(finit):
    """Collect the functions an extension registers through its ``finit(callback)`` entry: the callback receives
    ``(name, function)`` pairs, exactly the protocol of the reference helper, and the dict of them is returned."""
    fdict = {}

    def _list(name, func):
        fdict[name] = func
        return 0

    ret = finit(_list)
    if ret not in (None, 0):
        raise RuntimeError(f"cannot initialize with {finit}")
    return fdict
