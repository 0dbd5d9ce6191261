"""Named POSIX shared memory for tensors / graphs shared between processes of one host
(parity: /root/reference/src/runtime/shared_mem.cc, src/graph/shared_mem_manager.cc,
python/dgl/multiprocessing/pytorch.py:59-104 shared_tensor)."""
from __future__ import annotations

import ctypes
import json

import numpy as np
import torch

from .._native import loader

_DT = {"float32": torch.float32, "float64": torch.float64, "float16": torch.float16, "bfloat16": torch.bfloat16,
       "int32": torch.int32, "int64": torch.int64, "int8": torch.int8, "uint8": torch.uint8, "bool": torch.bool,
       "int16": torch.int16}
_keep = {}


def _wrap(ptr, nbytes, dtype, shape):
    buf = (ctypes.c_char * max(nbytes, 1)).from_address(ptr)
    t = torch.frombuffer(buf, dtype=torch.uint8, count=nbytes) if nbytes else torch.empty(0, dtype=torch.uint8)
    return t.view(dtype).view(*shape) if nbytes else torch.empty(shape, dtype=dtype)


def create_shared_tensor(name, shape, dtype):
    lib = loader.cpu_lib()
    nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty(0, dtype=dtype).element_size()
    p = ctypes.c_void_p()
    lib.fn("dglb200_shm_create", "slp")(name.encode(), nbytes, ctypes.addressof(p))
    t = _wrap(p.value, nbytes, dtype, tuple(shape))
    _keep[name] = (p.value, nbytes)
    return t


def open_shared_tensor(name, shape, dtype):
    lib = loader.cpu_lib()
    nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty(0, dtype=dtype).element_size()
    p = ctypes.c_void_p()
    lib.fn("dglb200_shm_open", "slp")(name.encode(), nbytes, ctypes.addressof(p))
    _keep[name + ":open"] = (p.value, nbytes)
    return _wrap(p.value, nbytes, dtype, tuple(shape))


def shm_exists(name):
    lib = loader.cpu_lib()
    return bool(lib.raw("dglb200_shm_exists", ctypes.c_int, [ctypes.c_char_p])(name.encode()))


def unlink(name):
    loader.cpu_lib().fn("dglb200_shm_unlink", "s")(name.encode())


def tensors_to_shared_memory(name, tensors, meta=None):
    """Copy a dict of tensors into shm segments ``name/<key>`` + a JSON index segment."""
    index = {"meta": meta or {}, "tensors": {}}
    out = {}
    for i, (k, t) in enumerate(tensors.items()):
        t = t.cpu().contiguous()
        seg = f"/{name}_t{i}"
        st = create_shared_tensor(seg, tuple(t.shape), t.dtype)
        st.copy_(t)
        index["tensors"][k] = {"seg": seg, "shape": list(t.shape), "dtype": str(t.dtype).replace("torch.", "")}
        out[k] = st
    blob = json.dumps(index).encode()
    it = create_shared_tensor(f"/{name}_index", (len(blob) + 8,), torch.uint8)
    it[:8] = torch.frombuffer(bytearray(len(blob).to_bytes(8, "little")), dtype=torch.uint8)
    it[8:] = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    return out


def tensors_from_shared_memory(name):
    hdr = open_shared_tensor(f"/{name}_index", (8,), torch.uint8)
    n = int.from_bytes(bytes(hdr.tolist()), "little")
    it = open_shared_tensor(f"/{name}_index", (n + 8,), torch.uint8)
    index = json.loads(bytes(it[8:].tolist()).decode())
    out = {k: open_shared_tensor(v["seg"], tuple(v["shape"]), _DT[v["dtype"]]) for k, v in index["tensors"].items()}
    return out, index["meta"]


def graph_to_shared_memory(g, name, formats=("coo", "csr", "csc")):
    """Copy a graph's structure into shared memory and return a graph backed by it."""
    from ..graph_index import UnitGraph
    tensors = {}
    for i, ug in enumerate(g._graphs):
        row, col = ug.coo()
        tensors[f"row{i}"] = row
        tensors[f"col{i}"] = col
    meta = {"ntypes": g._ntypes, "cet": [list(c) for c in g._canonical_etypes], "num_nodes": g._num_nodes}
    sh = tensors_to_shared_memory(name, tensors, meta)
    graphs = [UnitGraph.from_coo(ug.num_src, ug.num_dst, sh[f"row{i}"], sh[f"col{i}"], formats) for i, ug in enumerate(g._graphs)]
    return g._clone_with(graphs, g._node_frames, g._edge_frames)


def graph_from_shared_memory(name):
    from ..graph_index import UnitGraph
    from ..heterograph import DGLGraph
    t, meta = tensors_from_shared_memory(name)
    cet = [tuple(c) for c in meta["cet"]]
    nn = {nt: n for nt, n in zip(meta["ntypes"], meta["num_nodes"])}
    graphs = [UnitGraph.from_coo(nn[c[0]], nn[c[2]], t[f"row{i}"], t[f"col{i}"]) for i, c in enumerate(cet)]
    return DGLGraph(graphs, meta["ntypes"], cet, meta["num_nodes"])


def create_shared_mem_array(name, shape, dtype):
    """Reference spelling (python/dgl/ndarray.py:create_shared_mem_array) of :func:`create_shared_tensor`."""
    return create_shared_tensor(name, tuple(shape), dtype)


def get_shared_mem_array(name, shape, dtype):
    """Reference spelling (python/dgl/ndarray.py:get_shared_mem_array) of :func:`open_shared_tensor`."""
    return open_shared_tensor(name, tuple(shape), dtype)
