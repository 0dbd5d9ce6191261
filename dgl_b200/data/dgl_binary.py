"""Reader / writer for the reference's on-disk binary containers, so files produced by either framework load in the other:

* graph files of ``dgl.save_graphs`` (version 2, heterograph records; /root/reference/src/graph/serialize/
  heterograph_serialize.cc:47-200, heterograph_data.h:75-90, src/graph/heterograph.cc:564-584,
  src/graph/unit_graph.cc:433-445,820-832,1654-1745, src/graph/immutable_graph.cc:305-310,627-644,
  include/dgl/aten/coo.h:89-114, include/dgl/aten/csr.h:84-107),
* tensor dictionaries of ``dgl.data.save_tensors`` (src/graph/serialize/tensor_serialize.cc:25-66),
* the NDArray record both are made of (include/dgl/runtime/ndarray.h:619-665, src/runtime/ndarray.cc:366-407).

Design: the file is memory-mapped and walked once with ``struct``; every array body becomes a tensor over the mapping
(copy-on-write pages, so nothing is read until touched and nothing is copied), which makes partition files of tens of GB
open in milliseconds. Writing streams array bodies straight from tensor memory (``memoryview``), patching the three
forward offsets of the layout at the end. Everything is little-endian, as the reference writes on the platforms it runs on."""
from __future__ import annotations

import mmap
import os
import struct

import numpy as np
import torch

from ..base import DGLError

GRAPHS_MAGIC = 0xDD2E4FF046B4A13F
HETERO_MAGIC = 0xDD589FBE35224ABF
UNIT_MAGIC = 0xDD2E60F0F6B4A127
IMGRAPH_MAGIC = 0xDD3C5FFE20046ABF
COO_MAGIC = 0xDD61FFD305DFF127
CSR_MAGIC = 0xDD6CD31205DFF127
NDARRAY_MAGIC = 0xDD5E40F096B4A13F
TENSORS_MAGIC = 0xDD5A9FBE3FA2443F
_COO, _CSR, _CSC = 1, 2, 4
_HETERO_GRAPH_TYPE = 2  # GraphType::kHeteroGraph
_META_BYTES = 4096

# (type code, bits) <-> numpy dtype carrying the bytes; bfloat16 travels as int16 and is re-viewed in torch
_NP_OF = {(0, 8): np.int8, (0, 16): np.int16, (0, 32): np.int32, (0, 64): np.int64,
          (1, 8): np.uint8, (1, 16): np.uint16, (1, 32): np.uint32, (1, 64): np.uint64,
          (2, 16): np.float16, (2, 32): np.float32, (2, 64): np.float64, (4, 16): np.int16}
_CODE_OF = {torch.int8: (0, 8), torch.int16: (0, 16), torch.int32: (0, 32), torch.int64: (0, 64),
            torch.uint8: (1, 8), torch.bool: (1, 8), torch.float16: (2, 16), torch.float32: (2, 32),
            torch.float64: (2, 64), torch.bfloat16: (4, 16)}


def is_dgl_binary(path) -> bool:
    """True when the file starts with the reference's graph-container or tensor-container magic."""
    try:
        with open(path, "rb") as fh:
            head = fh.read(8)
    except OSError:
        return False
    return len(head) == 8 and struct.unpack("<Q", head)[0] in (GRAPHS_MAGIC, TENSORS_MAGIC)


# ---------------------------------------------------------------------------------------------------------- reading


class _Cursor:
    """Position in a read-only mapping; scalar reads via struct, array bodies as zero-copy views."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise DGLError(f"File {path} does not exist.")
        self.path = path
        self._fh = open(path, "rb")
        size = os.fstat(self._fh.fileno()).st_size
        if size < 16:
            raise DGLError(f"{path}: too short to be a DGL binary file")
        # ACCESS_COPY: private copy-on-write pages, so the tensors handed out are writable without touching the file
        self.buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_COPY)
        self.size = size
        self.pos = 0

    def take(self, fmt):
        n = struct.calcsize(fmt)
        if self.pos + n > self.size:
            raise DGLError(f"{self.path}: truncated at byte {self.pos}")
        out = struct.unpack_from(fmt, self.buf, self.pos)
        self.pos += n
        return out if len(out) > 1 else out[0]

    def u64(self):
        return self.take("<Q")

    def expect(self, magic, what):
        got = self.u64()
        if got != magic:
            raise DGLError(f"{self.path}: bad {what} marker 0x{got:016X} at byte {self.pos - 8}")

    def string(self):
        n = self.u64()
        if self.pos + n > self.size:
            raise DGLError(f"{self.path}: truncated string at byte {self.pos}")
        s = bytes(self.buf[self.pos:self.pos + n]).decode("utf-8")
        self.pos += n
        return s

    def ndarray(self):
        self.expect(NDARRAY_MAGIC, "NDArray")
        self.u64()  # reserved
        _dev_type, _dev_id, ndim = self.take("<iii")
        code, bits, lanes = self.take("<BBH")
        if lanes != 1 or (code, bits) not in _NP_OF:
            raise DGLError(f"{self.path}: unsupported array dtype (code={code}, bits={bits}, lanes={lanes})")
        shape = [self.take("<q") for _ in range(ndim)]
        nbytes = self.take("<q")
        count = int(np.prod(shape, dtype=np.int64)) if ndim else 1
        if nbytes != count * (bits // 8) or self.pos + nbytes > self.size:
            raise DGLError(f"{self.path}: inconsistent array record at byte {self.pos}")
        npdt = np.dtype(_NP_OF[(code, bits)])
        if nbytes:
            arr = np.frombuffer(self.buf, dtype=npdt, count=count, offset=self.pos)
        else:
            arr = np.empty(0, dtype=npdt)
        self.pos += nbytes
        if npdt in (np.uint16, np.uint32, np.uint64):
            t = torch.from_numpy(arr.astype({2: np.int16, 4: np.int32, 8: np.int64}[npdt.itemsize], copy=False))
        else:
            t = torch.from_numpy(arr)
        if code == 4:
            t = t.view(torch.bfloat16)
        return t.reshape(shape)

    def named_tensors(self):
        return [(self.string(), self.ndarray()) for _ in range(self.u64())]

    def csr_matrix(self):
        self.expect(CSR_MAGIC, "CSR matrix")
        num_cols, num_rows = self.take("<qq")
        indptr, indices, data = self.ndarray(), self.ndarray(), self.ndarray()
        self.take("<?")  # sorted flag
        return num_rows, num_cols, indptr, indices, data

    def coo_matrix(self):
        self.expect(COO_MAGIC, "COO matrix")
        num_cols, num_rows = self.take("<qq")
        row, col, data = self.ndarray(), self.ndarray(), self.ndarray()
        self.take("<??")  # row_sorted, col_sorted
        return num_rows, num_cols, row, col, data

    def meta_graph(self):
        """ImmutableGraph record: out-CSR over node types; returns [(src type id, dst type id)] by edge-type id."""
        self.expect(IMGRAPH_MAGIC, "metagraph")
        _, _, indptr, indices, data = self.csr_matrix()
        src = torch.repeat_interleave(torch.arange(indptr.numel() - 1), indptr[1:].long() - indptr[:-1].long())
        order = data.long() if data.numel() else torch.arange(indices.numel())
        pairs = [None] * indices.numel()
        for s, d, e in zip(src.tolist(), indices.tolist(), order.tolist()):
            pairs[e] = (s, d)
        return pairs

    def unit_graph(self):
        """One relation -> (num_src, num_dst, src ids, dst ids) in edge-id order."""
        self.expect(UNIT_MAGIC, "unit graph")
        saved, _allowed = self.take("<qq")
        saved = (saved & 0xFFFFFFFF) if (saved >> 32) else (1 << {0: 0, 1: 0, 2: 1, 3: 2}.get(saved & 0xFFFFFFFF, 0))
        out = None
        if saved & _COO:
            self.meta_graph()
            nr, nc, row, col, data = self.coo_matrix()
            if data.numel():
                inv = torch.empty_like(data, dtype=torch.int64)
                inv[data.long()] = torch.arange(data.numel())
                row, col = row[inv], col[inv]
            out = (nr, nc, row, col)
        for code in (_CSR, _CSC):
            if not saved & code:
                continue
            self.meta_graph()
            nr, nc, indptr, indices, data = self.csr_matrix()
            if out is not None:
                continue
            major = torch.repeat_interleave(torch.arange(nr, dtype=indices.dtype), indptr[1:].long() - indptr[:-1].long())
            minor = indices
            if data.numel():
                inv = torch.empty(data.numel(), dtype=torch.int64)
                inv[data.long()] = torch.arange(data.numel())
                major, minor = major[inv], minor[inv]
            out = (nr, nc, major, minor) if code == _CSR else (nc, nr, minor, major)
        if out is None:
            raise DGLError(f"{self.path}: relation stored without any sparse format")
        return out

    def hetero_graph_data(self):
        from ..frame import Frame
        from ..graph_index import UnitGraph
        from ..heterograph import DGLGraph
        self.expect(HETERO_MAGIC, "heterograph")
        meta = self.meta_graph()
        rels = [self.unit_graph() for _ in range(self.u64())]
        num_nodes = [self.take("<q") for _ in range(self.u64())]
        ndata = [dict(self.named_tensors()) for _ in range(self.u64())]
        edata = [dict(self.named_tensors()) for _ in range(self.u64())]
        ntypes = [self.string() for _ in range(self.u64())]
        etypes = [self.string() for _ in range(self.u64())]
        if len(meta) != len(rels) or len(etypes) != len(rels) or len(ntypes) != len(num_nodes):
            raise DGLError(f"{self.path}: heterograph record is inconsistent")
        cet = [(ntypes[s], etypes[i], ntypes[d]) for i, (s, d) in enumerate(meta)]
        graphs = [UnitGraph.from_coo(num_nodes[s], num_nodes[d], r[2], r[3]) for (s, d), r in zip(meta, rels)]
        g = DGLGraph(graphs, ntypes, cet, num_nodes)
        g._node_frames = [Frame(f, num_rows=n) for f, n in zip(ndata, num_nodes)]
        g._edge_frames = [Frame(f, num_rows=ug.num_edges()) for f, ug in zip(edata, graphs)]
        return g

    def header(self):
        self.expect(GRAPHS_MAGIC, "graph file")
        version, gtype, num_graph = self.take("<QQQ")
        if version != 2 or gtype != _HETERO_GRAPH_TYPE:
            raise DGLError(f"{self.path}: graph file version {version} / graph type {gtype} is not supported "
                           "(only the version-2 heterograph layout; re-save old files with the reference)")
        self.pos = _META_BYTES
        return num_graph, self.u64()


def read_labels(path):
    c = _Cursor(path)
    c.header()
    return dict(c.named_tensors())


def read_graphs(path, idx_list=None):
    """(list of DGLGraph, label dict) of a reference ``save_graphs`` file; ``idx_list`` seeks through the trailing
    offset table, so only the selected records are parsed."""
    c = _Cursor(path)
    num_graph, gdata_start = c.header()
    labels = dict(c.named_tensors())
    if idx_list is None:
        c.pos = gdata_start
        return [c.hetero_graph_data() for _ in range(num_graph)], labels
    table_bytes = struct.unpack_from("<Q", c.buf, c.size - 8)[0]
    c.pos = c.size - 8 - table_bytes
    offsets = [c.u64() for _ in range(c.u64())]
    graphs = []
    for i in idx_list:
        if not 0 <= int(i) < len(offsets):
            raise DGLError(f"ID {i} in idx_list is out of bound: the file holds {len(offsets)} graphs")
        c.pos = offsets[int(i)]
        graphs.append(c.hetero_graph_data())
    return graphs, labels


def read_tensors(path):
    c = _Cursor(path)
    c.expect(TENSORS_MAGIC, "tensor file")
    c.u64()  # element count, repeated by the vector header that follows
    return dict(c.named_tensors())


# ---------------------------------------------------------------------------------------------------------- writing


class _Sink:
    def __init__(self, fh):
        self.fh = fh
        self.count = 0

    def raw(self, b):
        self.fh.write(b)
        self.count += len(b)

    def put(self, fmt, *vals):
        self.raw(struct.pack(fmt, *vals))

    def string(self, s):
        b = s.encode("utf-8")
        self.put("<Q", len(b))
        self.raw(b)

    def ndarray(self, t):
        t = torch.as_tensor(t).detach()
        if t.device.type != "cpu":
            t = t.cpu()
        if t.dtype not in _CODE_OF:
            raise DGLError(f"dtype {t.dtype} cannot be stored in the DGL binary format")
        code, bits = _CODE_OF[t.dtype]
        t = t.contiguous()
        self.put("<QQiii", NDARRAY_MAGIC, 0, 1, 0, t.dim())  # context is always CPU (device type 1, id 0)
        self.put("<BBH", code, bits, 1)
        for s in t.shape:
            self.put("<q", s)
        nbytes = t.numel() * (bits // 8)
        self.put("<q", nbytes)
        if nbytes:
            body = t.view(torch.int16) if t.dtype == torch.bfloat16 else (t.view(torch.uint8) if t.dtype == torch.bool else t)
            mv = memoryview(body.numpy()).cast("B")
            self.fh.write(mv)
            self.count += nbytes

    def named_tensors(self, items):
        items = list(items)
        self.put("<Q", len(items))
        for k, v in items:
            self.string(k)
            self.ndarray(v)

    def csr_matrix(self, num_rows, num_cols, indptr, indices, data, is_sorted=False):
        self.put("<Qqq", CSR_MAGIC, num_cols, num_rows)
        self.ndarray(indptr)
        self.ndarray(indices)
        self.ndarray(data)
        self.put("<?", is_sorted)

    def meta_graph(self, num_types, pairs):
        """out-CSR of the type graph with explicit edge-type ids as data."""
        src = torch.tensor([p[0] for p in pairs], dtype=torch.int64)
        dst = torch.tensor([p[1] for p in pairs], dtype=torch.int64)
        order = torch.argsort(src, stable=True)
        indptr = torch.zeros(num_types + 1, dtype=torch.int64)
        indptr[1:] = torch.cumsum(torch.bincount(src, minlength=num_types), 0)
        self.put("<Q", IMGRAPH_MAGIC)
        self.csr_matrix(num_types, num_types, indptr, dst[order], order.clone(), False)

    def unit_graph(self, ug, same_type, formats, idtype):
        self.put("<Q", UNIT_MAGIC)
        code = 0
        for f in formats:
            code |= {"coo": _COO, "csr": _CSR, "csc": _CSC}[f]
        self.put("<qq", code | 0x100000000, 7 | 0x100000000)
        nsrc, ndst = ug.num_src, ug.num_dst
        meta = (1, [(0, 0)]) if same_type else (2, [(0, 1)])
        null = torch.empty(0, dtype=idtype)
        if code & _COO:
            row, col = ug.coo()
            self.meta_graph(*meta)
            self.put("<Qqq", COO_MAGIC, ndst, nsrc)
            self.ndarray(row.to(idtype))
            self.ndarray(col.to(idtype))
            self.ndarray(null)
            self.put("<??", False, False)
        if code & _CSR:
            indptr, indices, eids = ug.csr()
            self.meta_graph(*meta)
            self.csr_matrix(nsrc, ndst, indptr.to(idtype), indices.to(idtype), null if eids is None else eids.to(idtype))
        if code & _CSC:
            indptr, indices, eids = ug.csc()
            self.meta_graph(*meta)
            self.csr_matrix(ndst, nsrc, indptr.to(idtype), indices.to(idtype), null if eids is None else eids.to(idtype))

    def hetero_graph_data(self, g, formats):
        ntypes = list(g.ntypes)
        nid = {t: i for i, t in enumerate(ntypes)}
        pairs = [(nid[c[0]], nid[c[2]]) for c in g.canonical_etypes]
        self.put("<Q", HETERO_MAGIC)
        self.meta_graph(len(ntypes), pairs)
        self.put("<Q", len(g._graphs))
        for ug, (s, d) in zip(g._graphs, pairs):
            self.unit_graph(ug, s == d, formats, g.idtype)
        self.put("<Q", len(ntypes))
        for t in ntypes:
            self.put("<q", g.num_nodes(t))
        self.put("<Q", len(ntypes))
        for f in g._node_frames:
            self.named_tensors((k, f[k]) for k in f.keys())
        self.put("<Q", len(g._edge_frames))
        for f in g._edge_frames:
            self.named_tensors((k, f[k]) for k in f.keys())
        self.put("<Q", len(ntypes))
        for t in ntypes:
            self.string(t)
        self.put("<Q", len(g.canonical_etypes))
        for c in g.canonical_etypes:
            self.string(c[1])


def write_graphs(path, g_list, labels=None, formats=None):
    """Write a version-2 graph file the reference's ``dgl.load_graphs`` reads back (and :func:`read_graphs`)."""
    from ..heterograph import DGLGraph
    if isinstance(g_list, DGLGraph):
        g_list = [g_list]
    for g in g_list:
        if type(g) is not DGLGraph:
            raise DGLError("Invalid argument g_list. Must be a DGLGraph or a list of DGLGraphs.")
    if formats is None or formats == []:
        formats = ["coo"]
    elif isinstance(formats, str):
        formats = [formats]
    for f in formats:
        if f not in ("coo", "csr", "csc"):
            raise DGLError(f"unknown sparse format {f!r}")
    labels = labels or {}
    tmp = f"{path}.tmp{os.getpid()}"
    with open(tmp, "wb") as fh:
        sink = _Sink(fh)
        head = struct.pack("<QQQQ", GRAPHS_MAGIC, 2, _HETERO_GRAPH_TYPE, len(g_list))
        sink.raw(head + b"\0" * (_META_BYTES - len(head)))
        at = fh.tell()
        sink.put("<Q", 0)  # start of the graph records, patched below (lets readers skip the label block)
        sink.named_tensors(labels.items())
        gdata_start = sink.count
        offsets = []
        for g in g_list:
            offsets.append(sink.count)
            sink.hetero_graph_data(g, formats)
        table = struct.pack(f"<Q{len(offsets)}Q", len(offsets), *offsets)
        sink.raw(table)
        sink.put("<Q", len(table))
        fh.seek(at)
        fh.write(struct.pack("<Q", gdata_start))
    os.replace(tmp, path)


def write_tensors(path, tensor_dict):
    tmp = f"{path}.tmp{os.getpid()}"
    with open(tmp, "wb") as fh:
        sink = _Sink(fh)
        sink.put("<QQ", TENSORS_MAGIC, len(tensor_dict))
        sink.named_tensors(tensor_dict.items())
    os.replace(tmp, path)
