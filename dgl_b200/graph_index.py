"""Per-relation sparse structure with lazily materialised COO / CSR / CSC.

This is the counterpart of the reference's C++ ``UnitGraph``
(/root/reference/src/graph/unit_graph.h:389-401, unit_graph.cc:1418-1519) but
built directly on torch tensors: one relation = (num_src, num_dst) bipartite
structure whose formats are created on demand, can be restricted
(``formats([...])``), pinned in place and moved between devices. g-SpMM
consumes CSC (rows = destinations), its backward consumes CSR; g-SDDMM consumes
COO (src/array/kernel.cc:24,228 in the reference).
"""
from __future__ import annotations

import torch

from .base import DGLError

ALL_FORMATS = ("coo", "csr", "csc")


def _stable_argsort(x):
    return torch.sort(x, stable=True)[1]


def coo_to_csr(row, col, num_rows, idtype):
    """Return (indptr, indices, eids) with rows sorted stably (edge order kept
    inside each row)."""
    E = row.shape[0]
    if E == 0:
        z = torch.zeros(num_rows + 1, dtype=idtype, device=row.device)
        return z, col.clone(), torch.empty(0, dtype=idtype, device=row.device)
    # sorted already? (common for generated/ sampled graphs)
    if E > 1 and bool((row[1:] >= row[:-1]).all()):
        perm = None
        srow, scol = row, col
    else:
        perm = _stable_argsort(row)
        srow, scol = row[perm], col[perm]
    eids = perm.to(idtype) if perm is not None else None
    if row.is_cuda and row.dtype == idtype:
        from ._graph_kernels import row_counts_to_indptr
        return row_counts_to_indptr(srow.contiguous(), num_rows), scol.contiguous(), eids
    counts = torch.bincount(srow.long(), minlength=num_rows)
    indptr = torch.zeros(num_rows + 1, dtype=torch.int64, device=row.device)
    torch.cumsum(counts, 0, out=indptr[1:])
    return indptr.to(idtype), scol.contiguous(), eids


class UnitGraph:
    """One relation's adjacency. Edge ids are positions in the COO arrays."""

    __slots__ = ("num_src", "num_dst", "idtype", "_coo", "_csr", "_csc", "_formats", "_pinned", "_row_sorted",
                 "_degs", "_rev_of", "_dst_view")

    def __init__(self, num_src, num_dst, idtype, coo=None, csr=None, csc=None, formats=ALL_FORMATS):
        self.num_src = int(num_src)
        self.num_dst = int(num_dst)
        self.idtype = idtype
        self._coo = coo  # (row, col)
        self._csr = csr  # (indptr[num_src+1], indices(dst), eids|None)
        self._csc = csc  # (indptr[num_dst+1], indices(src), eids|None)
        self._formats = tuple(formats)
        self._pinned = False
        self._degs = {}
        self._dst_view = None
        self._rev_of = None  # set on reversed views: formats they build are cached on the viewed graph too

    # ---- construction -------------------------------------------------
    @staticmethod
    def from_coo(num_src, num_dst, row, col, formats=ALL_FORMATS):
        if row.dtype != col.dtype:
            raise DGLError("src/dst id dtypes differ")
        return UnitGraph(num_src, num_dst, row.dtype, coo=(row.contiguous(), col.contiguous()), formats=formats)

    @staticmethod
    def from_csr(num_src, num_dst, indptr, indices, eids=None, formats=ALL_FORMATS):
        return UnitGraph(num_src, num_dst, indptr.dtype, csr=(indptr, indices, eids), formats=formats)

    @staticmethod
    def from_csc(num_src, num_dst, indptr, indices, eids=None, formats=ALL_FORMATS):
        return UnitGraph(num_src, num_dst, indptr.dtype, csc=(indptr, indices, eids), formats=formats)

    # ---- properties ---------------------------------------------------
    @property
    def device(self):
        for f in (self._coo, self._csr, self._csc):
            if f is not None:
                return f[0].device
        return torch.device("cpu")

    def num_edges(self):
        if self._coo is not None:
            return int(self._coo[0].shape[0])
        if self._csr is not None:
            return int(self._csr[1].shape[0])
        return int(self._csc[1].shape[0])

    def created_formats(self):
        out = []
        if self._coo is not None:
            out.append("coo")
        if self._csr is not None:
            out.append("csr")
        if self._csc is not None:
            out.append("csc")
        return out

    def allowed_formats(self):
        return list(self._formats)

    def _check_allowed(self, fmt):
        if fmt not in self._formats:
            raise DGLError(f"format '{fmt}' is not allowed on this graph (allowed: {self._formats}); "
                           "use g.formats([...]) to clone with different restrictions")

    def _post_create(self, tensors):
        if self._pinned:
            for t in tensors:
                if t is not None and not t.is_pinned():
                    _pin_inplace(t)

    # ---- format access --------------------------------------------------
    def coo(self):
        if self._coo is None:
            self._check_allowed("coo")
            if self._csr is not None:
                indptr, indices, eids = self._csr
                deg = (indptr[1:] - indptr[:-1]).long()
                row = torch.repeat_interleave(torch.arange(self.num_src, device=indptr.device, dtype=self.idtype), deg)
                col = indices
                if eids is not None:
                    r2 = torch.empty_like(row)
                    c2 = torch.empty_like(col)
                    r2[eids.long()] = row
                    c2[eids.long()] = col
                    row, col = r2, c2
            else:
                indptr, indices, eids = self._csc
                deg = (indptr[1:] - indptr[:-1]).long()
                col = torch.repeat_interleave(torch.arange(self.num_dst, device=indptr.device, dtype=self.idtype), deg)
                row = indices
                if eids is not None:
                    r2 = torch.empty_like(row)
                    c2 = torch.empty_like(col)
                    r2[eids.long()] = row
                    c2[eids.long()] = col
                    row, col = r2, c2
            self._coo = (row.contiguous(), col.contiguous())
            self._post_create(self._coo)
        return self._coo

    def csr(self):
        """(indptr, indices=dst ids, eids|None), rows = source nodes."""
        if self._csr is None:
            self._check_allowed("csr")
            row, col = self.coo() if self._coo is not None or "coo" in self._formats else self._coo_temp()
            self._csr = coo_to_csr(row, col, self.num_src, self.idtype)
            self._post_create(self._csr)
            if self._rev_of is not None and self._rev_of._csc is None and "csc" in self._rev_of._formats:
                self._rev_of._csc = self._csr   # a format built on a reversed VIEW is cached on the graph it views
        return self._csr

    def csc(self):
        """(indptr, indices=src ids, eids|None), rows = destination nodes."""
        if self._csc is None:
            self._check_allowed("csc")
            row, col = self.coo() if self._coo is not None or "coo" in self._formats else self._coo_temp()
            self._csc = coo_to_csr(col, row, self.num_dst, self.idtype)
            self._post_create(self._csc)
            if self._rev_of is not None and self._rev_of._csr is None and "csr" in self._rev_of._formats:
                self._rev_of._csr = self._csc
        return self._csc

    def _coo_temp(self):
        # build COO without caching when 'coo' is a disallowed format
        saved = self._formats
        self._formats = ALL_FORMATS
        try:
            coo = self.coo()
        finally:
            self._formats = saved
            self._coo = None
        return coo

    def edges(self, order="eid"):
        """(src, dst, eid) with order in {'eid', 'srcdst', 'dst'... }"""
        if order in (None, "eid"):
            row, col = self.coo() if "coo" in self._formats or self._coo is not None else self._coo_temp()
            return row, col, None
        if order == "srcdst":
            indptr, indices, eids = self.csr() if ("csr" in self._formats or self._csr is not None) else \
                coo_to_csr(*self.coo(), self.num_src, self.idtype)
            deg = (indptr[1:] - indptr[:-1]).long()
            row = torch.repeat_interleave(torch.arange(self.num_src, device=indptr.device, dtype=self.idtype), deg)
            # sort dst within row for canonical srcdst order
            key = row.long() * max(self.num_dst, 1) + indices.long()
            perm = _stable_argsort(key)
            e = eids if eids is not None else torch.arange(indices.shape[0], device=indptr.device, dtype=self.idtype)
            return row[perm], indices[perm], e[perm]
        raise DGLError(f"unknown edge order {order}")

    def in_degrees(self):
        d = self._degs.get("in")
        if d is None:
            if self._csc is not None:
                ip = self._csc[0]
                d = ip[1:] - ip[:-1]
            else:
                _, col = self.coo() if self._coo is not None else self._coo_temp()
                d = torch.bincount(col.long(), minlength=self.num_dst).to(self.idtype)
            self._degs["in"] = d
        return d

    def out_degrees(self):
        d = self._degs.get("out")
        if d is None:
            if self._csr is not None:
                ip = self._csr[0]
                d = ip[1:] - ip[:-1]
            else:
                row, _ = self.coo() if self._coo is not None else self._coo_temp()
                d = torch.bincount(row.long(), minlength=self.num_src).to(self.idtype)
            self._degs["out"] = d
        return d

    # ---- transforms -----------------------------------------------------
    def dst_sorted_view(self):
        """(view, eid_map): a relation with the same edges whose edge ids are the positions of this graph's CSC, i.e.
        edges grouped by destination. g-SpMM / edge-softmax on the view need no edge-id indirection and g-SDDMM writes
        its result already in that order, so a layer that creates AND consumes its edge data internally (attention)
        never permutes an [E, ...] tensor. ``eid_map[i]`` = original id of view edge i (None if already ordered).
        The view is cached; the formats it builds (its CSR for the backward pass) live as long as the graph."""
        indptr, indices, eids = self.csc()
        if eids is None:
            return self, None
        v = getattr(self, "_dst_view", None)
        if v is None:
            v = UnitGraph.from_csc(self.num_src, self.num_dst, indptr, indices, None, ALL_FORMATS)
            self._dst_view = v
        return v, eids

    def reverse(self):
        """Transposed relation sharing storage (CSR <-> CSC swap)."""
        g = UnitGraph(self.num_dst, self.num_src, self.idtype,
                      coo=None if self._coo is None else (self._coo[1], self._coo[0]),
                      csr=self._csc, csc=self._csr, formats=_swap_formats(self._formats))
        g._pinned = self._pinned
        g._rev_of = self
        return g

    def _map(self, fn):
        def m(f):
            return None if f is None else tuple(None if t is None else fn(t) for t in f)

        g = UnitGraph(self.num_src, self.num_dst, self.idtype, m(self._coo), m(self._csr), m(self._csc), self._formats)
        return g

    def to(self, device, non_blocking=False):
        device = torch.device(device)
        if device == self.device:
            return self
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def astype(self, idtype):
        if idtype == self.idtype:
            return self
        g = self._map(lambda t: t.to(idtype))
        g.idtype = idtype
        return g

    def with_formats(self, formats):
        formats = tuple(formats)
        for f in formats:
            if f not in ALL_FORMATS:
                raise DGLError(f"unknown sparse format {f}")
        keep = {}
        for name in ALL_FORMATS:
            cur = getattr(self, "_" + name)
            keep[name] = cur if name in formats else None
        if all(v is None for v in keep.values()):
            # need at least one format: derive the first allowed one
            first = formats[0]
            keep[first] = getattr(self, first)() if first in self._formats else _derive(self, first)
        return UnitGraph(self.num_src, self.num_dst, self.idtype, keep["coo"], keep["csr"], keep["csc"], formats)

    def create_formats_(self):
        for f in self._formats:
            getattr(self, f)()

    # ---- pinning / streams ---------------------------------------------------
    def pin_memory_(self):
        if self.device.type != "cpu":
            raise DGLError("only CPU graphs can be pinned")
        for f in (self._coo, self._csr, self._csc):
            if f is not None:
                for t in f:
                    if t is not None and not t.is_pinned():
                        _pin_inplace(t)
        self._pinned = True

    def unpin_memory_(self):
        for f in (self._coo, self._csr, self._csc):
            if f is not None:
                for t in f:
                    if t is not None and t.is_pinned():
                        _unpin_inplace(t)
        self._pinned = False

    def is_pinned(self):
        return self._pinned

    def record_stream(self, stream):
        for f in (self._coo, self._csr, self._csc):
            if f is not None:
                for t in f:
                    if t is not None and t.is_cuda:
                        t.record_stream(stream)


def _derive(g, fmt):
    saved = g._formats
    g._formats = ALL_FORMATS
    try:
        return getattr(g, fmt)()
    finally:
        g._formats = saved


def _swap_formats(formats):
    m = {"coo": "coo", "csr": "csc", "csc": "csr"}
    return tuple(m[f] for f in formats)


def _pin_inplace(t):
    """cudaHostRegister the tensor's storage (in-place pin, reference:
    src/runtime/cuda/cuda_device_api.cc:238-247)."""
    if t.numel() == 0:
        return
    if torch.cuda.is_available():
        rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
        if int(rc) != 0:
            raise DGLError(f"cudaHostRegister failed with code {int(rc)}")


def _unpin_inplace(t):
    if t.numel() == 0:
        return
    if torch.cuda.is_available():
        torch.cuda.cudart().cudaHostUnregister(t.data_ptr())
