"""Symmetric heap over CUDA VMM + NVLS multicast (csrc/cuda/symm.cu).

``SymmetricHeap.alloc(nbytes)`` is a collective: every rank creates a physical allocation, the file descriptors that
name them are exchanged over unix sockets (SCM_RIGHTS), every rank maps all allocations back to back (peer r at
``base + r * stride``) and - when the switch supports it - binds them to one multicast object whose mapping gives
``multimem.ld_reduce`` (in-switch sum over the ranks) and ``multimem.st`` (store to every rank). torch.distributed is
only the bootstrap (rendezvous directory, barriers between the driver-level phases).

The reference has no counterpart: its collectives are NCCL calls (python/dgl/cuda/nccl.py, DDP).
"""
from __future__ import annotations

import ctypes
import os
import socket
import tempfile
import threading

import torch
import torch.distributed as dist

from .._native import loader
from ..base import DGLError


class _RawCuda:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class SymmBuffer:
    """One symmetric allocation: ``local`` (uint8 torch view of this rank's copy), ``peer_ptr(r)``, ``mc_ptr`` (0 without
    multicast) and the padded per-rank size ``stride``."""

    def __init__(self, heap, base, stride, nbytes, mc_ptr, handles, mc_handle):
        self.heap, self.base, self.stride, self.nbytes, self.mc_ptr = heap, base, stride, nbytes, mc_ptr
        self._handles, self._mc_handle = handles, mc_handle
        self.local_ptr = base + heap.rank * stride
        self._raw = _RawCuda(self.local_ptr, nbytes)
        self.local = torch.as_tensor(self._raw, device=heap.device)

    def peer_ptr(self, r):
        return self.base + r * self.stride

    def view(self, dtype, offset_bytes=0, numel=None):
        t = self.local[offset_bytes:].view(dtype)
        return t if numel is None else t[:numel]


class SymmetricHeap:
    def __init__(self, group=None, device=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise DGLError("SymmetricHeap needs an initialised process group (bootstrap only)")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lib = loader.cuda_lib()
        vmm, mc, gran = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int64(0)
        try:
            self.lib.fn("dglb200_symm_query", "ippp")(self.device.index, ctypes.addressof(vmm), ctypes.addressof(mc),
                                                      ctypes.addressof(gran))
        except Exception as e:  # noqa: BLE001 - an old driver: no VMM entry points
            self._why = str(e)
            vmm.value = mc.value = 0
        flags = torch.tensor([vmm.value, mc.value], device=self.device, dtype=torch.int32)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN, group=group)          # every rank must agree on the path
        self.vmm_ok, self.mc_ok = bool(flags[0].item()), bool(flags[1].item()) and os.environ.get("DGLB200_NVLS", "1") != "0"
        self.gran = max(int(gran.value), 1 << 21)
        box = [tempfile.mkdtemp(prefix="dglb200_symm_") if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        self._dir, self._seq, self._buffers = box[0], 0, []

    # ------------------------------------------------------------------ fd exchange (SCM_RIGHTS)
    def _sock_path(self, tag, r):
        return os.path.join(self._dir, f"{tag}_{r}.sock")

    def _all_gather_fd(self, fd, tag):
        """fds[r] = a local descriptor for rank r's object (own entry = ``fd``)."""
        path = self._sock_path(tag, self.rank)
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(path)
        srv.listen(self.world)
        dist.barrier(group=self.group)                      # every listener exists before anyone connects

        def serve():
            for _ in range(self.world - 1):
                conn, _ = srv.accept()
                socket.send_fds(conn, [b"f"], [fd])
                conn.close()

        th = threading.Thread(target=serve, daemon=True)
        th.start()
        fds = [None] * self.world
        fds[self.rank] = fd
        for r in range(self.world):
            if r == self.rank:
                continue
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(self._sock_path(tag, r))
            _, got, _, _ = socket.recv_fds(c, 16, 1)
            c.close()
            fds[r] = got[0]
        th.join()
        srv.close()
        dist.barrier(group=self.group)
        try:
            os.unlink(path)
        except OSError:
            pass
        return fds

    def _bcast_fd(self, fd, tag, src=0):
        """Every rank obtains a local descriptor for the object rank ``src`` exported."""
        if self.rank == src:
            path = self._sock_path(tag, src)
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(path)
            srv.listen(self.world)
        dist.barrier(group=self.group)
        out = fd
        if self.rank == src:
            for _ in range(self.world - 1):
                conn, _ = srv.accept()
                socket.send_fds(conn, [b"f"], [fd])
                conn.close()
            srv.close()
        else:
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(self._sock_path(tag, src))
            _, got, _, _ = socket.recv_fds(c, 16, 1)
            c.close()
            out = got[0]
        dist.barrier(group=self.group)
        return out

    def _all_ok(self, ok):
        """Collective AND over the ranks: a driver-level failure on one rank switches the feature off everywhere instead
        of leaving the others blocked in the next phase."""
        t = torch.tensor([1 if ok else 0], device=self.device, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(t.item())

    def _bind_multicast(self, local_handle, stride, tag):
        """(multicast pointer, multicast handle) of the team object over this allocation, or (0, 0) - and multicast is
        disabled for the heap - if any rank fails in any phase (no fabric manager, no NVSwitch, permissions ...)."""
        mh, mfd = ctypes.c_uint64(0), ctypes.c_int(-1)
        ok = True
        if self.rank == 0:
            try:
                self.lib.fn("dglb200_symm_mc_create", "ilpp")(self.world, stride, ctypes.addressof(mh), ctypes.addressof(mfd))
            except Exception as e:  # noqa: BLE001
                self._why, ok = f"cuMulticastCreate: {e}", False
        if not self._all_ok(ok):
            self.mc_ok = False
            return 0, 0
        got = self._bcast_fd(mfd.value, tag + "m", 0)
        p = ctypes.c_void_p()
        try:
            if self.rank != 0:
                self.lib.fn("dglb200_symm_import", "ip")(got, ctypes.addressof(mh))
            os.close(got)
            self.lib.fn("dglb200_symm_mc_add_device", "li")(mh.value, self.device.index)
        except Exception as e:  # noqa: BLE001
            self._why, ok = f"multicast team: {e}", False
        if not self._all_ok(ok):                            # also the barrier: all devices joined before anyone binds
            self.mc_ok = False
            return 0, 0
        try:
            self.lib.fn("dglb200_symm_mc_bind_map", "illllp")(self.device.index, mh.value, local_handle, stride, self.gran,
                                                              ctypes.addressof(p))
        except Exception as e:  # noqa: BLE001
            self._why, ok = f"cuMulticastBindMem: {e}", False
        if not self._all_ok(ok):
            self.mc_ok = False
            return 0, 0
        return p.value, mh.value

    # ------------------------------------------------------------------ allocation
    def alloc(self, nbytes, multicast=True):
        """Collective. Returns a :class:`SymmBuffer`; its memory is zero-initialised."""
        if not self.vmm_ok:
            raise DGLError("VMM symmetric memory is not available on this system")
        stride = (int(nbytes) + self.gran - 1) // self.gran * self.gran
        tag = f"b{self._seq}"
        self._seq += 1
        h, fd = ctypes.c_uint64(0), ctypes.c_int(-1)
        self.lib.fn("dglb200_symm_create", "ilpp")(self.device.index, stride, ctypes.addressof(h), ctypes.addressof(fd))
        fds = self._all_gather_fd(fd.value, tag)
        handles = (ctypes.c_uint64 * self.world)()
        for r, f in enumerate(fds):
            if r == self.rank:
                handles[r] = h.value
            else:
                hh = ctypes.c_uint64(0)
                self.lib.fn("dglb200_symm_import", "ip")(f, ctypes.addressof(hh))
                handles[r] = hh.value
                os.close(f)
        os.close(fd.value)
        base = ctypes.c_void_p()
        self.lib.fn("dglb200_symm_map", "ipillp")(self.device.index, ctypes.cast(handles, ctypes.c_void_p), self.world, stride,
                                                  self.gran, ctypes.addressof(base))
        mc_ptr, mc_handle = 0, 0
        if multicast and self.mc_ok and self.world > 1:
            mc_ptr, mc_handle = self._bind_multicast(h.value, stride, tag)
        buf = SymmBuffer(self, base.value, stride, int(nbytes), mc_ptr, list(handles), mc_handle)
        buf.local.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._buffers.append(buf)
        return buf
