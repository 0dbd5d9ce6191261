"""Symmetric peer-mapped memory + device-side collectives over NVLink 5 / NVSwitch.

``P2PComm`` allocates buffers with cudaMalloc, exchanges CUDA-IPC handles through the (NCCL/gloo) process
group once at creation time, and afterwards every rank holds a table of peer pointers it can pass to
kernels: barrier, one-shot all-reduce, halo-fused g-SpMM, sharded-table gather / scatter-add.
Replaces the reference's python/dgl/cuda/nccl.py + DDP all-reduce + cooperative all-to-all
 for the single-node 8xB200 case; NCCL remains the bootstrap and the baseline.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from .._native import loader
from ..base import DGLError, dtype_code, id_code


class _RawCuda:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class SymmetricBuffer:
    """One allocation per rank of identical size; ``local`` is a torch view of this rank's copy,
    ``ptrs`` the table of all ranks' device pointers (index = rank)."""

    def __init__(self, comm, nbytes):
        self.comm = comm
        self.nbytes = int(nbytes)
        lib = comm.lib
        p = ctypes.c_void_p()
        lib.fn("dglb200_comm_malloc", "lp")(self.nbytes, ctypes.addressof(p))
        self.local_ptr = p.value
        ptrs = [None] * comm.world_size
        ptrs[comm.rank] = self.local_ptr
        if comm.world_size > 1:
            handle = (ctypes.c_char * 64)()
            lib.fn("dglb200_comm_ipc_get", "pp")(self.local_ptr, ctypes.addressof(handle))
            gathered = [None] * comm.world_size
            dist.all_gather_object(gathered, bytes(handle.raw), group=comm.group)
            for r, h in enumerate(gathered):
                if r == comm.rank:
                    continue
                buf = ctypes.create_string_buffer(h, 64)
                q = ctypes.c_void_p()
                lib.fn("dglb200_comm_ipc_open", "pp")(ctypes.addressof(buf), ctypes.addressof(q))
                ptrs[r] = q.value
        self.ptrs = ptrs
        self.ptr_array = (ctypes.c_void_p * comm.world_size)(*ptrs)
        self._raw = _RawCuda(self.local_ptr, self.nbytes)
        self.local_bytes = torch.as_tensor(self._raw, device=comm.device)

    def _close_peer_mappings(self):
        for r, p in enumerate(self.ptrs):
            if r != self.comm.rank and p:
                try:
                    self.comm.lib.fn("dglb200_comm_ipc_close", "p")(p)
                except Exception:  # noqa: BLE001 - best effort during teardown
                    pass
        self.ptrs = [p if r == self.comm.rank else None for r, p in enumerate(self.ptrs)]

    def _free_local(self):
        if self.local_ptr:
            self.local_bytes = None
            self._raw = None
            try:
                self.comm.lib.fn("dglb200_comm_free", "p")(self.local_ptr)
            except Exception:  # noqa: BLE001
                pass
            self.local_ptr = 0

    def view(self, dtype, shape=None):
        t = self.local_bytes.view(dtype)
        return t if shape is None else t[: int(torch.tensor(shape).prod())].view(*shape)

    def ptr_table(self):
        return ctypes.cast(self.ptr_array, ctypes.c_void_p)


class P2PComm:
    """Per-process communicator over peer-mapped memory. Requires all ranks on one NVLink domain."""

    def __init__(self, group=None, device=None, timeout_s=20.0):
        if not torch.cuda.is_available():
            raise DGLError("P2PComm needs CUDA devices")
        self.group = group
        if dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world_size = dist.get_world_size(group)
        else:
            self.rank, self.world_size = 0, 1
        if self.world_size > 8:
            raise DGLError("P2PComm supports up to 8 ranks (one NVSwitch domain)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lib = loader.cuda_lib()
        self.timeout_s = float(timeout_s)
        self._flags = SymmetricBuffer(self, 8 * 8 * 4)  # 4 barrier channels x 8 ranks x u64
        self._epoch = torch.zeros(4, dtype=torch.int64, device=self.device)
        self._error = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._grad = None
        self._bufs = []
        # NVLS path (VMM symmetric heap + multicast, parallel/symm.py): in-switch reduction for the gradient all-reduce,
        # multicast stores for the parameter broadcast, one-round-trip barrier. DGLB200_NVLS=0 keeps the CUDA-IPC path.
        self.heap = None
        if self.world_size > 1:
            import os as _os
            if _os.environ.get("DGLB200_NVLS", "1") != "0":
                try:
                    from .symm import SymmetricHeap
                    heap = SymmetricHeap(group=self.group, device=self.device)
                    if heap.vmm_ok and heap.mc_ok:
                        flags = heap.alloc(256)
                        if flags.mc_ptr:
                            self.heap, self._mc_flags = heap, flags
                            self._mc_epoch = torch.zeros(8, dtype=torch.int64, device=self.device)
                except Exception as e:  # noqa: BLE001 - any bootstrap problem: stay on the IPC path
                    self.heap, self._nvls_error = None, str(e)
            dist.barrier(group=self.group)

    @property
    def nvls(self):
        """True when gradient / parameter buffers can live in multicast-mapped symmetric memory."""
        return self.heap is not None

    def nvls_barrier(self, channel=0):
        """Barrier through the multicast counter: one ``multimem.red`` per rank + local spinning (graph-capturable)."""
        self.lib.fn("dglb200_symm_barrier", "pppipdp")(
            self._mc_flags.mc_ptr + 4 * channel, self._mc_flags.local_ptr + 4 * channel,
            self._mc_epoch.data_ptr() + 8 * channel, self.world_size, self._error.data_ptr(), self.timeout_s,
            loader.stream_ptr())

    def close(self):
        """Collective teardown: every rank unmaps its peers' buffers, then (after a barrier) frees its own allocations.
        Views handed out by ``SymmetricBuffer.view`` must not be used afterwards."""
        if getattr(self, "_closed", False):
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        if self.world_size > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        bufs = list(self._bufs) + [self._flags]
        for b in bufs:
            b._close_peer_mappings()
        if self.world_size > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        for b in bufs:
            b._free_local()
        self._bufs = []

    # ------------------------------------------------------------------ memory
    def alloc(self, nbytes):
        b = SymmetricBuffer(self, nbytes)
        self._bufs.append(b)
        return b

    def grad_buffer(self, numel):
        self._grad = self.alloc(int(numel) * 4)
        return self._grad.view(torch.float32)[:numel]

    def halo_plan(self, indptr, indices_global, shard_rows):
        """Build the static halo-exchange plan of this rank's partition (see ``HaloPlan``)."""
        return HaloPlan(self, indptr, indices_global, shard_rows)

    def peer_grad_ptrs(self):
        return self._grad.ptr_array

    # ------------------------------------------------------------------ collectives
    def barrier(self, channel=0):
        """Device-side barrier enqueued on the current stream (no host sync; CUDA-graph capturable)."""
        flag_ptrs = (ctypes.c_void_p * self.world_size)(*[p + channel * 64 for p in self._flags.ptrs])
        self._keep = flag_ptrs
        self.lib.fn("dglb200_comm_barrier", "piippdp")(
            ctypes.cast(flag_ptrs, ctypes.c_void_p), self.rank, self.world_size,
            self._epoch.data_ptr() + 8 * channel, self._error.data_ptr(), self.timeout_s, loader.stream_ptr())

    def publish_and_barrier(self, step_dev=None):
        self.barrier(0)

    def check_error(self):
        if int(self._error.item()) != 0:
            raise DGLError("device-side barrier timed out (a peer rank is stuck or dead)")

    def all_reduce(self, buf: SymmetricBuffer, out, numel, scale=1.0, out_bf16=None, sync=True):
        """out = scale * sum over ranks of ``buf`` (fp32). One-shot: every rank pulls every peer's copy
        with 16-byte P2P loads; ``sync`` adds the barriers that order it against producers/consumers."""
        if sync:
            self.barrier(1)
        self.lib.fn("dglb200_comm_allreduce", "pipplfp")(
            buf.ptr_table(), self.world_size, out.data_ptr(), 0 if out_bf16 is None else out_bf16.data_ptr(),
            int(numel), float(scale), loader.stream_ptr())
        if sync:
            self.barrier(2)

    TWO_SHOT_MIN_BYTES = 2 << 20

    def fused_allreduce_adam(self, param, m, v, shadow, step_dev, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                             weight_decay=0.0, two_shot=None):
        """Gradient all-reduce FUSED with the Adam update (+ bf16 shadow weights). Gradients must already sit in
        ``grad_buffer``. No NCCL call, graph-capturable.
          * one-shot (small models, latency-bound): every rank pulls all peers' gradients inside the Adam kernel;
          * two-shot (>= 2 MiB of gradients on >= 4 ranks, bandwidth-bound): each rank reduces + updates only ITS
            slice (reduce-scatter fused with Adam; optimizer state is effectively sharded) and publishes the new fp32
            values in a symmetric buffer, then pulls the other ranks' slices (all-gather fused with the bf16 cast) -
            2(W-1)/W x bytes per GPU instead of (W-1) x bytes."""
        numel = int(param.numel())
        W = self.world_size
        if two_shot is None:
            two_shot = W >= 4 and numel * 4 >= self.TWO_SHOT_MIN_BYTES
        peers = ctypes.cast(self._grad.ptr_array, ctypes.c_void_p)
        sh = 0 if shadow is None else shadow.data_ptr()
        self.barrier(0)
        if not two_shot or W == 1:
            self.lib.fn("dglb200_engine_allreduce_adam", "ppippplfffffpppp")(
                param.data_ptr(), peers, W, m.data_ptr(), v.data_ptr(), sh, numel, lr, betas[0], betas[1], eps,
                weight_decay, step_dev.data_ptr(), 0, 0, loader.stream_ptr())
            self.barrier(1)
            return
        if getattr(self, "_param_pub", None) is None or self._param_pub.nbytes < numel * 4:
            self._param_pub = self.alloc(numel * 4)
        slice_ = -(-numel // W)
        slice_ = (slice_ + 3) // 4 * 4
        self.lib.fn("dglb200_engine_allreduce_adam_rs", "ppiipppllfffffppp")(
            param.data_ptr(), peers, W, self.rank, m.data_ptr(), v.data_ptr(), sh, numel, slice_, lr, betas[0], betas[1],
            eps, weight_decay, step_dev.data_ptr(), self._param_pub.local_ptr, loader.stream_ptr())
        self.barrier(1)
        self.lib.fn("dglb200_engine_param_allgather", "ppiipllpp")(
            param.data_ptr(), self._param_pub.ptr_table(), W, self.rank, sh, numel, slice_, step_dev.data_ptr(),
            loader.stream_ptr())
        self.barrier(2)   # nobody overwrites its published slice / gradients before every rank has pulled them

    def halo_spmm(self, indptr, indices, shards: SymmetricBuffer, shard_rows, feat_dim, dtype, out, mean=True,
                  edge_weight=None):
        """Aggregation over a range-partitioned feature table with remote rows fetched over NVLink inside
        the kernel (no materialised halo exchange). ``indices`` hold GLOBAL source ids."""
        self.lib.fn("dglb200_comm_halo_spmm", "iipplpilipipp")(
            dtype_code(dtype), id_code(indptr.dtype), indptr.data_ptr(), indices.data_ptr(), indptr.shape[0] - 1,
            shards.ptr_table(), self.world_size, int(shard_rows), int(feat_dim), out.data_ptr(), 1 if mean else 0,
            0 if edge_weight is None else edge_weight.data_ptr(), loader.stream_ptr())
        return out

    def halo_spmm_backward(self, indptr, indices, dshards: SymmetricBuffer, shard_rows, feat_dim, dout, mean=True):
        self.lib.fn("dglb200_comm_halo_spmm_bwd", "iipplpilipip")(
            dtype_code(dout.dtype), id_code(indptr.dtype), indptr.data_ptr(), indices.data_ptr(),
            indptr.shape[0] - 1, dshards.ptr_table(), self.world_size, int(shard_rows), int(feat_dim),
            dout.data_ptr(), 1 if mean else 0, loader.stream_ptr())

    def gather(self, shards: SymmetricBuffer, shard_rows, idx, out):
        row_bytes = out.element_size() * int(out[0].numel()) if out.shape[0] else 0
        if out.shape[0]:
            self.lib.fn("dglb200_comm_p2p_gather", "ipilplplp")(
                id_code(idx.dtype), shards.ptr_table(), self.world_size, int(shard_rows), idx.data_ptr(),
                idx.shape[0], out.data_ptr(), row_bytes, loader.stream_ptr())
        return out

    def scatter_add(self, shards: SymmetricBuffer, shard_rows, idx, vals):
        if idx.shape[0]:
            self.lib.fn("dglb200_comm_p2p_scatter_add", "ipilplpip")(
                id_code(idx.dtype), shards.ptr_table(), self.world_size, int(shard_rows), idx.data_ptr(),
                idx.shape[0], vals.data_ptr(), int(vals.shape[1]), loader.stream_ptr())


class HaloPlan:
    """Static halo-exchange plan of one partition (the graph structure does not change between layers / steps).

    The reference moves halo features with collectives before every layer (DistDGL pulls them through the KVStore,
    GraphBolt's CooperativeConv does an all-to-all, python/dgl/graphbolt/impl/cooperative_conv.py:27-57); an NCCL
    all-gather of the whole table is the dense equivalent. Here the plan
      * deduplicates the remote source ids once: every halo row crosses NVLink exactly once per aggregation;
      * relabels the sources into the local table ``[local rows | halo rows]``;
      * orders the destination rows INTERIOR first (all sources local), then BOUNDARY rows,
    and ``spmm`` runs ONE cooperative kernel (csrc/cuda/comm.cu ``halo_pipe_kernel``) in which some warps of every
    CTA pull the halo chunks from the owners' HBM with peer loads while all other warps aggregate the interior rows;
    boundary rows start as soon as the device-wide chunk counter says the halo is complete - no NCCL call, no second
    launch, no host synchronisation.
    """

    MAX_CHUNKS = 2048

    def __init__(self, comm: "P2PComm", indptr, indices_global, shard_rows, puller_warps=None):
        self.comm, self.rows = comm, int(shard_rows)
        dev = comm.device
        self.indptr = indptr.contiguous()
        idx = indices_global.long()
        lo = comm.rank * self.rows
        remote = (idx < lo) | (idx >= lo + self.rows)
        halo = torch.unique(idx[remote])                      # sorted => grouped by owner (range partition)
        self.halo_ids = halo.to(indices_global.dtype).contiguous()
        self.n_halo = int(halo.shape[0])
        ext = idx - lo
        if self.n_halo:
            ext[remote] = self.rows + torch.searchsorted(halo, idx[remote])
        self.indices = ext.to(indices_global.dtype).contiguous()
        n = int(indptr.shape[0]) - 1
        has_remote = torch.zeros(n, dtype=torch.bool, device=dev)
        if idx.numel():
            deg = (indptr[1:] - indptr[:-1]).long()
            row_of = torch.repeat_interleave(torch.arange(n, device=dev), deg)
            has_remote[row_of[remote]] = True
        self.order = torch.sort(has_remote.to(torch.uint8), stable=True)[1].contiguous()   # interior rows first
        self.n_interior = n - int(has_remote.sum())
        m = torch.tensor([self.n_halo], device=dev)
        if comm.world_size > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=comm.group)
        self.halo_cap = int(m.item())
        self.chunk_rows = max(64, -(-self.n_halo // self.MAX_CHUNKS))
        self.counters = torch.zeros(4, dtype=torch.int32, device=dev)
        # measured on 2xB200 (profiles/comm_2gpu.json): with (almost) no interior rows every warp should pull first
        self.puller_warps = int(puller_warps) if puller_warps else (8 if self.n_interior * 20 < n else 4)

    # ---- two-phase formulation: bulk halo pull on a side stream || balanced g-SpMM of the interior rows, then the boundary rows
    def _split(self):
        """Row-permuted CSC (interior rows first) cut into an interior and a boundary matrix over the ``[local | halo]`` table."""
        if getattr(self, "_split_cache", None) is None:
            dev = self.indptr.device
            n = int(self.order.shape[0])
            ip = self.indptr.long()
            deg = ip[1:] - ip[:-1]
            order = self.order.long()
            deg_p = deg[order]
            ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            ptr[1:] = torch.cumsum(deg_p, 0)
            total = int(ptr[-1])
            owner = torch.repeat_interleave(torch.arange(n, device=dev), deg_p, output_size=total)
            pos = ip[:-1][order][owner] + (torch.arange(total, device=dev) - ptr[:-1][owner])
            idx = self.indices[pos].contiguous()
            cut = int(ptr[self.n_interior])
            it = self.indptr.dtype
            self._split_cache = dict(
                interior=(ptr[: self.n_interior + 1].to(it).contiguous(), idx[:cut].contiguous(), None),
                boundary=((ptr[self.n_interior:] - cut).to(it).contiguous(), idx[cut:].contiguous(), None),
                inv_order=torch.argsort(order), inv_deg=(1.0 / deg.clamp(min=1).to(torch.float32)).unsqueeze(1),
                side=torch.cuda.Stream(device=dev))
        return self._split_cache

    def spmm_balanced(self, buf, table, out, mean=True):
        """Same contract as ``spmm``. The halo rows are copied out of the owners' shards by ONE gather kernel on a side
        stream (sorted ids: sequential reads per owner, every row crosses NVLink once) while the main stream aggregates the
        interior rows with the load-balanced g-SpMM (csrc/cuda/spmm.cu: row-parallel vector kernel + edge-balanced items for
        hub rows); the boundary rows follow once the pull has landed. On hub-heavy graphs this is several times faster than
        the one-warp-per-row interior loop of ``halo_pipe_kernel`` (measured: products RMAT, 1 GPU, F = 256: 57 ms -> see
        profiles/partitioned.md)."""
        from .._sparse_ops import _gspmm_balanced
        sp = self._split()
        F = int(table.shape[1])
        n, ni = self.rows, self.n_interior
        main = torch.cuda.current_stream(table.device)
        if self.n_halo:
            sp["side"].wait_stream(main)
            with torch.cuda.stream(sp["side"]):
                self.comm.gather(buf, self.rows, self.halo_ids, table[n: n + self.n_halo])
        staged = torch.empty((n, F), dtype=table.dtype, device=table.device)
        if ni:
            _gspmm_balanced(sp["interior"], table, None, ni, F, out=staged[:ni])
        if self.n_halo:
            main.wait_stream(sp["side"])
        if n - ni:
            _gspmm_balanced(sp["boundary"], table, None, n - ni, F, out=staged[ni:])
        torch.index_select(staged, 0, sp["inv_order"], out=out)
        if mean:
            out.mul_(sp["inv_deg"].to(out.dtype))
        return out

    def alloc_features(self, feat_dim, dtype):
        """Symmetric feature buffer with room for the halo rows behind the local shard. Returns (buffer, table);
        ``table[:rows]`` is this rank's shard (what the peers read), ``table[rows:]`` receives the halo rows."""
        es = torch.empty(0, dtype=dtype).element_size()
        buf = self.comm.alloc((self.rows + self.halo_cap) * feat_dim * es)
        table = buf.view(dtype)[: (self.rows + self.halo_cap) * feat_dim].view(self.rows + self.halo_cap, feat_dim)
        return buf, table

    def spmm(self, buf, table, out, mean=True):
        """out[d] = sum / mean over the in-neighbours of d of the (distributed) feature rows. The caller must make sure
        every rank has finished writing its local rows (``comm.barrier``) before calling, and must not overwrite them
        before every rank has finished (another ``comm.barrier``)."""
        F = int(table.shape[1])
        self.comm.lib.fn("dglb200_comm_halo_pipe", "iipppll" "ppil" "pliip" "ipipp")(
            dtype_code(table.dtype), id_code(self.indptr.dtype), self.indptr.data_ptr(), self.indices.data_ptr(),
            self.order.data_ptr(), int(self.order.shape[0]), self.n_interior,
            table.data_ptr(), buf.ptr_table(), self.comm.world_size, self.rows,
            self.halo_ids.data_ptr(), self.n_halo, self.chunk_rows, self.puller_warps, self.counters.data_ptr(),
            F, out.data_ptr(), 1 if mean else 0, self.comm._error.data_ptr(), loader.stream_ptr())
        return out
