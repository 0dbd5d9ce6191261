"""GraphSAGE: the DGL-style module and the fused B200 training engine.

``GraphSAGE`` is the reference model of /root/reference/examples/multigpu/node_classification_sage.py:61-80
(3x SAGEConv-mean, ReLU, dropout 0.5) written against the public ``dgl_b200.nn`` API.

``FusedSAGETrainer`` is the B200-first execution of the same training step: sampling, block
construction, aggregation, GEMMs, loss, backward, gradient all-reduce and Adam run from ONE CUDA
graph with no host synchronisation (see csrc/cuda/engine.cu for the design notes). Mini-batches are
mathematically the same computation as DataLoader + NeighborSampler + SAGEConv('mean') + Adam.
"""
from __future__ import annotations

import ctypes
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .._native import loader
from ..base import DGLError, dtype_code, id_code
from ..nn import SAGEConv

__all__ = ["GraphSAGE", "FusedSAGETrainer"]


class GraphSAGE(nn.Module):
    """Multi-layer GraphSAGE-mean on message-flow graphs."""

    def __init__(self, in_size, hid_size, out_size, num_layers=3, dropout=0.5, aggregator="mean"):
        super().__init__()
        self.layers = nn.ModuleList()
        sizes = [in_size] + [hid_size] * (num_layers - 1) + [out_size]
        for i in range(num_layers):
            self.layers.append(SAGEConv(sizes[i], sizes[i + 1], aggregator))
        self.dropout = nn.Dropout(dropout)
        self.hid_size = hid_size
        self.out_size = out_size

    def forward(self, blocks, x):
        h = x
        for l, (layer, block) in enumerate(zip(self.layers, blocks)):
            h = layer(block, h)
            if l != len(self.layers) - 1:
                h = F.relu(h)
                h = self.dropout(h)
        return h

    def full_forward(self, g, x):
        h = x
        for l, layer in enumerate(self.layers):
            h = layer(g, h)
            if l != len(self.layers) - 1:
                h = F.relu(h)
        return h


def _p(t):
    return 0 if t is None else t.data_ptr()


def _pad8(n):
    return (n + 7) // 8 * 8


class FusedSAGETrainer:
    """Whole-step fused GraphSAGE trainer.

    Parameters
    ----------
    graph : DGLGraph on the GPU (homogeneous). Its CSC is used (in-neighbour sampling).
    feat : [N, F] node features (GPU). Stored in ``dtype`` (bf16 by default).
    labels : [N] class ids (GPU).
    fanouts : list, ``fanouts[i]`` = fan-out of layer i (same convention as NeighborSampler).
    comm : optional :class:`dgl_b200.parallel.P2PComm` for multi-GPU data-parallel training; gradients
        are averaged by a peer-memory kernel fused with the Adam update (no NCCL call).
    """

    def __init__(self, graph, feat, labels, num_classes, hidden=256, fanouts=(10, 10, 10), batch_size=1024, lr=1e-3,
                 weight_decay=5e-4, dropout=0.5, dtype=torch.bfloat16, seed=0, comm=None, use_cuda_graph=True,
                 betas=(0.9, 0.999), eps=1e-8, calibrate=True, cap_margin=1.5, use_tc=None):
        if not feat.is_cuda:
            raise DGLError("FusedSAGETrainer needs CUDA tensors (B200); use GraphSAGE + DataLoader on CPU")
        self.dev = feat.device
        self.lib = loader.cuda_lib()
        indptr, indices, _ = graph._graphs[0].csc()
        self._graph_ref = graph                    # used by predict() (generic sampler + module path)
        E = indices.shape[0]
        N = graph.num_nodes()
        if E >= 2 ** 31 or N >= 2 ** 31:
            raise DGLError("fused engine uses int32 ids; graph too large for one shard")
        self.indptr = indptr.to(torch.int32).contiguous()
        self.indices = indices.to(torch.int32).contiguous()
        self.N, self.E = N, E
        self.dtype = dtype
        # tensor-core path (tcgen05 fused gather+GEMM kernel): bf16, hidden % 32 == 0, feature rows padded to a
        # multiple of 64 elements (zero columns: mathematically identical, their weights receive zero gradient)
        tc_ok = dtype in (torch.bfloat16, torch.float32) and hidden in (128, 256) and num_classes <= 64
        if use_tc is None:
            from ..global_config import get_gemm_path
            path = get_gemm_path()                      # DGLB200_GEMM = auto | tc | cublas
            use_tc = tc_ok
            if path == "cublas":
                use_tc = False
            elif path == "tc" and not use_tc:
                raise ValueError("DGLB200_GEMM=tc needs hidden in {128, 256} and <= 64 classes")
        elif use_tc and not tc_ok:
            raise ValueError("the tcgen05 engine needs hidden in {128, 256} and <= 64 classes")
        self.use_tc = bool(use_tc)
        # bf16: kind::f16 MMAs on bf16 storage. fp32: fp32 storage + 3xTF32 split products (fp32-level accuracy, see
        # csrc/cuda/tc_common.cuh) - the engine never drops below the precision the caller asked for.
        self.precision_note = (
            ("bf16 storage/operands, fp32 accumulate (tcgen05 kind::f16)" if dtype == torch.bfloat16 else
             "fp32 storage, fp32 accumulate, 3xTF32 error-compensated tcgen05 products (fp32-equivalent)")
            if self.use_tc else ("bf16 operands via cuBLAS" if dtype == torch.bfloat16 else "fp32 via cuBLAS (TF32 off)"))
        self.F_orig = feat.shape[1]
        if self.use_tc and feat.shape[1] % 64:
            Fp = (feat.shape[1] + 63) // 64 * 64
            padded = torch.zeros((feat.shape[0], Fp), dtype=dtype, device=feat.device)
            padded[:, : feat.shape[1]] = feat.to(dtype)
            self.feat = padded
        else:
            self.feat = feat.to(dtype).contiguous()
        feat = self.feat
        self.labels = labels.to(torch.int32).contiguous()
        self.F_in = feat.shape[1]
        self.H = hidden
        self.C = num_classes
        self.L = len(fanouts)
        self.fanouts = list(fanouts)
        self.B = batch_size
        self.lr, self.wd, self.p_drop = lr, weight_decay, dropout
        self.betas, self.eps = betas, eps
        self.seed = seed
        # comm = P2PComm            -> gradients averaged by the peer-memory kernel fused with Adam (the product path)
        # comm = "nccl"             -> BASELINE: torch.distributed all_reduce (NCCL) of the flat gradient, then the same
        #                              Adam kernel on the local buffer (also selected by DGLB200_COMM=nccl when the
        #                              process group is initialised and no P2PComm is given)
        self.nccl_group = None
        if comm is None:
            from ..global_config import get_comm_path
            import torch.distributed as _dist
            if get_comm_path() == "nccl" and _dist.is_available() and _dist.is_initialized() and _dist.get_world_size() > 1:
                comm = "nccl"
        if isinstance(comm, str):
            if comm != "nccl":
                raise DGLError("comm must be a P2PComm, 'nccl' or None")
            import torch.distributed as _dist
            self.nccl_group = _dist.group.WORLD
            self._nccl_world, self._nccl_rank = _dist.get_world_size(), _dist.get_rank()
            comm = None
        self.comm = comm
        if dtype not in (torch.bfloat16, torch.float32):
            raise DGLError("FusedSAGETrainer supports dtype bfloat16 (tensor-core path) or float32")
        self.world = 1 if comm is None else comm.world_size
        self.rank = 0 if comm is None else comm.rank
        if self.nccl_group is not None:
            self.rank = self._nccl_rank
        if self.F_in % 2 or hidden % 8:
            raise DGLError("feature sizes must be even (F_in) / multiples of 8 (hidden)")
        # ---- layer dims: layer l (1-based) maps in_dims[l-1] -> out_dims[l-1]
        self.in_dims = [self.F_in] + [hidden] * (self.L - 1)
        self.out_dims = [hidden] * (self.L - 1) + [num_classes]
        # ---- capacities: caps[l] = max |S_l|, S_L = seeds, S_{l-1} = unique(S_l U nbrs)
        caps = [0] * (self.L + 1)
        caps[self.L] = batch_size
        for l in range(self.L, 1, -1):
            caps[l - 1] = caps[l] * (1 + self.fanouts[l - 1])
        self.caps = caps
        self.max_caps = list(caps)
        self._alloc_params()
        if calibrate:
            self._calibrate_caps(cap_margin)
        self._alloc_buffers()
        self.use_graph = use_cuda_graph
        self.graph = None
        self._side_stream = torch.cuda.Stream(self.dev)
        self._zero_stream = torch.cuda.Stream(self.dev)
        self._probe_mm()
        self.gpu_launches_per_step = 0

    # ------------------------------------------------------------------ parameters
    def _alloc_params(self):
        dev = self.dev
        # Layers that run on the tcgen05 kernel keep their weight in the K-major layout the tensor core wants:
        # storage [n_pad, 2*in] (= W^T, output rows padded to the MMA N granularity, padding rows stay exactly 0).
        # self.W / self.Wb / self.gW are [2*in, out] *views* of that storage, so everything else is layout agnostic.
        self.tc_last = self.use_tc and self.C <= 64
        self.tc_layer = [self.use_tc and (l < self.L - 1 or self.tc_last) for l in range(self.L)]
        self.n_pad = [(64 if l == self.L - 1 else self.out_dims[l]) if self.tc_layer[l] else self.out_dims[l]
                      for l in range(self.L)]
        segs = []
        off = 0
        for l in range(self.L):
            k, n = 2 * self.in_dims[l], self.n_pad[l]
            segs.append(("W", l, off, (k, n)))
            off += _pad8(k * n)
            segs.append(("b", l, off, (n,)))
            off += _pad8(n)
        self.numel = off
        self.segs = segs
        # fp32 tensor-core operands: hi / lo planes of the TF32 split, refreshed by the optimizer epilogue
        split = self.use_tc and self.dtype == torch.float32
        self._nvls = None
        if self.comm is not None and getattr(self.comm, "nvls", False):
            # NVLS: gradients, parameters and their operand copies live in ONE multicast-mapped symmetric allocation:
            # the optimizer kernel reduces gradients in the switch (multimem.ld_reduce) and multicasts the update
            pad = lambda nbytes: (nbytes + 255) // 256 * 256  # noqa: E731
            o_grad, o_param = 0, pad(4 * (off + 8))
            o_sh = o_param + pad(4 * off)
            o_hi = o_sh + pad(2 * off)
            o_lo = o_hi + (pad(4 * off) if split else 0)
            total = o_lo + (pad(4 * off) if split else 0)
            buf = self.comm.heap.alloc(total)
            if buf.mc_ptr:
                self._nvls = dict(buf=buf, grad=buf.mc_ptr + o_grad, param=buf.mc_ptr + o_param, shadow=buf.mc_ptr + o_sh,
                                  hi=buf.mc_ptr + o_hi if split else 0, lo=buf.mc_ptr + o_lo if split else 0)
                self.zero_block = buf.view(torch.float32, o_grad, off + 8)
                self.param = buf.view(torch.float32, o_param, off)
                self.shadow = buf.view(torch.bfloat16, o_sh, off)
                self.shadow_hi = buf.view(torch.float32, o_hi, off) if split else None
                self.shadow_lo = buf.view(torch.float32, o_lo, off) if split else None
        if self._nvls is None:
            self.param = torch.zeros(off, dtype=torch.float32, device=dev)
            self.shadow = torch.zeros(off, dtype=torch.bfloat16, device=dev)
            self.shadow_hi = torch.zeros(off, dtype=torch.float32, device=dev) if split else None
            self.shadow_lo = torch.zeros(off, dtype=torch.float32, device=dev) if split else None
            # flat gradient + 8 trailing words (loss, accuracy counter): one memset per step clears all of it
            self.zero_block = (torch.zeros(off + 8, dtype=torch.float32, device=dev) if self.comm is None
                               else self.comm.grad_buffer(off + 8))
        self.adam_m = torch.zeros_like(self.param)
        self.adam_v = torch.zeros_like(self.param)
        self.grad = self.zero_block[:off]
        self.loss_dev = self.zero_block[off:off + 1]
        self.correct_dev = self.zero_block[off + 1:off + 2].view(torch.int32)
        gen = torch.Generator(device="cpu")
        gen.manual_seed(self.seed)
        gain = math.sqrt(2.0)
        self.W, self.b, self.Wb, self.gW, self.gb = [], [], [], [], []
        self.Wt = [None] * self.L      # [n_pad, 2*in] K-major operand of the fused layer kernel (bf16, or fp32 hi plane)
        self.Wt_lo = [None] * self.L   # fp32 only: lo plane
        self.gWt = [None] * self.L     # gradient in the same layout
        self.b_pad = [None] * self.L   # bias incl. padding entries
        for kind, l, o, shape in segs:
            n = int(torch.tensor(shape).prod())
            n_out = self.out_dims[l]
            if kind == "W":
                k2, npad = shape

                def wview(buf, o=o, n=n, k2=k2, npad=npad, l=l, n_out=n_out):
                    if self.tc_layer[l]:
                        return buf[o:o + n].view(npad, k2).t()[:, :n_out]  # [2*in, out] strided view of W^T storage
                    return buf[o:o + n].view(k2, npad)

                view = wview(self.param)
                k_in = self.in_dims[l]
                # same init as SAGEConv: xavier_uniform(gain=relu) on fc_self.weight / fc_neigh.weight [out, in]
                bound = gain * math.sqrt(6.0 / (k_in + n_out))
                w_self = (torch.rand(n_out, k_in, generator=gen) * 2 - 1) * bound
                w_neigh = (torch.rand(n_out, k_in, generator=gen) * 2 - 1) * bound
                if l == 0 and self.F_orig != self.F_in:
                    w_self[:, self.F_orig:] = 0
                    w_neigh[:, self.F_orig:] = 0
                view.copy_(torch.cat([w_self.t(), w_neigh.t()], 0).to(dev))
                self.W.append(view)
                self.Wb.append(wview(self.shadow) if self.dtype == torch.bfloat16 else view)
                self.gW.append(wview(self.grad))
                if self.tc_layer[l]:
                    if split:
                        self.Wt[l] = self.shadow_hi[o:o + n].view(npad, k2)
                        self.Wt_lo[l] = self.shadow_lo[o:o + n].view(npad, k2)
                    else:
                        self.Wt[l] = self.shadow[o:o + n].view(npad, k2)
                    self.gWt[l] = self.grad[o:o + n].view(npad, k2)
            else:
                self.b.append(self.param[o:o + n][:n_out])
                self.gb.append(self.grad[o:o + n][:n_out])
                self.b_pad[l] = self.param[o:o + n]
                self.gb_pad = getattr(self, "gb_pad", [None] * self.L)
                self.gb_pad[l] = self.grad[o:o + n]
        self._refresh_shadows()
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)

    def _refresh_shadows(self):
        """Operand copies of the fp32 master weights: bf16 shadow, or the TF32 hi / lo planes."""
        self.shadow.copy_(self.param)
        if self.shadow_hi is not None:
            hi = (self.param.view(torch.int32) + 0x1000).bitwise_and(-8192).view(torch.float32)  # round-to-nearest TF32
            self.shadow_hi.copy_(hi)
            self.shadow_lo.copy_(self.param - hi)

    def state_dict(self):
        return {"param": self.param.clone(), "m": self.adam_m.clone(), "v": self.adam_v.clone(),
                "step": self.step_dev.clone()}

    def load_state_dict(self, sd):
        self.param.copy_(sd["param"]); self.adam_m.copy_(sd["m"]); self.adam_v.copy_(sd["v"])
        self.step_dev.copy_(sd["step"]); self._refresh_shadows()

    def export_to_module(self, module: GraphSAGE):
        """Copy the trained weights into a ``GraphSAGE`` module (fc_self / fc_neigh / bias)."""
        with torch.no_grad():
            for l, layer in enumerate(module.layers):
                k = self.in_dims[l]
                ko = self.F_orig if l == 0 else k
                layer.fc_self.weight.copy_(self.W[l][:ko].t())
                layer.fc_neigh.weight.copy_(self.W[l][k:k + ko].t())
                layer.fc_self.bias.copy_(self.b[l])

    def import_from_module(self, module: GraphSAGE):
        with torch.no_grad():
            for l, layer in enumerate(module.layers):
                k = self.in_dims[l]
                ko = self.F_orig if l == 0 else k
                self.W[l].zero_()
                self.W[l][:ko].copy_(layer.fc_self.weight.t())
                self.W[l][k:k + ko].copy_(layer.fc_neigh.weight.t())
                self.b[l].copy_(layer.fc_self.bias)
            self._refresh_shadows()

    # ------------------------------------------------------------------ capacity calibration
    def _calibrate_caps(self, margin, trials=24):
        """Static buffer capacities default to the worst case B*prod(1+fanout). On real graphs the number of
        unique frontier nodes is far smaller, so we sample a few random batches (sampling + compaction only),
        take the largest frontier seen and keep ``margin`` headroom (rounded to 128 rows). A device-side
        overflow flag guards the bound: if a later batch exceeds it the step is flagged, capacities are
        raised to the worst case and the graph is recaptured (see ``check_overflow``)."""
        from ..ops.cuda_struct import cuda_unique_compact
        from ..sampling.neighbor import sample_rel  # noqa: F401
        L = self.L
        gen = torch.Generator(device="cpu")
        gen.manual_seed(self.seed + 17)
        seen = [0] * (L + 1)
        for _ in range(trials):
            cur = torch.randint(0, self.N, (self.B,), generator=gen).to(self.dev, torch.int32)
            for l in range(L, 1, -1):
                f = self.fanouts[l - 1]
                n = cur.shape[0]
                nbr = torch.empty((n, f), dtype=torch.int32, device=self.dev)
                cnt = torch.empty(n, dtype=torch.int32, device=self.dev)
                self.lib.fn("dglb200_cuda_sample_ell", "ippiplpilpppppip")(
                    id_code(torch.int32), _p(self.indptr), _p(self.indices), 0, _p(cur), n, 0, f, 12345 + l, 0,
                    _p(nbr), 0, _p(cnt), 0, 1, loader.stream_ptr())
                cur, _ = cuda_unique_compact(cur, nbr.reshape(-1))
                seen[l - 1] = max(seen[l - 1], int(cur.shape[0]))
        for l in range(1, L):
            want = int(math.ceil(seen[l] * margin / 128.0)) * 128
            self.caps[l] = max(128, min(self.max_caps[l], want))
        self.calibrated = {"seen_max": seen[1:L], "caps": self.caps[1:L]}

    def check_overflow(self):
        """True if a step exceeded the calibrated capacities (that step dropped frontier nodes). The engine
        then falls back to worst-case capacities and recaptures."""
        if int(self.overflow.item()) == 0:
            return False
        self.overflow.zero_()
        self.caps = list(self.max_caps)
        self._alloc_buffers()
        self.graph = None
        return True

    # ------------------------------------------------------------------ buffers
    def _alloc_buffers(self):
        dev, L = self.dev, self.L
        i32 = torch.int32
        # host->device staging: ring of pinned slots [B seeds | int64 count] + one device landing buffer
        self.RING = 4
        self.stage_pinned = [torch.full((self.B + 2,), -1, dtype=i32).pin_memory() for _ in range(self.RING)]
        self.stage_events = [None] * self.RING
        self.stage_dev = torch.full((self.B + 2,), -1, dtype=i32, device=dev)
        self._slot = 0
        self.loss_pinned = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.S = [None] * (L + 1)      # node sets (global ids), S[L] = seeds
        self.n = [None] * (L + 1)      # device counts
        self.nbr = [None] * (L + 1)    # ELL neighbour ids for layer l (global for l=1 else local in S[l-1])
        self.nbr_g = [None] * (L + 1)  # global ids before relabel
        self.cnt = [None] * (L + 1)
        self.ws = [None] * (L + 1)     # compaction workspaces
        for l in range(L, 0, -1):
            cap, f = self.caps[l], self.fanouts[l - 1]
            if l == L:
                self.S[l] = self.stage_dev[: self.B]
                self.n[l] = self.stage_dev[self.B:].view(torch.int64)
            else:
                self.n[l] = torch.zeros(1, dtype=torch.int64, device=dev)
            self.nbr_g[l] = torch.empty((cap, f), dtype=i32, device=dev)
            self.cnt[l] = torch.empty(cap, dtype=i32, device=dev)
            if l > 1:
                N = cap + cap * f
                out_cap = self.caps[l - 1]
                hcap = 16
                while hcap < 2 * N + 2:
                    hcap <<= 1
                # hash table of the cooperative frontier-compaction kernel (left clean by every call)
                self.ws[l] = dict(keys=torch.full((hcap,), -1, dtype=i32, device=dev),
                                  vals=torch.full((hcap,), -2, dtype=i32, device=dev), cap=hcap,
                                  slot=torch.empty(N, dtype=i32, device=dev),
                                  counter=torch.zeros(1, dtype=i32, device=dev))
                self.S[l - 1] = torch.full((out_cap,), -1, dtype=i32, device=dev)
                self.nbr[l] = torch.empty((cap, f), dtype=i32, device=dev)
            else:
                self.nbr[l] = self.nbr_g[l]
        if not hasattr(self, "overflow"):   # survive a re-allocation after a capacity overflow
            self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
            self.edges_total = torch.zeros(16, dtype=torch.int64, device=dev)  # running count of sampled edges
            self.overflow_steps = 0
        # activations
        self.A = [None] * (L + 1)
        self.dH = [None] * (L + 1)
        for l in range(1, L + 1):
            self.A[l] = torch.zeros((self.caps[l], 2 * self.in_dims[l - 1]), dtype=self.dtype, device=dev)
            if l < L:
                self.dH[l] = torch.zeros((self.caps[l], self.out_dims[l - 1]), dtype=torch.float32, device=dev)
        self.dlogits = torch.empty((self.B, self.C), dtype=torch.float32, device=dev)
        self.Hbuf = [None] * (L + 1)
        if self.use_tc:
            for l in range(1, L):
                self.Hbuf[l] = torch.zeros((self.caps[l], self.out_dims[l - 1]), dtype=self.dtype, device=dev)
        if self.tc_last:
            self.logits_pad = torch.zeros((self.B, 64), dtype=torch.float32, device=dev)
            # CE gradient in the operand format of the backward kernels: bf16, or the fp32 hi / lo planes
            self.dlogits16 = torch.zeros((self.B, 64), dtype=self.dtype, device=dev)
            self.dlogits_lo = torch.zeros((self.B, 64), dtype=torch.float32, device=dev) if self.shadow_hi is not None else None
            # fp32: dlogits16 / dlogits_lo are the TF32 hi / lo planes (TMA operands of the dA kernel), dlogits_full the
            # unsplit value (the dW kernel splits its operands itself)
            self.dlogits_full = (torch.zeros((self.B, 64), dtype=torch.float32, device=dev)
                                 if self.shadow_hi is not None else self.dlogits16)
        # dZ of the hidden layers that feed a dA kernel (layers 2..L-1), written once by the dW kernel's producers
        self.dZ = [None] * (L + 1)
        self.dZ_lo = [None] * (L + 1)
        if self.use_tc:
            for l in range(2, L):
                self.dZ[l] = torch.zeros((self.caps[l], self.out_dims[l - 1]), dtype=self.dtype, device=dev)
                if self.shadow_hi is not None:
                    self.dZ_lo[l] = torch.zeros((self.caps[l], self.out_dims[l - 1]), dtype=torch.float32, device=dev)
            # all dH buffers in one allocation: ONE memset node per step clears them (side stream, off the critical path)
            tot = sum(self.caps[l] * self.out_dims[l - 1] for l in range(1, L))
            self.dH_all = torch.zeros(max(tot, 1), dtype=torch.float32, device=dev)
            o = 0
            for l in range(1, L):
                k = self.caps[l] * self.out_dims[l - 1]
                self.dH[l] = self.dH_all[o:o + k].view(self.caps[l], self.out_dims[l - 1])
                o += k
        thresh = min(255, int(min(self.p_drop, 1.0) * 256.0 + 0.5)) if self.p_drop > 0 else 0
        self.keep_scale = 256.0 / (256.0 - thresh)  # the kernels quantise the drop probability to 1/256

    def _probe_mm(self):
        """Can torch.mm emit an fp32 result from bf16 operands (keeps dW in full precision)?"""
        self.mm_fp32_out = False
        try:
            a = torch.zeros(8, 8, dtype=self.dtype, device=self.dev)
            r = torch.mm(a, a, out_dtype=torch.float32)
            self.mm_fp32_out = r.dtype == torch.float32
        except Exception:  # noqa: BLE001
            self.mm_fp32_out = False

    def _mm32(self, a, b):
        if a.dtype == torch.float32:
            return torch.mm(a, b)
        if self.mm_fp32_out:
            return torch.mm(a, b, out_dtype=torch.float32)
        return torch.mm(a, b).float()

    # ------------------------------------------------------------------ the step (device side)
    def _launch(self, name, sig, *args, kernels=1):
        self.lib.fn(name, sig + "p")(*args, loader.stream_ptr())
        self._launch_count += kernels
        self._my_kernels += kernels

    def _step_impl(self):
        L, dt, it = self.L, dtype_code(self.dtype), id_code(torch.int32)
        self._launch_count = 0
        self._my_kernels = 0
        if self.use_tc and L > 1:
            # the fp32 gradient buffers the dA kernels scatter into: cleared by one memset that overlaps sampling + forward
            zs = self._zero_stream
            zs.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(zs):
                self.dH_all.zero_()
            self._launch_count += 1
        # ---------------- sampling + compaction, output layer first
        for l in range(L, 0, -1):
            cap, f = self.caps[l], self.fanouts[l - 1]
            self._launch("dglb200_cuda_sample_ell", "ippiplpilpppppi", it, _p(self.indptr), _p(self.indices), 0,
                         _p(self.S[l]), cap, _p(self.n[l]), f, self.seed * 7919 + l, _p(self.step_dev),
                         _p(self.nbr_g[l]), 0, _p(self.cnt[l]), _p(self.edges_total), (1 if l > 1 else 0) | 2)
            if l > 1:
                w = self.ws[l]
                self._launch("dglb200_frontier_compact", "pplpipplpplpppp", _p(self.S[l]), _p(self.n[l]), cap,
                             _p(self.nbr_g[l]), f, _p(w["keys"]), _p(w["vals"]), w["cap"], _p(w["slot"]),
                             _p(self.S[l - 1]), self.caps[l - 1], _p(self.nbr[l]), _p(self.n[l - 1]), _p(w["counter"]),
                             _p(self.overflow))
        # ---------------- forward
        Hs = [None] * (L + 1)
        X = self.feat
        for l in range(1, L + 1):
            cap, f, Fi = self.caps[l], self.fanouts[l - 1], self.in_dims[l - 1]
            if self.tc_layer[l - 1]:
                # gather + mean + concat + GEMM + bias (+ ReLU + dropout) in ONE tcgen05 kernel; the output layer
                # emits fp32 logits (bias included) into a 64-column padded buffer
                last = l == L
                self._launch("dglb200_engine_sage_layer_tc", "ippppp" + "l" + "iii" + "ppppp" + "iflpip", dt, _p(X),
                             _p(self.S[1]) if l == 1 else 0, _p(self.nbr[l]), _p(self.cnt[l]), _p(self.n[l]), cap, f, Fi,
                             self.n_pad[l - 1], _p(self.Wt[l - 1]), _p(self.Wt_lo[l - 1]), _p(self.b_pad[l - 1]),
                             _p(self.logits_pad if last else self.Hbuf[l]), _p(self.A[l]), 0 if last else 1,
                             0.0 if last else float(self.p_drop), self.seed * 104729 + l, _p(self.step_dev),
                             1 if last else 0, 0)
                if not last:
                    Hs[l] = self.Hbuf[l]
                    X = self.Hbuf[l]
                continue
            self._launch("dglb200_engine_agg_concat", "iippppplii" + "pi", dt, it, _p(X), _p(self.S[1]) if l == 1 else 0,
                         _p(self.nbr[l]), _p(self.cnt[l]), _p(self.n[l]), cap, f, Fi, _p(self.A[l]), 0)
            if l < L:
                Z = torch.mm(self.A[l], self.Wb[l - 1])
                self._launch_count += 1
                self._launch("dglb200_engine_bias_relu_dropout", "ippplifl" + "pi", dt, _p(Z), _p(self.b[l - 1]),
                             _p(self.n[l]), cap, self.out_dims[l - 1], float(self.p_drop), self.seed * 104729 + l,
                             _p(self.step_dev), 1)
                Hs[l] = Z
                X = Z
            else:
                logits = self._mm32(self.A[l], self.Wb[l - 1])
                self._launch_count += 1
        # one memset clears loss / accuracy counters and the whole flat gradient (they are adjacent views)
        self.zero_block.zero_()
        self._launch_count += 1
        if self.tc_last:
            split = self.dlogits_lo is not None
            self._launch("dglb200_engine_ce_loss2", "pipppplippipp" + "ppp", _p(self.logits_pad), 64, 0, _p(self.labels),
                         _p(self.S[L]), _p(self.n[L]), self.B, self.C, _p(self.dlogits_full) if split else 0,
                         0 if split else _p(self.dlogits16), 64, _p(self.gb[L - 1]),
                         _p(self.loss_dev), _p(self.correct_dev), _p(self.dlogits16) if split else 0, _p(self.dlogits_lo))
        else:
            self._launch("dglb200_engine_ce_loss2", "pipppplippipp" + "ppp", _p(logits), self.C, _p(self.b[L - 1]),
                         _p(self.labels), _p(self.S[L]), _p(self.n[L]), self.B, self.C, _p(self.dlogits), 0, 0,
                         _p(self.gb[L - 1]), _p(self.loss_dev), _p(self.correct_dev), 0, 0)
        # ---------------- backward
        # The weight-gradient GEMM of a layer does not feed the dA -> scatter -> dReLU chain that produces the next
        # layer's dZ, so it runs on a side stream (forked/joined with events, capturable) and overlaps with that chain.
        main = torch.cuda.current_stream(self.dev)
        side = self._side_stream
        if self.use_tc:
            self._backward_tc(main, side)
            self._finish_step()
            return
        dZ = self.dlogits
        keep = [dZ]  # every dZ stays referenced until the join: the allocator must not recycle a block the side stream reads
        for l in range(L, 0, -1):
            cap, f, Fi = self.caps[l], self.fanouts[l - 1], self.in_dims[l - 1]
            A = self.A[l]
            if self.tc_layer[l - 1]:
                # weights live as W^T [n_pad, 2*in]: dW^T = dZ^T A, dA = dZ W^T - no transposes materialised
                if l > 1:
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        self.gWt[l - 1].copy_(self._mm32(dZ.t(), A))
                else:
                    self.gWt[l - 1].copy_(self._mm32(dZ.t(), A))
                self._launch_count += 2
                dA = torch.mm(dZ, self.Wt[l - 1]) if l > 1 else None
                dA_dt = dt
            elif l == L:
                self.gW[l - 1].copy_(torch.mm(A.float().t(), dZ))
                self._launch_count += 3
                dA = torch.mm(dZ, self.W[l - 1].t()) if l > 1 else None
                dA_dt = dtype_code(torch.float32)
            else:
                self.gW[l - 1].copy_(self._mm32(A.t(), dZ))
                self._launch_count += 2
                dA = torch.mm(dZ, self.Wb[l - 1].t()) if l > 1 else None
                dA_dt = dt
            if l > 1:
                self._launch_count += 1
                dH = self.dH[l - 1]
                self._launch("dglb200_engine_zero_rows", "ppll", _p(dH), _p(self.n[l - 1]), self.caps[l - 1], Fi * 4)
                self._launch("dglb200_engine_scatter_grad", "iippppliip", dA_dt, it, _p(dA), _p(self.nbr[l]),
                             _p(self.cnt[l]), _p(self.n[l]), cap, f, Fi, _p(dH))
                dZn = torch.empty((self.caps[l - 1], Fi), dtype=self.dtype, device=self.dev)
                scale = 1.0 / (1.0 - self.p_drop) if self.p_drop > 0 else 1.0
                self._launch("dglb200_engine_relu_bwd", "ippplifpp", dt, _p(dH), _p(Hs[l - 1]), _p(self.n[l - 1]),
                             self.caps[l - 1], Fi, float(scale), _p(dZn), _p(self.gb[l - 2]))
                dZ = dZn
                keep.append(dZn)
        if any(self.tc_layer[1:]):  # the side stream only carries work (and joined the capture) on that branch
            main.wait_stream(side)
        self._finish_step()

    def _backward_tc(self, main, side):
        """Backward on tcgen05: per layer one weight-gradient kernel (dZ built on the fly from the incoming fp32
        gradient + ReLU/dropout mask, bias gradient in the same pass) and one input-gradient kernel whose epilogue IS
        the scatter into the layer below. No cuBLAS, no elementwise kernels."""
        L, dt = self.L, dtype_code(self.dtype)
        sig_dw = "ipppiifpiipl" + "pllpppi"
        sig_da = "ipppppppliiipi"
        # layer L: dZ = dlogits (CE kernel). Swapped roles (M side = A columns) because n_pad = 64 < one MMA M tile.
        side.wait_stream(main)
        with torch.cuda.stream(side):
            k2 = 2 * self.in_dims[L - 1]
            self._launch("dglb200_engine_sage_bwd_dw", sig_dw, dt, 0, 0, _p(self.A[L]), k2, k2, 1.0,
                         _p(self.dlogits_full), 64, 64, _p(self.n[L]), self.caps[L], _p(self.gWt[L - 1]), 1, k2, 0, 0, 0, 0)
        dz, dz_lo = self.dlogits16, self.dlogits_lo
        for l in range(L, 1, -1):
            Fi = self.in_dims[l - 1]
            if l == L:
                main.wait_stream(self._zero_stream)  # dH buffers were cleared off the critical path (see _step_impl)
            self._launch("dglb200_engine_sage_bwd_da", sig_da, dt, _p(dz), _p(dz_lo), _p(self.Wt[l - 1]),
                         _p(self.Wt_lo[l - 1]), _p(self.nbr[l]), _p(self.cnt[l]), _p(self.n[l]), self.caps[l],
                         self.fanouts[l - 1], self.n_pad[l - 1], Fi, _p(self.dH[l - 1]), 0)
            lo = l - 1  # weight gradient of the layer below; its dZ feeds the next dA kernel
            k2 = 2 * self.in_dims[lo - 1]
            self._launch("dglb200_engine_sage_bwd_dw", sig_dw, dt, _p(self.dH[lo]), _p(self.Hbuf[lo]), 0,
                         self.out_dims[lo - 1], self.out_dims[lo - 1], float(self.keep_scale), _p(self.A[lo]), k2, k2,
                         _p(self.n[lo]), self.caps[lo], _p(self.gWt[lo - 1]), k2, 1, _p(self.gb_pad[lo - 1]),
                         _p(self.dZ[lo]), _p(self.dZ_lo[lo]), 0)
            dz, dz_lo = self.dZ[lo], self.dZ_lo[lo]
        main.wait_stream(side)

    def _finish_step(self):
        # ---------------- gradient all-reduce fused with Adam
        if self._nvls is not None:
            # NVLS two-shot: in-switch reduction of this rank's slice, Adam on the slice, multicast of the new values
            nv = self._nvls
            self.comm.nvls_barrier(0)   # every rank's gradients are complete and visible
            self._launch("dglb200_symm_allreduce_adam", "pppppppp" + "lii" + "fffff" + "p", nv["grad"], nv["param"],
                         _p(self.param), _p(self.adam_m), _p(self.adam_v),
                         nv["shadow"] if self.dtype == torch.bfloat16 else 0, nv["hi"], nv["lo"], self.numel, self.world,
                         self.rank, float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                         float(self.wd), _p(self.step_dev), kernels=2)
            self.comm.nvls_barrier(1)   # every slice has landed everywhere before the next forward reads the weights
            self._launch_count += 2
            self._my_kernels += 2
            self.loss_pinned.copy_(self.loss_dev, non_blocking=True)
            self.gpu_launches_per_step = self._my_kernels
            self.total_launches_per_step = self._launch_count
            return
        if self.comm is not None:
            self.comm.barrier(0)  # every rank's gradients are complete and visible
            self._launch_count += 1
            self._my_kernels += 1
            peers = self.comm.peer_grad_ptrs()
        else:
            if self.nccl_group is not None:
                # baseline communication path: NCCL all-reduce (capturable), pre-scaled so the kernel's 1/world is a no-op
                import torch.distributed as _dist
                self.grad.mul_(1.0 / self._nccl_world)
                _dist.all_reduce(self.grad, group=self.nccl_group)
                self._launch_count += 2
            peers = (ctypes.c_void_p * 1)(self.grad.data_ptr())
        self._peers_keepalive = peers
        self._launch("dglb200_engine_allreduce_adam", "ppippplfffffppp", _p(self.param), ctypes.cast(peers, ctypes.c_void_p),
                     self.world, _p(self.adam_m), _p(self.adam_v), _p(self.shadow) if self.dtype == torch.bfloat16 else 0, self.numel, float(self.lr),
                     float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.wd), _p(self.step_dev),
                     _p(self.shadow_hi), _p(self.shadow_lo), kernels=2)
        if self.comm is not None:
            self.comm.barrier(1)  # all ranks finished reading before anyone rewrites its gradients
            self._launch_count += 1
            self._my_kernels += 1
        self.loss_pinned.copy_(self.loss_dev, non_blocking=True)
        self.gpu_launches_per_step = self._my_kernels
        self.total_launches_per_step = self._launch_count

    # ------------------------------------------------------------------ public API
    def _capture(self):
        # warm-up on a side stream (cuBLAS handles / workspaces), then capture
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        saved = self.state_dict()
        with torch.cuda.stream(s):
            for _ in range(2):
                self._step_impl()  # runs on whatever seeds are staged (valid ids: load_seeds precedes capture)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.load_state_dict(saved)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_impl()
        self.load_state_dict(saved)
        torch.cuda.synchronize(self.dev)

    def load_seeds(self, seeds):
        """Stage a batch of seed ids (CPU tensor/list, <= batch_size): write them to a pinned ring slot and
        enqueue the host->device copy on the current stream. Never blocks unless the ring wraps onto a
        slot whose copy has not executed yet."""
        n = len(seeds)
        if n > self.B:
            raise DGLError(f"batch of {n} seeds exceeds batch_size={self.B}")
        k = self._slot
        self._slot = (k + 1) % self.RING
        if self.stage_events[k] is not None:
            self.stage_events[k].synchronize()
        buf = self.stage_pinned[k]
        buf[:n] = torch.as_tensor(seeds, dtype=torch.int32)
        if n < self.B:
            buf[n:self.B] = -1
        buf[self.B:].view(torch.int64)[0] = n
        self.stage_dev.copy_(buf, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.stage_events[k] = ev

    def step_async(self, seeds):
        """Enqueue one training step: H2D of the seeds from pinned memory -> sample -> ... -> Adam -> D2H of
        the loss into pinned memory. Does not synchronise with the host."""
        self.load_seeds(seeds)
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._step_impl()

    def train_step(self, seeds):
        """One training step on ``seeds``; returns the mini-batch loss as a Python float (syncs)."""
        self.step_async(seeds)
        torch.cuda.current_stream(self.dev).synchronize()
        return float(self.loss_pinned[0])

    def check_step_overflow(self):
        """Synchronous variant for ``train_step`` loops: True if any step since the last check overflowed the
        calibrated frontier capacities (the engine then switches to worst-case capacities)."""
        if self.check_overflow():
            self.overflow_steps += 1
            return True
        return False

    def last_loss(self):
        return float(self.loss_pinned[0])

    def fit(self, batches, on_step=None):
        """Pipelined training over an iterable of seed batches; returns the list of per-step losses.

        Every step still copies its seeds from pinned host memory to the device and its loss (plus the frontier
        overflow flag) back into pinned host memory, but the host reads the result of step ``i`` only after it has
        enqueued step ``i + RING - 1``, so the CPU never idles the GPU (the reference's loop reads ``loss.item()``
        synchronously, examples/multigpu/node_classification_sage.py:229-256). ``on_step(i, loss)`` is called as losses
        arrive. A step whose sampled frontier exceeded the calibrated capacities trained on a truncated neighbourhood:
        it is counted in ``overflow_steps`` and the engine falls back to worst-case capacities before the next step."""
        import collections
        if not hasattr(self, "loss_ring"):
            self.loss_ring = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(self.RING)]
            self.ovf_ring = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(self.RING)]
        pending = collections.deque()
        losses = []
        state = {"overflow": False}

        def drain_one():
            k, ev = pending.popleft()
            ev.synchronize()
            losses.append(float(self.loss_ring[k][0]))
            if int(self.ovf_ring[k][0]) != 0:
                state["overflow"] = True
            if on_step is not None:
                on_step(len(losses) - 1, losses[-1])

        for i, seeds in enumerate(batches):
            if len(pending) == self.RING:
                drain_one()
            if state["overflow"]:
                while pending:
                    drain_one()
                torch.cuda.synchronize(self.dev)
                self.overflow_steps += 1
                self.check_overflow()          # worst-case capacities, buffers re-allocated, graph re-captured lazily
                state["overflow"] = False
            self.step_async(seeds)
            k = i % self.RING
            self.loss_ring[k].copy_(self.loss_dev, non_blocking=True)
            self.ovf_ring[k].copy_(self.overflow, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((k, ev))
        while pending:
            drain_one()
        if state["overflow"]:
            self.overflow_steps += 1
            self.check_overflow()
        return losses

    def total_sampled_edges(self):
        """Edges sampled since construction (device counter; syncs)."""
        return int(self.edges_total.sum().item())

    def sampled_edges_upper_bound(self):
        return sum(self.caps[l] * self.fanouts[l - 1] for l in range(1, self.L + 1))

    def last_sampled_edges(self):
        """Exact number of sampled edges of the last step (reads device counters; syncs)."""
        tot = 0
        for l in range(1, self.L + 1):
            n = int(self.n[l].item())
            tot += int(self.cnt[l][:n].sum().item())
        return tot

    @torch.no_grad()
    def predict(self, seeds, fanouts=None, batch_size=4096):
        """Sampled-neighbourhood inference for ``seeds`` (no dropout): returns fp32 logits [len(seeds), C].
        The trained weights are exported into a ``GraphSAGE`` module and evaluated on blocks drawn by the generic
        ``NeighborSampler`` (``fanouts`` defaults to the training fan-outs; pass ``[-1] * L`` for full neighbourhoods)."""
        from ..dataloading import NeighborSampler
        mod = GraphSAGE(self.F_orig, self.H, self.C, self.L, 0.0).to(self.dev).eval()
        self.export_to_module(mod)
        sampler = NeighborSampler(list(fanouts) if fanouts is not None else list(self.fanouts))
        g = self._graph_ref
        seeds = torch.as_tensor(seeds).to(self.dev).to(g.idtype)
        outs = []
        for chunk in seeds.split(int(batch_size)):
            inp, _, blocks = sampler.sample(g, chunk)
            x = self.feat[inp.long(), : self.F_orig].float()
            outs.append(mod(blocks, x).float())
        return torch.cat(outs) if outs else torch.empty((0, self.C), device=self.dev)
