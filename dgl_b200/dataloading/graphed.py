"""Whole-step CUDA-graph execution of a user-defined ``nn.Module`` mini-batch pipeline.

``FusedSAGETrainer`` (models/graphsage.py) owns its kernels end to end; this is the generic counterpart for models written
against the public ``dgl.nn`` API: neighbour sampling, frontier compaction, the feature gather, the model's forward, the
loss, autograd's backward and the optimizer step are captured ONCE as a CUDA graph over capacity-shaped buffers and
replayed per mini-batch - the host only writes the seed ids into a pinned buffer. Nothing in the step reads a size back:

* sampling uses the one-launch-per-layer ELL sampler + cooperative compaction of ``sampling/fused_blocks.py`` on
  worst-case capacities with device-side counts, keyed by a device step counter so every replay draws fresh samples;
* the model sees ``CapacityBlock`` objects (duck-typed blocks: ELL neighbour table, per-row counts, device-side row
  count). ``SAGEConv`` consumes them through the tcgen05 fused layer (``nn/fused_sage.py``), which masks on the device
  count; row-wise layers (Linear, activations, dropout, norms over the feature dimension) are capacity-safe as they are;
* padded seed slots carry label ``ignore_index`` so the loss averages over the real seeds only;
* the optimizer must be capturable (``torch.optim.Adam(..., capturable=True)``, or any optimizer without host reads).

The reference has no counterpart: its loaders return variable-shaped blocks whose construction synchronises with the host
several times per layer (python/dgl/dataloading/neighbor_sampler.py:152-206, src/array/cuda/rowwise_sampling.cu).
"""
from __future__ import annotations

import torch

from ..base import DGLError
from ..sampling.fused_blocks import FusedBlockSampler, _csc_int32, _Workspace


class CapacityBlock:
    """A message-flow layer at fixed capacity: ``nbr[i, :cnt[i]]`` are the (local) source rows of destination ``i`` for
    ``i < n_dst_dev``; rows beyond that are padding. Quacks like a DGLBlock as far as the fused SAGE layer needs."""

    is_block = True

    def __init__(self, nbr, cnt, fanout, n_src_cap, n_dst_cap, n_dst_dev, global_input=None):
        self._ell_cache = (nbr, cnt, fanout)
        self._n_dst_dev = n_dst_dev
        if global_input is not None:        # (self ids, neighbour ids) as rows of the feature table handed to the model
            self._global_input = global_input
        self._caps = (int(n_src_cap), int(n_dst_cap))
        self.canonical_etypes = [("_N", "_E", "_N")]
        self.ntypes = self.srctypes = self.dsttypes = ["_N"]
        self.device = nbr.device

    def num_src_nodes(self, *_):
        return self._caps[0]

    def num_dst_nodes(self, *_):
        return self._caps[1]

    number_of_src_nodes, number_of_dst_nodes = num_src_nodes, num_dst_nodes

    def local_scope(self):
        import contextlib
        return contextlib.nullcontext()


class GraphedTrainer:
    """``trainer = GraphedTrainer(g, feat, labels, model, optimizer, fanouts, batch_size)``; ``trainer.step(seeds)``
    enqueues one training step and returns the (device) loss tensor of that step without synchronising.

    ``model(blocks, x)`` receives ``len(fanouts)`` CapacityBlocks (input layer first) and the gathered input rows
    ``[capacity_0, F]``; it must return ``[batch_size, C]`` logits for the seed slots. ``amp``: run the forward under
    ``torch.autocast(bf16)``.

    ``gather_inputs=False``: the model's first layer must be a ``SAGEConv`` (mean); it then receives the WHOLE feature
    table and a block addressed by global ids, and the fused layer reads the table in place - the input frontier is
    neither compacted nor gathered (it is the largest tensor of a step: ~0.7 ms of 2.3 ms at batch 1024, fan-out 10/10/10,
    128 features). Needs a feature width that is a multiple of 64 (bf16) / 32 (fp32)."""

    def __init__(self, g, feat, labels, model, optimizer, fanouts, batch_size, amp=False, loss_fn=None, seed=0,
                 ignore_index=-100, gather_inputs=True):
        if g.device.type != "cuda" or not feat.is_cuda:
            raise DGLError("GraphedTrainer needs the graph and the features on the GPU")
        if not all(isinstance(f, int) and 1 <= f <= 64 for f in fanouts):
            raise DGLError("fan-outs must be integers in [1, 64]")
        self.dev = g.device
        self.indptr, self.indices, _ = _csc_int32(g)
        self.feat, self.labels, self.model, self.opt = feat, labels.long(), model, optimizer
        self.gather_inputs = bool(gather_inputs)
        if not self.gather_inputs:
            if feat.shape[1] % 64:
                raise DGLError("gather_inputs=False needs a feature width that is a multiple of 64")
            if amp and feat.dtype != torch.bfloat16:
                self.feat = feat.to(torch.bfloat16)      # cast the table ONCE, not inside every step
        self.B, self.L, self.amp, self.ignore = int(batch_size), len(fanouts), bool(amp), int(ignore_index)
        self.loss_fn = loss_fn or (lambda logits, y: torch.nn.functional.cross_entropy(logits.float(), y, ignore_index=self.ignore))
        self.sampler = FusedBlockSampler(fanouts, seed)
        self.ws = _Workspace(self.B, list(fanouts), self.dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.RING = 4
        self._pinned = [torch.full((self.B,), -1, dtype=torch.int32).pin_memory() for _ in range(self.RING)]
        self._events = [None] * self.RING
        self._slot = 0
        self.loss = torch.zeros((), device=self.dev)
        self.graph = None
        self.edges = torch.zeros(1, dtype=torch.int64, device=self.dev)

    # ---- one step on the staged seeds (ws.seeds); everything below is capturable
    def _step_impl(self):
        ws, L = self.ws, self.L
        self.step_dev += 1
        ws.n[L].copy_((ws.seeds >= 0).sum().view(1))
        table_mode = not self.gather_inputs
        S, nbr_local = self.sampler.enqueue(ws, self.indptr, self.indices, step_dev=self.step_dev, compact_input=not table_mode)
        blocks = []
        for l in range(1, L + 1):
            if l == 1 and table_mode:
                blocks.append(CapacityBlock(ws.nbr_g[1], ws.cnt[1], self.sampler.fanouts[0], self.feat.shape[0], S[1].shape[0],
                                            ws.n[1], global_input=(S[1], ws.nbr_g[1])))
            else:
                blocks.append(CapacityBlock(nbr_local[l], ws.cnt[l], self.sampler.fanouts[l - 1], S[l - 1].shape[0],
                                            S[l].shape[0], ws.n[l]))
        x = self.feat if table_mode else self.feat[S[0].clamp(min=0).long()]
        seeds = ws.seeds.long()
        y = torch.where(seeds >= 0, self.labels[seeds.clamp(min=0)], torch.full_like(seeds, self.ignore))
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
            logits = self.model(blocks, x)
        # padded seed slots hold whatever the capacity rows contained: neutralise them before the loss (a NaN there would
        # survive `ignore_index` through the softmax backward)
        logits = torch.where((seeds >= 0).unsqueeze(1), logits[: self.B], torch.zeros((), dtype=logits.dtype, device=self.dev))
        loss = self.loss_fn(logits, y)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        self.loss.copy_(loss.detach())
        self.edges += sum(t.sum() for t in ws.totals[1:])

    def _capture(self):
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(3):                      # warm-up: lazy initialisations, optimizer state, autotuned workspaces
                self._step_impl()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_impl()
        torch.cuda.synchronize(self.dev)

    def load_seeds(self, seeds):
        n = len(seeds)
        if n > self.B:
            raise DGLError(f"batch of {n} seeds exceeds batch_size={self.B}")
        k = self._slot
        self._slot = (k + 1) % self.RING
        if self._events[k] is not None:
            self._events[k].synchronize()
        buf = self._pinned[k]
        buf[:n] = torch.as_tensor(seeds, dtype=torch.int32)
        buf[n:] = -1
        self.ws.seeds.copy_(buf, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._events[k] = ev

    def step(self, seeds):
        """Stage ``seeds`` (host tensor / list, at most ``batch_size``) and run one training step; returns the loss tensor
        on the device (valid once the stream reaches it)."""
        self.load_seeds(seeds)
        if self.graph is None:
            self._capture()
            self.load_seeds(seeds)                  # the warm-up consumed the staged batch
        self.graph.replay()
        return self.loss

    def sampled_edges(self):
        return int(self.edges.item())
