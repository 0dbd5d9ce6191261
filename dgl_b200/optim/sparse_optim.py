"""SparseAdagrad / SparseAdam: row-wise updates of only the embedding rows touched in the step.

Parity: /root/reference/python/dgl/optim/pytorch/sparse_optim.py (SparseGradOptimizer :17, multi-GPU exchange
:184-433, state_dict :455-500). For sharded tables gradients are pushed to the owning rank with a peer-memory
scatter-add kernel and each rank updates its own shard."""
from __future__ import annotations

import abc

import torch
import torch.distributed as dist

from ..base import DGLError
from ..nn.sparse_emb import NodeEmbedding


class SparseGradOptimizer(abc.ABC):
    def __init__(self, params, lr):
        self._params = params
        self._lr = lr
        self._rank = None
        self._world_size = None
        self._clean_grad = False
        self._opt_meta = {}
        for emb in params:
            assert isinstance(emb, NodeEmbedding), "DGL SparseOptimizer only supports dgl.nn.NodeEmbedding"

    def zero_grad(self):
        self._clean_grad = True
        for emb in self._params:
            emb.reset_trace()

    def _collect(self, emb):
        idx, grads = [], []
        for i, data in emb._trace:
            if data.grad is not None:
                idx.append(i.reshape(-1))
                grads.append(data.grad.reshape(-1, emb.embedding_dim))
        if not idx:
            return None, None
        return torch.cat(idx), torch.cat(grads)

    def step(self):
        with torch.no_grad():
            for emb in self._params:
                idx, grad = self._collect(emb)
                if emb._shards is not None:
                    comm = emb._comm
                    # push: scatter-add every (idx, grad) row into the owner's gradient accumulator over NVLink
                    state = self._state(emb)
                    gacc, cnt = state["gacc"], state["cnt"]
                    comm.barrier(3)
                    if idx is not None:
                        g = torch.zeros((idx.shape[0], emb._dim_pad), dtype=torch.float32, device=comm.device)
                        g[:, : emb.embedding_dim] = grad.to(comm.device)
                        comm.scatter_add(gacc, emb._shard_rows, idx.to(comm.device).contiguous(), g)
                        ones = torch.ones((idx.shape[0], 4), dtype=torch.float32, device=comm.device)
                        comm.scatter_add(cnt, emb._shard_rows, idx.to(comm.device).contiguous(), ones)
                    comm.barrier(3)
                    gl = gacc.view(torch.float32)[: emb._shard_rows * emb._dim_pad].view(emb._shard_rows, emb._dim_pad)
                    cl = cnt.view(torch.float32)[: emb._shard_rows * 4].view(emb._shard_rows, 4)[:, 0]
                    rows = torch.nonzero(cl > 0).squeeze(1)
                    if rows.numel():
                        gm = gl[rows, : emb.embedding_dim] / cl[rows].unsqueeze(1)
                        self.update_rows(emb, rows, gm, emb._tensor[:, : emb.embedding_dim])
                        gl[rows] = 0
                        cl_full = cnt.view(torch.float32)[: emb._shard_rows * 4].view(emb._shard_rows, 4)
                        cl_full[rows] = 0
                    comm.barrier(3)
                elif idx is not None:
                    dev = emb.weight.device
                    idx, grad = idx.to(dev), grad.to(dev)
                    uniq, inv, cnt = torch.unique(idx, return_inverse=True, return_counts=True)
                    gsum = torch.zeros((uniq.shape[0], grad.shape[1]), device=dev, dtype=grad.dtype)
                    gsum.index_add_(0, inv, grad)
                    gsum = gsum / cnt.unsqueeze(1)
                    self.update_rows(emb, uniq, gsum, emb.weight)
                emb.reset_trace()
        self._clean_grad = False

    def _state(self, emb):
        raise NotImplementedError

    def setup(self, params):
        """Allocate the optimizer state of every embedding (idempotent; the constructors call the same path)."""
        for emb in params:
            self._state(emb)

    def update(self, idx, grad, emb):
        """Apply one update to rows ``idx`` of ``emb`` with gradient ``grad`` (reference sparse_optim.py:update):
        duplicates in ``idx`` are averaged first, exactly as :meth:`step` does."""
        with torch.no_grad():
            table = emb.weight if emb._shards is None else emb._tensor[:, : emb.embedding_dim]
            idx, grad = idx.to(table.device), grad.to(table.device)
            uniq, inv, cnt = torch.unique(idx, return_inverse=True, return_counts=True)
            gsum = torch.zeros((uniq.shape[0], grad.shape[1]), device=table.device, dtype=grad.dtype).index_add_(0, inv, grad)
            self.update_rows(emb, uniq, gsum / cnt.unsqueeze(1), table)

    @abc.abstractmethod
    def update_rows(self, emb, rows, grad, table):
        """apply the optimizer update to table[rows] with averaged gradient ``grad``"""

    def state_dict(self, **kwargs):
        return {"state": {emb.name: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._opt_meta.get(emb.name, {}).items()
                                     if k not in ("gacc", "cnt")} for emb in self._params},
                "param_groups": [{"lr": self._lr}]}

    def load_state_dict(self, state_dict, **kwargs):
        for emb in self._params:
            st = state_dict["state"].get(emb.name, {})
            cur = self._state(emb)
            for k, v in st.items():
                if torch.is_tensor(v) and k in cur:
                    cur[k].copy_(v)
                else:
                    cur[k] = v
        self._lr = state_dict["param_groups"][0]["lr"]

    @property
    def param_groups(self):
        return [{"lr": self._lr}]


def _alloc_like(emb, ncols=None):
    if emb._shards is not None:
        return torch.zeros((emb._shard_rows, ncols or emb.embedding_dim), dtype=torch.float32, device=emb._comm.device)
    return torch.zeros((emb.num_embeddings, ncols or emb.embedding_dim), dtype=torch.float32, device=emb.weight.device)


class SparseAdagrad(SparseGradOptimizer):
    def __init__(self, params, lr, eps=1e-10):
        super().__init__(params, lr)
        self._eps = eps
        for emb in params:
            self._state(emb)

    def _state(self, emb):
        st = self._opt_meta.get(emb.name)
        if st is None:
            st = {"sum": _alloc_like(emb)}
            if emb._shards is not None:
                st["gacc"] = emb._comm.alloc(emb._shard_rows * emb._dim_pad * 4)
                st["cnt"] = emb._comm.alloc(emb._shard_rows * 4 * 4)
            self._opt_meta[emb.name] = st
            emb.set_optm_state((st["sum"],))
        return st

    def update_rows(self, emb, rows, grad, table):
        st = self._state(emb)
        gsq = grad * grad
        st["sum"][rows] += gsq
        std = st["sum"][rows].sqrt_().add_(self._eps)
        table[rows] -= self._lr * grad / std


class SparseAdam(SparseGradOptimizer):
    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-08, use_uva=None, dtype=torch.float32):
        super().__init__(params, lr)
        self._lr, self._beta1, self._beta2, self._eps = lr, betas[0], betas[1], eps
        for emb in params:
            self._state(emb)

    def _state(self, emb):
        st = self._opt_meta.get(emb.name)
        if st is None:
            st = {"step": _alloc_like(emb, 1), "mem": _alloc_like(emb), "power": _alloc_like(emb)}
            if emb._shards is not None:
                st["gacc"] = emb._comm.alloc(emb._shard_rows * emb._dim_pad * 4)
                st["cnt"] = emb._comm.alloc(emb._shard_rows * 4 * 4)
            self._opt_meta[emb.name] = st
            emb.set_optm_state((st["step"], st["mem"], st["power"]))
        return st

    def update_rows(self, emb, rows, grad, table):
        st = self._state(emb)
        st["step"][rows] += 1
        t = st["step"][rows]
        m = st["mem"][rows] * self._beta1 + (1 - self._beta1) * grad
        v = st["power"][rows] * self._beta2 + (1 - self._beta2) * grad * grad
        st["mem"][rows] = m
        st["power"][rows] = v
        mh = m / (1 - torch.pow(torch.tensor(self._beta1, device=t.device), t))
        vh = v / (1 - torch.pow(torch.tensor(self._beta2, device=t.device), t))
        table[rows] -= self._lr * mh / (vh.sqrt() + self._eps)
