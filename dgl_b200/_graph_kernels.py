"""Python bindings of csrc/cuda/graph_kernels.cu (random walks, node2vec, LABOR-0, negative sampling, FPS, k-NN,
neighbour matching, partition ops, CSR<->COO helpers). Every function takes CUDA tensors and enqueues on the
current stream; callers (sampling/, geometry.py, transforms/, partition.py, graph_index.py) pick these on CUDA inputs
and keep their torch formulations for CPU tensors.
"""
from __future__ import annotations

import ctypes

import torch

from ._native import loader
from .base import dtype_code, id_code


def _lib():
    return loader.cuda_lib()


def _p(t):
    return 0 if t is None else t.data_ptr()


def random_walk(rels, step_rel, restart, seeds, length, seed, return_eids=True):
    """rels: list of (indptr, indices, eids|None, prob|None) CSR relations (<= 16); step_rel: [L] int32 device tensor."""
    idt = seeds.dtype
    n = seeds.shape[0]
    traces = torch.empty((n, length + 1), dtype=idt, device=seeds.device)
    eids = torch.empty((n, length), dtype=idt, device=seeds.device) if return_eids else None
    keep = []
    ptrs = (ctypes.c_void_p * (4 * len(rels)))()
    for r, (ip, ix, ei, pr) in enumerate(rels):
        pr = None if pr is None else pr.to(device=seeds.device, dtype=torch.float32).contiguous()
        keep.append(pr)
        ptrs[4 * r], ptrs[4 * r + 1], ptrs[4 * r + 2], ptrs[4 * r + 3] = ip.data_ptr(), ix.data_ptr(), _p(ei) or None, _p(pr) or None
    _lib().fn("dglb200_cuda_random_walk", "ipippplilppp")(
        id_code(idt), ctypes.cast(ptrs, ctypes.c_void_p), len(rels), step_rel.data_ptr(), _p(restart), seeds.data_ptr(), n,
        int(length), int(seed) & 0x7FFFFFFFFFFFFFFF, traces.data_ptr(), _p(eids), loader.stream_ptr())
    return traces, eids


def node2vec_walk(indptr, indices, eids, seeds, length, p, q, seed, return_eids=True):
    idt = seeds.dtype
    n = seeds.shape[0]
    traces = torch.empty((n, length + 1), dtype=idt, device=seeds.device)
    te = torch.empty((n, length), dtype=idt, device=seeds.device) if return_eids else None
    _lib().fn("dglb200_cuda_node2vec_walk", "ippppliddlppp")(
        id_code(idt), indptr.data_ptr(), indices.data_ptr(), _p(eids), seeds.data_ptr(), n, int(length), float(p), float(q),
        int(seed) & 0x7FFFFFFFFFFFFFFF, traces.data_ptr(), _p(te), loader.stream_ptr())
    return traces, te


def labor0(indptr, indices, eids, seeds, fanout, seed):
    """LABOR-0 on a CSC: returns (src, seed position, eid, counts per seed)."""
    idt = seeds.dtype
    n = seeds.shape[0]
    dev = seeds.device
    counts = torch.empty(n, dtype=torch.int64, device=dev)
    fn = _lib().fn("dglb200_cuda_labor0", "iipppplilpppppp")
    sd = int(seed) & 0x7FFFFFFFFFFFFFFF
    fn(id_code(idt), 0, indptr.data_ptr(), indices.data_ptr(), _p(eids), seeds.data_ptr(), n, int(fanout), sd,
       counts.data_ptr(), 0, 0, 0, 0, loader.stream_ptr())
    offs = torch.cumsum(counts, 0) - counts
    total = int(counts.sum())
    src = torch.empty(total, dtype=idt, device=dev)
    pos = torch.empty(total, dtype=idt, device=dev)
    eid = torch.empty(total, dtype=idt, device=dev)
    if total:
        fn(id_code(idt), 1, indptr.data_ptr(), indices.data_ptr(), _p(eids), seeds.data_ptr(), n, int(fanout), sd,
           0, offs.data_ptr(), src.data_ptr(), pos.data_ptr(), eid.data_ptr(), loader.stream_ptr())
    return src, pos, eid, counts


def labor_sample(indptr, indices, eids, seeds, fanout, prob, iters, seed1, seed2=0, mix=0.0, num_src=None, cache=None,
                 return_state=False):
    """LABOR-i on a CSC, entirely on the device (csrc/cuda/labor.cu). ``iters`` importance iterations (< 0: until the
    importances stop moving, at most 10), ``prob`` optional per-edge weights (edge-id indexed), ``(seed1, seed2, mix)``
    the continuous-seed variate. Returns (src, seed position, eid, importance weight[, (c, pi)]). One host read (the number
    of sampled edges) per call; the importance buffers are indexed by vertex id and live in ``cache`` between calls."""
    lib = _lib()
    idt, dev, n = seeds.dtype, seeds.device, seeds.shape[0]
    num_src = int(num_src if num_src is not None else (int(indices.max()) + 1 if indices.numel() else 1))
    cache = {} if cache is None else cache
    key = ("labor_pi", dev.index, num_src)
    bufs = cache.get(key)
    if bufs is None:
        bufs = cache[key] = (torch.ones(num_src, dtype=torch.float32, device=dev), torch.zeros(num_src, dtype=torch.float32, device=dev))
    pi, pi_next = bufs
    c = torch.empty(n, dtype=torch.float32, device=dev)
    pcode = -1 if prob is None else dtype_code(prob.dtype)
    common = (id_code(idt),)
    st = loader.stream_ptr
    solve = lib.fn("dglb200_cuda_labor_solve_c", "ipppplipippp")
    ppass = lib.fn("dglb200_cuda_labor_pi_pass", "iipppplpppipp")
    pick = lib.fn("dglb200_cuda_labor_pick", "iipppplppipllfpppppppp")
    g = (indptr.data_ptr(), indices.data_ptr(), _p(eids), seeds.data_ptr(), n)

    def run_solve():
        solve(id_code(idt), *g, int(fanout), pi.data_ptr(), pcode, _p(prob), c.data_ptr(), st())

    def run_pass(mode):
        ppass(id_code(idt), mode, *g, c.data_ptr(), pi.data_ptr(), pi_next.data_ptr(), pcode, _p(prob), st())

    run_solve()
    rounds = 10 if iters < 0 else int(iters)
    for _ in range(rounds):
        run_pass(0)          # pi_next[t] = max_s c_s pi_t a_ts
        run_pass(1)          # pi[t] <- pi_next[t]
        run_pass(2)          # pi_next[t] <- 0 (ready for the next round / call)
        run_solve()
    counts = torch.empty(n, dtype=torch.int64, device=dev)
    wsum = torch.empty(n, dtype=torch.float32, device=dev)
    s1, s2 = int(seed1) & 0x7FFFFFFFFFFFFFFF, int(seed2) & 0x7FFFFFFFFFFFFFFF
    pick(id_code(idt), 0, *g, c.data_ptr(), pi.data_ptr(), pcode, _p(prob), s1, s2, float(mix), counts.data_ptr(),
         wsum.data_ptr(), 0, 0, 0, 0, 0, st())
    offs = torch.cumsum(counts, 0) - counts
    total = int(counts.sum())
    src = torch.empty(total, dtype=idt, device=dev)
    pos = torch.empty(total, dtype=idt, device=dev)
    eid = torch.empty(total, dtype=idt, device=dev)
    w = torch.empty(total, dtype=torch.float32, device=dev)
    if total:
        pick(id_code(idt), 1, *g, c.data_ptr(), pi.data_ptr(), pcode, _p(prob), s1, s2, float(mix), 0, wsum.data_ptr(),
             offs.data_ptr(), src.data_ptr(), pos.data_ptr(), eid.data_ptr(), w.data_ptr(), st())
    state = None
    if return_state:
        touched = torch.unique(indices[_expand_rows(indptr, seeds)].long()) if n else torch.empty(0, dtype=torch.long, device=dev)
        state = (c.clone(), touched, pi[touched].clone())
    if rounds:
        run_pass(3)          # pi[t] <- 1 for the touched vertices: the cached buffer is neutral again
    return (src, pos, eid, w, state) if return_state else (src, pos, eid, w)


def _expand_rows(indptr, rows):
    s = indptr[rows.long()].long()
    deg = indptr[rows.long() + 1].long() - s
    total = int(deg.sum())
    owner = torch.repeat_interleave(torch.arange(rows.shape[0], device=rows.device), deg, output_size=total)
    first = torch.cumsum(deg, 0) - deg
    return s[owner] + (torch.arange(total, device=rows.device) - first[owner])


def neg_sample(indptr, indices, sorted_rows, n_src, n_dst, exclude_self, trials, seed):
    idt = indptr.dtype
    dev = indptr.device
    u = torch.empty(trials, dtype=idt, device=dev)
    v = torch.empty(trials, dtype=idt, device=dev)
    ok = torch.empty(trials, dtype=torch.uint8, device=dev)
    _lib().fn("dglb200_cuda_neg_sample", "ippillillpppp")(
        id_code(idt), indptr.data_ptr(), indices.data_ptr(), 1 if sorted_rows else 0, int(n_src), int(n_dst),
        1 if exclude_self else 0, int(trials), int(seed) & 0x7FFFFFFFFFFFFFFF, u.data_ptr(), v.data_ptr(), ok.data_ptr(),
        loader.stream_ptr())
    return u, v, ok.bool()


def fps(pos, npoints, start):
    """pos: [B, N, C] CUDA float tensor, start: [B] int64 -> [B, npoints] int64."""
    B, N, C = pos.shape
    pos = pos.contiguous()
    dist = torch.empty((B, N), dtype=torch.float32, device=pos.device)
    out = torch.empty((B, npoints), dtype=torch.int64, device=pos.device)
    _lib().fn("dglb200_cuda_fps", "ipllii" "pppp")(
        dtype_code(pos.dtype), pos.data_ptr(), B, N, C, int(npoints), start.data_ptr(), dist.data_ptr(), out.data_ptr(),
        loader.stream_ptr())
    return out


def knn_brute(data, data_off, query, query_off, k, return_dist=False):
    """k nearest data points (global ids) of every query inside its segment. *_off: [n_seg+1] int64 device tensors."""
    C = data.shape[1]
    nq = query.shape[0]
    data, query = data.contiguous(), query.contiguous()
    out = torch.empty((nq, k), dtype=torch.int64, device=data.device)
    dist = torch.empty((nq, k), dtype=torch.float32, device=data.device) if return_dist else None
    n_seg = data_off.shape[0] - 1
    max_q = int((query_off[1:] - query_off[:-1]).max()) if n_seg else 0
    _lib().fn("dglb200_cuda_knn_brute", "ippppiliippp")(
        dtype_code(data.dtype), data.data_ptr(), data_off.data_ptr(), query.data_ptr(), query_off.data_ptr(), n_seg, max_q,
        C, int(k), out.data_ptr(), _p(dist), loader.stream_ptr())
    return (out, dist) if return_dist else out


def neighbor_matching(indptr, indices, eids, weights, n, seed, rounds=12):
    idt = indptr.dtype
    match = torch.full((n,), -1, dtype=idt, device=indptr.device)
    prop = torch.empty(n, dtype=idt, device=indptr.device)
    w = None if weights is None else weights.to(torch.float32).contiguous()
    _lib().fn("dglb200_cuda_neighbor_matching", "ippppllippp")(
        id_code(idt), indptr.data_ptr(), indices.data_ptr(), _p(eids), _p(w), int(n), int(seed) & 0x7FFFFFFFFFFFFFFF,
        int(rounds), match.data_ptr(), prop.data_ptr(), loader.stream_ptr())
    return match


def partition_owner(idx, num_parts, ranges=None, want_counts=True, want_local=False):
    """owner (int32) of every index under a remainder (ranges=None) or range partition; optional counts / local ids."""
    n = idx.shape[0]
    dev = idx.device
    owner = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.zeros(num_parts, dtype=torch.int64, device=dev) if want_counts else None
    local = torch.empty_like(idx) if want_local else None
    rg = None if ranges is None else ranges.to(device=dev, dtype=torch.int64).contiguous()
    _lib().fn("dglb200_cuda_partition_owner", "iplippppp")(
        id_code(idx.dtype), idx.data_ptr(), n, int(num_parts), _p(rg), owner.data_ptr(), _p(counts), _p(local),
        loader.stream_ptr())
    return owner, counts, local


def partition_perm(owner, num_parts):
    """Stable permutation grouping positions by owner (radix sort on ceil(log2 P) key bits)."""
    n = owner.shape[0]
    dev = owner.device
    lib = _lib()
    f = lib.raw("dglb200_cuda_partition_perm_temp_bytes", ctypes.c_int64, [ctypes.c_int64])
    tb = int(f(n))
    temp = torch.empty(max(tb, 1), dtype=torch.uint8, device=dev)
    iota = torch.empty(n, dtype=torch.int64, device=dev)
    keys = torch.empty(n, dtype=torch.int32, device=dev)
    perm = torch.empty(n, dtype=torch.int64, device=dev)
    lib.fn("dglb200_cuda_partition_perm", "plippppl" "p")(
        owner.data_ptr(), n, int(num_parts), iota.data_ptr(), keys.data_ptr(), perm.data_ptr(), temp.data_ptr(), tb,
        loader.stream_ptr())
    return perm


def expand_indptr(indptr, nnz=None, node_ids=None):
    """CSR -> COO row ids (optionally mapped through node_ids)."""
    n_rows = indptr.shape[0] - 1
    if nnz is None:
        nnz = int(indptr[-1])
    rows = torch.empty(nnz, dtype=indptr.dtype, device=indptr.device)
    _lib().fn("dglb200_cuda_expand_indptr", "ipllppp")(
        id_code(indptr.dtype), indptr.data_ptr(), n_rows, int(nnz), _p(node_ids), rows.data_ptr(), loader.stream_ptr())
    return rows


def row_counts_to_indptr(row, num_rows):
    """indptr of a COO row array (rows need not be sorted): atomic histogram + scan."""
    indptr = torch.zeros(num_rows + 1, dtype=row.dtype, device=row.device)
    _lib().fn("dglb200_cuda_row_hist", "iplpp")(id_code(row.dtype), row.data_ptr(), row.shape[0],
                                                indptr.data_ptr() + row.element_size(), loader.stream_ptr())
    return torch.cumsum(indptr, 0, dtype=row.dtype)


def keyed_sum(keys, vals, num_keys):
    """out[k] = sum of the rows of ``vals`` ([n, C] fp32) whose key is k; few keys, shared-memory privatised."""
    vals = vals.contiguous().float()
    out = torch.zeros((num_keys, vals.shape[1]), dtype=torch.float32, device=vals.device)
    _lib().fn("dglb200_cuda_keyed_sum", "iplpiipp")(id_code(keys.dtype), keys.data_ptr(), keys.shape[0], vals.data_ptr(),
                                                   vals.shape[1], int(num_keys), out.data_ptr(), loader.stream_ptr())
    return out


class _KeyedGather(torch.autograd.Function):
    """table[keys] whose backward is a keyed sum (no serialised global atomics on a handful of table rows)."""

    @staticmethod
    def forward(ctx, table, keys):
        ctx.save_for_backward(keys)
        ctx.shape = table.shape
        return table.index_select(0, keys.long())

    @staticmethod
    def backward(ctx, grad):
        (keys,) = ctx.saved_tensors
        R = ctx.shape[0]
        g2 = grad.reshape(grad.shape[0], -1)
        if grad.is_cuda and R * g2.shape[1] <= 8192 and keys.dtype in (torch.int32, torch.int64):
            return keyed_sum(keys.contiguous(), g2, R).to(grad.dtype).view(ctx.shape), None
        return torch.zeros(ctx.shape, dtype=grad.dtype, device=grad.device).index_add_(0, keys.long(), grad), None


def keyed_gather(table, keys):
    return _KeyedGather.apply(table, keys)
