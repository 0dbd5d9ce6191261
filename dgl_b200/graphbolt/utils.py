"""GraphBolt tensor utilities (parity: /root/reference/python/dgl/graphbolt/internal/sample_utils.py:
unique_and_compact, unique_and_compact_csc_formats, compact_csc_format; external utils add_reverse_edges,
exclude_seed_edges)."""
from __future__ import annotations

from collections import defaultdict

import torch

from ..transforms.functional import unique_and_compact as _uc
from .base import CSCFormatBase, etype_str_to_tuple, isin


def unique_and_compact(nodes, rank=0, world_size=1):
    """Deduplicate node ids (per type); returns (unique, compacted_list)."""
    def one(lst):
        cat = torch.cat(lst)
        uniq, local = _uc(None, cat)
        out, off = [], 0
        for x in lst:
            out.append(local[off: off + x.shape[0]])
            off += x.shape[0]
        return uniq, out
    if isinstance(nodes, dict):
        u, c = {}, {}
        for k, v in nodes.items():
            lst = v if isinstance(v, (list, tuple)) else [v]
            u[k], c[k] = one(list(lst))
        return u, c
    lst = nodes if isinstance(nodes, (list, tuple)) else [nodes]
    return one(list(lst))


def unique_and_compact_csc_formats(csc_formats, unique_dst_nodes, rank=0, world_size=1):
    """Compact a sampled CSC whose ``indices`` hold global source ids: destination nodes keep their order and
    come first among the unique sources (the block invariant)."""
    if isinstance(csc_formats, dict):
        by_src = defaultdict(list)
        for et, csc in csc_formats.items():
            s, _, _ = etype_str_to_tuple(et)
            by_src[s].append((et, csc))
        uniq, out = {}, {}
        ntypes = set(unique_dst_nodes.keys()) | set(by_src.keys())
        for nt in ntypes:
            first = unique_dst_nodes.get(nt)
            lst = by_src.get(nt, [])
            ids = torch.cat([csc.indices for _, csc in lst]) if lst else None
            if ids is None:
                uniq[nt] = first
                continue
            if first is None:
                first = ids.new_empty(0)
            u, local = _uc(first, ids)
            uniq[nt] = u
            off = 0
            for et, csc in lst:
                n = csc.indices.shape[0]
                out[et] = CSCFormatBase(indptr=csc.indptr, indices=local[off: off + n])
                off += n
        return uniq, out
    u, local = _uc(unique_dst_nodes, csc_formats.indices)
    return u, CSCFormatBase(indptr=csc_formats.indptr, indices=local)


def compact_csc_format(csc_formats, dst_nodes, dst_timestamps=None):
    """Relabel WITHOUT deduplication: every sampled edge gets its own source slot after the dst nodes."""
    if isinstance(csc_formats, dict):
        # per source node type: [dst nodes of that type | one fresh slot per sampled edge, relation after relation]
        rows = {nt: [ids] for nt, ids in dst_nodes.items()}
        nxt = {nt: int(ids.shape[0]) for nt, ids in dst_nodes.items()}
        compacted = {}
        for et, csc in csc_formats.items():
            s, _, _ = etype_str_to_tuple(et)
            n = csc.indices.shape[0]
            start = nxt.get(s, 0)
            rows.setdefault(s, []).append(csc.indices)
            nxt[s] = start + n
            compacted[et] = CSCFormatBase(indptr=csc.indptr,
                                          indices=torch.arange(start, start + n, device=csc.indices.device,
                                                               dtype=csc.indices.dtype))
        return {nt: torch.cat(v) for nt, v in rows.items()}, compacted
    n_dst = dst_nodes.shape[0]
    n = csc_formats.indices.shape[0]
    original_row_ids = torch.cat([dst_nodes, csc_formats.indices])
    compacted = CSCFormatBase(indptr=csc_formats.indptr,
                              indices=torch.arange(n_dst, n_dst + n, device=csc_formats.indices.device,
                                                   dtype=csc_formats.indices.dtype))
    return original_row_ids, compacted


def add_reverse_edges(edges, reverse_etypes_mapping=None):
    """Append reversed seed edges (for exclude_seed_edges)."""
    if isinstance(edges, dict):
        out = {k: v for k, v in edges.items()}
        for et, e in edges.items():
            ret = reverse_etypes_mapping[et]
            rev = e.flip(dims=(1,))
            out[ret] = torch.cat([out[ret], rev]) if ret in out else rev
        return out
    return torch.cat([edges, edges.flip(dims=(1,))])


def exclude_edges_from_subgraph(sg, edges):
    from .minibatch import SampledSubgraphImpl

    def one(csc, row_ids, col_ids, eids, ex):
        deg = (csc.indptr[1:] - csc.indptr[:-1]).long()
        dst_local = torch.repeat_interleave(torch.arange(deg.shape[0], device=deg.device), deg)
        src_g = row_ids[csc.indices.long()]
        dst_g = col_ids[dst_local]
        big = int(max(int(row_ids.max()) if row_ids.numel() else 0, int(col_ids.max()) if col_ids.numel() else 0, int(ex.max()) if ex.numel() else 0)) + 1
        key = src_g.long() * big + dst_g.long()
        exk = ex[:, 0].long() * big + ex[:, 1].long()
        keep = ~isin(key, exk)
        new_deg = torch.zeros_like(deg).index_add_(0, dst_local[keep], torch.ones_like(dst_local[keep]))
        indptr = torch.zeros_like(csc.indptr)
        indptr[1:] = torch.cumsum(new_deg, 0)
        return CSCFormatBase(indptr=indptr, indices=csc.indices[keep]), (eids[keep] if eids is not None else None)

    if isinstance(sg.sampled_csc, dict):
        new_csc, new_e = {}, {}
        for et, csc in sg.sampled_csc.items():
            s, _, d = etype_str_to_tuple(et)
            if et in edges:
                c, e = one(csc, sg.original_row_node_ids[s], sg.original_column_node_ids[d],
                           None if sg.original_edge_ids is None else sg.original_edge_ids.get(et), edges[et])
            else:
                c, e = csc, None if sg.original_edge_ids is None else sg.original_edge_ids.get(et)
            new_csc[et] = c
            if e is not None:
                new_e[et] = e
        return SampledSubgraphImpl(new_csc, sg.original_column_node_ids, sg.original_row_node_ids, new_e or None)
    c, e = one(sg.sampled_csc, sg.original_row_node_ids, sg.original_column_node_ids, sg.original_edge_ids, edges)
    return SampledSubgraphImpl(c, sg.original_column_node_ids, sg.original_row_node_ids, e)


def exclude_seed_edges(minibatch, include_reverse_edges=False, reverse_etypes_mapping=None, asynchronous=False):
    """Remove the seed edges (and optionally their reverses) from every sampled layer (link prediction)."""
    edges = minibatch.seeds
    if include_reverse_edges:
        edges = add_reverse_edges(edges, reverse_etypes_mapping)
    minibatch.sampled_subgraphs = [sg.exclude_edges(edges) for sg in minibatch.sampled_subgraphs]
    return minibatch
