"""Chunk readers (reference tools/distpartitioning/dataset_utils.py): every rank reads ONLY the chunk files / row ranges
it is responsible for."""
import os

import numpy as np
import torch

from . import constants as C
from .array_readwriter import get_array_parser


def type_offsets(meta):
    """Homogeneous id space: start of every node / edge type."""
    nn_ = [int(sum(s)) for s in meta[C.STR_NUM_NODES_PER_CHUNK]]
    ne_ = [int(sum(s)) for s in meta[C.STR_NUM_EDGES_PER_CHUNK]]
    noff = np.concatenate([[0], np.cumsum(nn_)]).astype(np.int64)
    eoff = np.concatenate([[0], np.cumsum(ne_)]).astype(np.int64)
    return noff, eoff


def read_rows(in_dir, spec, chunk_sizes, lo, hi):
    """Rows [lo, hi) of a chunked array described by ``spec`` ({'format': {...}, 'data': [files]})."""
    parser = get_array_parser(spec[C.STR_FORMAT][C.STR_NAME])
    offs = np.concatenate([[0], np.cumsum(chunk_sizes)])
    parts = []
    for i, f in enumerate(spec[C.STR_DATA]):
        a, b = max(lo, int(offs[i])), min(hi, int(offs[i + 1]))
        if a < b:
            arr = parser.read(os.path.join(in_dir, f))
            parts.append(np.asarray(arr[a - int(offs[i]): b - int(offs[i])]))
    if not parts:
        return None
    return torch.from_numpy(np.ascontiguousarray(np.concatenate(parts)))


def read_my_edges(in_dir, meta, rank, world):
    """Edges of the chunks assigned to this rank (chunk c of every edge type -> rank c % world) in GLOBAL ids:
    returns int64 [m, 4] = (global src, global dst, edge type id, original global edge id) and the per-etype row ranges
    read (for the edge features)."""
    noff, eoff = type_offsets(meta)
    ntid = {t: i for i, t in enumerate(meta[C.STR_NODE_TYPE])}
    rows, ranges = [], []
    for ei, et in enumerate(meta[C.STR_EDGE_TYPE]):
        s, _, d = et.split(":")
        spec = meta[C.STR_EDGES][et]
        parser = get_array_parser(spec[C.STR_FORMAT][C.STR_NAME])
        sizes = meta[C.STR_NUM_EDGES_PER_CHUNK][ei]
        offs = np.concatenate([[0], np.cumsum(sizes)])
        for c, f in enumerate(spec[C.STR_DATA]):
            if c % world != rank or sizes[c] == 0:
                continue
            e = torch.from_numpy(np.ascontiguousarray(np.asarray(parser.read(os.path.join(in_dir, f))))).long()
            m = e.shape[0]
            rows.append(torch.stack([e[:, 0] + int(noff[ntid[s]]), e[:, 1] + int(noff[ntid[d]]),
                                     torch.full((m,), ei, dtype=torch.int64),
                                     torch.arange(m, dtype=torch.int64) + int(eoff[ei]) + int(offs[c])], 1))
            ranges.append((et, int(offs[c]), int(offs[c + 1])))
    if not rows:
        return torch.empty((0, 4), dtype=torch.int64), ranges
    return torch.cat(rows), ranges
