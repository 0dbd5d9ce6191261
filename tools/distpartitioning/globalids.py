"""Contiguous ("shuffled") global id assignment (reference tools/distpartitioning/globalids.py): partition p numbers its
items [start_p, start_p + n_p) where start is the exclusive prefix sum of the per-partition counts."""
import torch

from .gloo_wrapper import allgather_ints


def assign_shuffle_global_ids(rank, world, count):
    counts = allgather_ints(count, world)
    starts = [0]
    for c in counts[:-1]:
        starts.append(starts[-1] + c)
    ids = torch.arange(starts[rank], starts[rank] + counts[rank], dtype=torch.int64)
    ranges = [[s, s + c] for s, c in zip(starts, counts)]
    return ids, ranges
