"""Work splitting for (Distributed)ItemSampler (parity: graphbolt/internal/item_sampler_utils.py).

Items are dealt to replicas (and to the data-loading workers inside a replica) in whole batches so that every batch
is produced by exactly one worker and replicas differ by at most one batch."""
from __future__ import annotations

__all__ = ["count_split", "calculate_range"]


def count_split(total, num_workers, worker_id, batch_size=1):
    """(items assigned to ``worker_id``, items assigned to all earlier workers) when ``total`` items are handed out
    batch by batch, round-robin over whole rounds and then left to right for the remainder."""
    per_round = num_workers * batch_size
    rounds, rest = divmod(total, per_round)
    full, tail = divmod(rest, batch_size)          # the remainder fills `full` workers and gives `tail` to the next
    if worker_id < full:
        extra = batch_size
    elif worker_id == full:
        extra = tail
    else:
        extra = 0
    before = rounds * batch_size * worker_id + min(worker_id * batch_size, rest)
    return rounds * batch_size + extra, before


def calculate_range(distributed, total, num_replicas, rank, num_workers, worker_id, batch_size, drop_last,
                    drop_uneven_inputs):
    """(start offset, number of items assigned, number of items actually emitted) for one worker of one replica.

    ``drop_last`` removes a trailing short batch; ``drop_uneven_inputs`` makes every replica emit the same number of
    batches (the count of the least-loaded replica) so collective calls inside the training loop stay aligned."""
    if not distributed:
        return 0, total, (total // batch_size * batch_size if drop_last else total)
    assigned, start = count_split(total, num_replicas, rank, batch_size)
    if drop_uneven_inputs:
        if drop_last:
            emitted = total // (batch_size * num_replicas) * batch_size
        else:
            least, _ = count_split(total, num_replicas, num_replicas - 1, batch_size)
            emitted = min(-(-least // batch_size) * batch_size, assigned)
    else:
        emitted = assigned // batch_size * batch_size if drop_last else assigned
    if num_workers > 1:
        dropped, dropped_before = count_split(assigned - emitted, num_workers, worker_id)
        emitted, emitted_before = count_split(emitted, num_workers, worker_id, batch_size)
        assigned = emitted + dropped
        start += emitted_before + dropped_before
    return start, assigned, emitted
