"""Distributed id lookup service (reference tools/distpartitioning/dist_lookup.py).

The global node-id space [0, N) is split into ``world`` contiguous ranges; the owner of a range holds, for its ids, (a)
the partition id read from the assignment files and (b) - once the partitions have numbered their nodes - the shuffled
global id. ``get_partition_ids`` / ``get_shuffle_nids`` are COLLECTIVE calls (every rank must call them, possibly with
an empty query): ids are bucketed by range owner, exchanged with all-to-all-v, answered locally, and sent back."""
import torch

from .gloo_wrapper import alltoallv_cpu


class DistLookupService:
    def __init__(self, rank, world, num_nodes, partition_ids_of_my_range):
        self.rank, self.world, self.N = rank, world, int(num_nodes)
        self.bounds = torch.tensor([self.N * r // world for r in range(world + 1)], dtype=torch.int64)
        self.lo, self.hi = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.part_ids = torch.as_tensor(partition_ids_of_my_range, dtype=torch.int64)
        assert self.part_ids.shape[0] == self.hi - self.lo
        self.shuffle_ids = torch.full((self.hi - self.lo,), -1, dtype=torch.int64)

    def owner_of(self, gids):
        return torch.searchsorted(self.bounds[1:].contiguous(), gids, right=True)

    def _query(self, gids, table):
        gids = torch.as_tensor(gids, dtype=torch.int64)
        owner = self.owner_of(gids)
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=self.world)
        send = list(torch.split(gids[order], counts.tolist()))
        asked = alltoallv_cpu(self.rank, self.world, send)
        answers = [table[q - self.lo] for q in asked]
        back = alltoallv_cpu(self.rank, self.world, answers)
        out = torch.empty_like(gids)
        out[order] = torch.cat(back) if back else gids.new_empty(0)
        return out

    def get_partition_ids(self, gids):
        return self._query(gids, self.part_ids)

    def set_shuffle_nids(self, orig_gids, new_gids):
        """COLLECTIVE: every partition announces (original id -> shuffled id) of the nodes it owns to the range owners."""
        orig_gids = torch.as_tensor(orig_gids, dtype=torch.int64)
        owner = self.owner_of(orig_gids)
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=self.world).tolist()
        pairs = torch.stack([orig_gids[order], torch.as_tensor(new_gids, dtype=torch.int64)[order]], 1)
        for got in alltoallv_cpu(self.rank, self.world, list(torch.split(pairs, counts))):
            if got.shape[0]:
                self.shuffle_ids[got[:, 0] - self.lo] = got[:, 1]

    def get_shuffle_nids(self, gids):
        return self._query(gids, self.shuffle_ids)
