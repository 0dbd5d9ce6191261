"""Stage timings of the host feature path (cache lookup, io_uring read, combine, insert) at different thread budgets.
Host-only (no GPU needed); used to size the I/O pool and the OpenMP teams on many-core boxes."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from dgl_b200._native import loader
from dgl_b200.graphbolt.impl.features import CPUCachedFeature, DiskBasedFeature


def main():
    n, dim = 2_000_000, 128
    path = "/tmp/dglb200_profile_feat.npy"
    if not os.path.exists(path):
        t = np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=(n, dim))
        t[:] = 1.0
        t.flush()
        del t
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "torch threads", torch.get_num_threads())
    for threads in (None, 32, 16, 8):
        if threads is not None:
            torch.set_num_threads(threads)
            loader.cpu_lib().fn("dglb200_cpu_set_num_threads", "i")(threads)
        disk = DiskBasedFeature(path)
        feat = CPUCachedFeature(disk, 256 << 20)
        g = torch.Generator().manual_seed(0)

        def T(f, *a):
            t0 = time.perf_counter()
            r = f(*a)
            return r, (time.perf_counter() - t0) * 1e3

        acc = [0.0] * 5
        for it in range(6):
            ids = torch.randint(0, n, (110_000,), generator=g).unique()
            (vals, hit, miss, mk), t1 = T(feat._feature.query, ids)
            m, t2 = T(disk.read, mk)
            out = torch.empty((ids.shape[0], dim))
            _, t3 = T(lambda: out.__setitem__(hit, vals))
            _, t4 = T(lambda: out.__setitem__(miss, m))
            _, t5 = T(feat._feature.replace, mk, m)
            if it >= 2:
                for i, t in enumerate((t1, t2, t3, t4, t5)):
                    acc[i] += t / 4
        print(f"threads={threads}: query {acc[0]:.1f} disk {acc[1]:.1f} put_hit {acc[2]:.1f} put_miss {acc[3]:.1f} "
              f"replace {acc[4]:.1f} ms (sum {sum(acc):.1f})")


if __name__ == "__main__":
    main()
