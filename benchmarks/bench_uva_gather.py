"""UVA row gather from pinned host memory: GB/s vs index order and thread budget (PCIe Gen5 x16 bound)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgl_b200.ops import cuda_struct as cs

n, F = 11_000_000, 128
feat = torch.randn(n, F).pin_memory()
idx = torch.randperm(n)[:1_000_000].cuda()


def timed(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it


res = []
for sort in (False, True):
    cs._uva_sort_min_rows[0] = 1 if sort else 1 << 62
    for thr in (10240, 40960, 148 * 1024, 0):
        cs.set_max_uva_threads(thr)
        ms = timed(lambda: cs.uva_index_select(feat, idx))
        r = dict(bench="uva_gather", sorted=sort, max_threads=thr, ms=ms, GBps=idx.shape[0] * F * 4 / ms / 1e6)
        print(json.dumps(r), flush=True)
        res.append(r)
ref = feat[idx.cpu()]
cs._uva_sort_min_rows[0] = 1
assert torch.equal(cs.uva_index_select(feat, idx).cpu(), ref)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
