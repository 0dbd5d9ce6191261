"""Per-op trace (dgl_b200.utils.trace) of the GraphBolt UVA pipeline: where a step's time goes."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.argv = ['x']
import torch, torch.nn.functional as F
import bench_uva_pipeline as b
class A: pass
args = A(); args.batch = 1024; args.fanouts = [10, 10, 10]; args.cache_mb = 0; args.workers = 0; args.resident = "host"; args.amp = False
n, u, v, feat, labels, train = b.make_graph(0.1)
model, opt, it, ce, ci = b.run_ours(args, n, u, v, feat, labels, train)
import dgl_b200 as dgl
def step():
    with dgl.utils.trace.range("next_batch"):
        blocks, x, y = next(it)
    with dgl.utils.trace.range("model_fwd_bwd_opt"):
        loss = F.cross_entropy(model(blocks, x), y)
        opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(5): step()
with dgl.utils.trace.trace() as t:
    for _ in range(20): step()
print(t.report())
