"""Per-kernel time table of one training epoch of a bench_configs.py config (torch.profiler, CUDA activities)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F
import bench_configs as bc

cfg, impl = sys.argv[1], sys.argv[2]
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
dgl = bc.load_impl(impl)
model, fwd, labels, idx, c = getattr(bc, cfg)(dgl, scale, torch.device("cuda"), torch.float32)
opt = torch.optim.Adam(model.parameters(), lr=5e-3)


def step():
    logits = fwd()
    loss = F.cross_entropy(logits[idx].float(), labels[idx] if labels.shape[0] == logits.shape[0] else labels)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
