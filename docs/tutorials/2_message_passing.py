"""Blitz 2 - writing your own GNN layer with message passing (reference tutorials/blitz/3_message_passing.py).

``update_all(message_fn, reduce_fn)`` with BUILT-IN functions lowers to one fused g-SpMM kernel; user-defined functions
(``lambda edges: ...``) run through the degree-bucketing executor and are the flexible (slower) fallback."""
import torch
import torch.nn as nn

import dgl_b200 as dgl
import dgl_b200.function as fn


class MySAGEConv(nn.Module):
    """h_v' = W [h_v || mean_{u in N(v)} h_u]"""

    def __init__(self, in_feat, out_feat):
        super().__init__()
        self.linear = nn.Linear(in_feat * 2, out_feat)

    def forward(self, g, h):
        with g.local_scope():
            g.ndata["h"] = h
            g.update_all(fn.copy_u("h", "m"), fn.mean("m", "h_N"))       # built-ins -> one g-SpMM launch
            return self.linear(torch.cat([h, g.ndata["h_N"]], 1))


g = dgl.rand_graph(100, 600)
h = torch.randn(100, 8)
print(MySAGEConv(8, 4)(g, h).shape)

# edge-weighted aggregation: u_mul_e + sum
g.edata["w"] = torch.rand(600, 1)
g.ndata["h"] = h
g.update_all(fn.u_mul_e("h", "w", "m"), fn.sum("m", "h_sum"))

# the same with user-defined functions (degree bucketing)
g.update_all(lambda edges: {"m": edges.src["h"] * edges.data["w"]}, lambda nodes: {"h_udf": nodes.mailbox["m"].sum(1)})
assert torch.allclose(g.ndata["h_sum"], g.ndata["h_udf"], atol=1e-5)

# edge-wise computation: attention-like scores with g-SDDMM
g.apply_edges(fn.u_dot_v("h", "h", "score"))
print(g.edata["score"].shape)
