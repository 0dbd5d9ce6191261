"""Message-passing operators on DGLGraph with edge-softmax fusion (parity: /root/reference/python/dgl/mpops/)."""
from ..ops import edge_softmax  # noqa: F401


def u_mul_e_sum(g, x, w):
    from .. import ops
    return ops.u_mul_e_sum(g, x, w)


def fused_gat(g, feat_src, el, er, slope=0.2):
    """leaky_relu(el[u] + er[v]) -> edge softmax -> weighted sum (three sm_100a kernels)."""
    import torch.nn.functional as F
    from .. import ops
    e = F.leaky_relu(ops.u_add_v(g, el, er), slope)
    a = edge_softmax(g, e)
    return ops.u_mul_e_sum(g, feat_src, a)

from .edgewise import *  # noqa: F401,F403,E402
from . import edgewise  # noqa: F401,E402
