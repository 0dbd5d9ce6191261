"""``dgl.nn.functional`` (parity: /root/reference/python/dgl/nn/functional/__init__.py)."""
from ..ops import edge_softmax  # noqa: F401
