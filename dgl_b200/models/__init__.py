"""Model zoo + fused training engines."""
from .graphsage import FusedSAGETrainer, GraphSAGE  # noqa: F401
from .gcn import GCN  # noqa: F401
from .gat import GAT  # noqa: F401
from .rgcn import RGCN  # noqa: F401
