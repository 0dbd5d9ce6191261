"""Model zoo + fused training engines."""
from .graphsage import FusedSAGETrainer, GraphSAGE  # noqa: F401
