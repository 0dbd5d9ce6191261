"""Internal helpers of GraphBolt (parity: /root/reference/python/dgl/graphbolt/internal/)."""
from .item_sampler_utils import *  # noqa: F401,F403
from .sample_utils import *  # noqa: F401,F403
from .utils import *  # noqa: F401,F403
