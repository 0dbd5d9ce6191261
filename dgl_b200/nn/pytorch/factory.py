"""reference nn/pytorch/factory.py: KNNGraph, SegmentedKNNGraph, RadiusGraph."""
from .. import KNNGraph, RadiusGraph, SegmentedKNNGraph  # noqa: F401
from ...transforms.functional import pairwise_squared_distance  # noqa: F401,E402
