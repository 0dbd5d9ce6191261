"""reference nn/pytorch/factory.py: KNNGraph, SegmentedKNNGraph, RadiusGraph."""
from .. import KNNGraph, RadiusGraph, SegmentedKNNGraph  # noqa: F401
