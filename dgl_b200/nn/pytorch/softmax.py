"""reference nn/pytorch/softmax.py"""
from ...ops import edge_softmax  # noqa: F401
