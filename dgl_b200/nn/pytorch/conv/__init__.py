"""reference nn/pytorch/conv: every convolution layer."""
from ...conv import *  # noqa: F401,F403
from ... import conv as _c  # noqa: E402
globals().update({k: getattr(_c, k) for k in dir(_c) if not k.startswith("__")})
