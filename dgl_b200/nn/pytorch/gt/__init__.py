from ...gt import *  # noqa: F401,F403
from ... import gt as _m  # noqa: E402
globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith('__')})
