"""Module-style transforms (filled in by transforms/module_impl.py)."""
from .module_impl import *  # noqa: F401,F403
