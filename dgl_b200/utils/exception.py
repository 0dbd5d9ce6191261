"""Carrying an exception (with its traceback text) from a worker thread/process to the consumer
(parity: /root/reference/python/dgl/utils/exception.py, used by the prefetching dataloader)."""
from __future__ import annotations

import sys
import traceback

__all__ = ["KeyErrorMessage", "ExceptionWrapper"]


class KeyErrorMessage(str):
    """``KeyError`` reprs its argument; this str subclass reprs as itself so multi-line messages stay readable."""

    def __repr__(self):
        return str(self)


class ExceptionWrapper:
    """Snapshot of ``sys.exc_info()`` that can be pickled, queued and re-raised elsewhere."""

    def __init__(self, exc_info=None, where="in background"):
        if exc_info is None:
            exc_info = sys.exc_info()
        self.exc_type = exc_info[0]
        self.exc_msg = "".join(traceback.format_exception(*exc_info))
        self.where = where

    def reraise(self):
        """Raise the captured exception type with the original traceback text folded into the message."""
        msg = f"Caught {self.exc_type.__name__} {self.where}.\nOriginal {self.exc_msg}"
        if self.exc_type is KeyError:
            msg = KeyErrorMessage(msg)
        elif getattr(self.exc_type, "message", None):
            raise self.exc_type(message=msg)
        try:
            exc = self.exc_type(msg)
        except TypeError:
            # exception types with mandatory extra constructor arguments
            raise RuntimeError(msg) from None
        raise exc
