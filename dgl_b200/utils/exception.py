"""Moving an exception across a thread / process boundary.

Worker threads of the prefetching loaders and sampler processes cannot raise into the consumer directly. They capture
the failure as an :class:`ExceptionWrapper` — exception *type* plus the fully formatted traceback *text*, both
picklable — push it through the result queue, and the consumer calls :meth:`ExceptionWrapper.reraise`.
(Role of the reference's python/dgl/utils/exception.py.)"""
from __future__ import annotations

import sys
import traceback

__all__ = ["KeyErrorMessage", "ExceptionWrapper"]


class KeyErrorMessage(str):
    """A ``str`` whose ``repr`` is itself. ``KeyError.__str__`` shows ``repr(args[0])``, which would turn the embedded
    multi-line traceback into one escaped line; wrapping the message in this type keeps it readable."""

    def __repr__(self):
        return str.__str__(self)


def _rebuild(exc_type, text):
    """Instance of ``exc_type`` carrying ``text``; types whose constructor does not accept a single message fall back to
    ``RuntimeError`` (the text already names the original type)."""
    if exc_type is KeyError:
        return KeyError(KeyErrorMessage(text))
    if getattr(exc_type, "message", None):          # exception classes built around a ``message=`` keyword
        return exc_type(message=text)
    try:
        return exc_type(text)
    except TypeError:
        return RuntimeError(text)


class ExceptionWrapper:
    """Picklable record of an exception: ``exc_type``, formatted traceback ``exc_msg`` and a short ``where`` note
    ("in prefetcher thread", "in sampler worker 3") that is prepended when the error resurfaces."""

    __slots__ = ("exc_type", "exc_msg", "where")

    def __init__(self, exc_info=None, where="in background"):
        etype, evalue, tb = exc_info if exc_info is not None else sys.exc_info()
        self.exc_type = etype
        self.exc_msg = "".join(traceback.format_exception(etype, evalue, tb))
        self.where = where

    def __getstate__(self):
        return self.exc_type, self.exc_msg, self.where

    def __setstate__(self, state):
        self.exc_type, self.exc_msg, self.where = state

    def reraise(self):
        """Raise an exception of the recorded type whose message embeds the original traceback."""
        text = f"Caught {self.exc_type.__name__} {self.where}.\nOriginal {self.exc_msg}"
        raise _rebuild(self.exc_type, text)
