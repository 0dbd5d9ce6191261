"""Package logger.

One logger, ``"dgl-core"`` (the name the reference uses, python/dgl/logging.py, so existing logging configs keep
matching), with a stream handler that tags records with process and thread — useful because samplers, prefetchers and
RPC servers all log concurrently. INFO by default; ``DGL_LOG_DEBUG=1`` in the environment or
:func:`enable_verbose_logging` switches to DEBUG (and sets ``DMLC_LOG_DEBUG`` for native code that consults it)."""
import logging as _logging
import os as _os

LOGGER_NAME = "dgl-core"
_FORMAT = "%(asctime)s %(levelname)s %(processName)s-%(threadName)s %(name)s: %(message)s"


def get_logger():
    """The configured package logger (handler attached exactly once)."""
    log = _logging.getLogger(LOGGER_NAME)
    if not any(getattr(h, "_dglb200", False) for h in log.handlers):
        handler = _logging.StreamHandler()
        handler._dglb200 = True
        handler.setLevel(_logging.DEBUG)
        handler.setFormatter(_logging.Formatter(_FORMAT))
        log.addHandler(handler)
        log.propagate = False
        log.setLevel(_logging.INFO)
    return log


def enable_verbose_logging():
    """DEBUG level for the package logger and for native components reading ``DMLC_LOG_DEBUG``."""
    _os.environ["DMLC_LOG_DEBUG"] = "1"
    log = get_logger()
    log.setLevel(_logging.DEBUG)
    log.debug("logging level set to DEBUG")


get_logger()
if _os.environ.get("DGL_LOG_DEBUG") == "1":
    enable_verbose_logging()
