"""Small filesystem helpers shared by the command-line tools (reference: tools/files.py)."""
import contextlib
import os


@contextlib.contextmanager
def setdir(path):
    """Create ``path`` if needed, make it the working directory inside the block and restore the previous one after."""
    os.makedirs(path, exist_ok=True)
    previous = os.getcwd()
    os.chdir(path)
    try:
        yield path
    finally:
        os.chdir(previous)
