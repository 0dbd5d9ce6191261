"""Dataset base class with an on-disk cache of the processed form.

Contract (same as the reference's ``dgl.data.DGLDataset``, python/dgl/data/dgl_dataset.py): subclasses implement
``process`` (build the graphs in memory) and usually ``save`` / ``load`` / ``has_cache``; construction then runs

    cached and not force_reload ?  load()  :  download() -> process() -> save()

A failing ``load`` (stale or truncated cache) falls back to the second branch. The cache directory is
``<save_dir>/<name>_<hash>`` where the hash digests ``hash_key`` — the constructor arguments that change the processed
result — so differently configured instances never share a cache."""
from __future__ import annotations

import hashlib
import os
import traceback


def get_download_dir():
    """Root for raw and processed data: ``$DGL_DOWNLOAD_DIR`` or ``~/.dgl`` (created on first use)."""
    path = os.environ.get("DGL_DOWNLOAD_DIR") or os.path.join(os.path.expanduser("~"), ".dgl")
    os.makedirs(path, exist_ok=True)
    return path


def _readonly(attr, doc):
    return property(lambda self: getattr(self, attr), doc=doc)


class DGLDataset:
    url = _readonly("_url", "Where the raw data comes from (may be None).")
    name = _readonly("_name", "Dataset name; also the raw sub-directory and the cache prefix.")
    raw_dir = _readonly("_raw_dir", "Directory holding raw inputs.")
    save_dir = _readonly("_save_dir", "Directory holding processed caches.")
    verbose = _readonly("_verbose", "Whether progress is printed.")
    hash = _readonly("_hash", "8-hex-digit digest of ``hash_key``.")

    def __init__(self, name, url=None, raw_dir=None, save_dir=None, hash_key=(), force_reload=False, verbose=False,
                 transform=None):
        self._name, self._url, self._transform = name, url, transform
        self._force_reload, self._verbose = force_reload, verbose
        self._hash_key = hash_key
        self._hash = hashlib.sha1(str(hash_key).encode("utf-8")).hexdigest()[:8]
        self._raw_dir = raw_dir if raw_dir is not None else get_download_dir()
        self._save_dir = save_dir if save_dir is not None else self._raw_dir
        self._load()

    # ---- hooks for subclasses ---------------------------------------------------------------------------------------
    def download(self):
        """Fetch raw data into :attr:`raw_path`. There is no network in this environment: builtin datasets synthesise
        their content in :meth:`process` instead."""

    def process(self):
        """Build the in-memory dataset. Must be overridden."""
        raise NotImplementedError

    def save(self):
        """Write the processed form under :attr:`save_path` (optional)."""

    def load(self):
        """Read what :meth:`save` wrote (optional)."""

    def has_cache(self):
        """Whether :meth:`load` can be used instead of :meth:`process`."""
        return False

    def __getitem__(self, idx):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    # ---- lifecycle ------------------------------------------------------------------------------------------------------
    @property
    def raw_path(self):
        return os.path.join(self._raw_dir, self._name)

    @property
    def save_path(self):
        return os.path.join(self._save_dir, f"{self._name}_{self._hash}")

    def _say(self, text):
        if self._verbose:
            print(text)

    def _download(self):
        if not os.path.exists(self.raw_path):
            os.makedirs(self._raw_dir, exist_ok=True)
            self.download()

    def _restore(self):
        """Try the cache; False when there is none or it cannot be read."""
        if self._force_reload or not self.has_cache():
            return False
        try:
            self.load()
        except KeyboardInterrupt:
            raise
        except Exception:  # noqa: BLE001 - any failure means "rebuild"
            self._say(traceback.format_exc() + "Loading from cache failed, re-processing.")
            return False
        self._say("Done loading data from cached files.")
        return True

    def _load(self):
        if self._restore():
            return
        self._download()
        self.process()
        os.makedirs(self.save_path, exist_ok=True)
        self.save()
        self._say("Done saving data into cached files.")

    def _get_hash(self):
        return self._hash

    def __repr__(self):
        return f'Dataset("{self.name}", num_graphs={len(self)}, save_path={self.save_path})'


class DGLBuiltinDataset(DGLDataset):
    """Base of the datasets shipped with the package (same lifecycle; kept as a separate type for ``isinstance``)."""
