"""DGLDataset base class with processed-data caching (parity: /root/reference/python/dgl/data/dgl_dataset.py)."""
from __future__ import annotations

import abc
import hashlib
import os
import traceback


def get_download_dir():
    default_dir = os.path.join(os.path.expanduser("~"), ".dgl")
    dirname = os.environ.get("DGL_DOWNLOAD_DIR", default_dir)
    os.makedirs(dirname, exist_ok=True)
    return dirname


class DGLDataset:
    def __init__(self, name, url=None, raw_dir=None, save_dir=None, hash_key=(), force_reload=False, verbose=False,
                 transform=None):
        self._name = name
        self._url = url
        self._force_reload = force_reload
        self._verbose = verbose
        self._hash_key = hash_key
        self._hash = self._get_hash()
        self._transform = transform
        self._raw_dir = get_download_dir() if raw_dir is None else raw_dir
        self._save_dir = self._raw_dir if save_dir is None else save_dir
        self._load()

    def download(self):
        """No network in this environment: builtin datasets synthesise their data in ``process``."""

    def save(self):
        pass

    def load(self):
        pass

    @abc.abstractmethod
    def process(self):
        pass

    def has_cache(self):
        return False

    def _download(self):
        if os.path.exists(self.raw_path):
            return
        os.makedirs(self.raw_dir, exist_ok=True)
        self.download()

    def _load(self):
        load_flag = not self._force_reload and self.has_cache()
        if load_flag:
            try:
                self.load()
                if self.verbose:
                    print("Done loading data from cached files.")
            except KeyboardInterrupt:
                raise
            except Exception:  # noqa: BLE001
                load_flag = False
                if self.verbose:
                    print(traceback.format_exc())
                    print("Loading from cache failed, re-processing.")
        if not load_flag:
            self._download()
            self.process()
            os.makedirs(self.save_path, exist_ok=True)
            self.save()
            if self.verbose:
                print("Done saving data into cached files.")

    def _get_hash(self):
        hash_func = hashlib.sha1()
        hash_func.update(str(self._hash_key).encode("utf-8"))
        return hash_func.hexdigest()[:8]

    @property
    def url(self):
        return self._url

    @property
    def name(self):
        return self._name

    @property
    def raw_dir(self):
        return self._raw_dir

    @property
    def raw_path(self):
        return os.path.join(self.raw_dir, self.name)

    @property
    def save_dir(self):
        return self._save_dir

    @property
    def save_path(self):
        return os.path.join(self.save_dir, self.name + "_" + self._hash)

    @property
    def verbose(self):
        return self._verbose

    @property
    def hash(self):
        return self._hash

    @abc.abstractmethod
    def __getitem__(self, idx):
        pass

    @abc.abstractmethod
    def __len__(self):
        pass

    def __repr__(self):
        return f'Dataset("{self.name}", num_graphs={len(self)}, save_path={self.save_path})'


class DGLBuiltinDataset(DGLDataset):
    pass
