"""Format registry for chunk files (reference tools/distpartitioning/array_readwriter/): numpy, csv, parquet."""
import numpy as np

_REGISTRY = {}


def register_array_parser(name):
    def deco(cls):
        _REGISTRY[name] = cls
        return cls
    return deco


def get_array_parser(name="numpy", **kwargs):
    if name not in _REGISTRY:
        raise KeyError(f"unknown array format {name!r}; known: {sorted(_REGISTRY)}")
    return _REGISTRY[name](**kwargs)


@register_array_parser("numpy")
class NumpyArrayParser:
    def read(self, path):
        return np.load(path, mmap_mode="r")

    def write(self, path, arr):
        np.save(path, arr)


@register_array_parser("csv")
class CSVArrayParser:
    def __init__(self, delimiter=","):
        self.delimiter = delimiter

    def read(self, path):
        import pandas as pd
        return pd.read_csv(path, header=None, delimiter=self.delimiter).to_numpy()

    def write(self, path, arr):
        np.savetxt(path, arr, delimiter=self.delimiter, fmt="%s")


@register_array_parser("parquet")
class ParquetArrayParser:
    def read(self, path):
        import pyarrow.parquet as pq
        t = pq.read_table(path)
        cols = [t.column(i).to_numpy() for i in range(t.num_columns)]
        return np.stack(cols, 1) if len(cols) > 1 else cols[0]

    def write(self, path, arr):
        import pyarrow as pa
        import pyarrow.parquet as pq
        arr = np.asarray(arr)
        arr2 = arr.reshape(arr.shape[0], -1)
        pq.write_table(pa.table({f"c{i}": arr2[:, i] for i in range(arr2.shape[1])}), path)
