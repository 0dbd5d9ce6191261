"""Datasets (parity: /root/reference/python/dgl/data/). No network access in this environment, so builtin
datasets are synthesised with the published shapes (see synthetic.py); loaders for on-disk data are kept."""
from .synthetic import SHAPES, rmat_edges, rmat_graph, shaped_dataset  # noqa: F401


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    import importlib
    if name == "sbm":            # as in the reference, ``dgl.data.sbm`` is the module; the generator is ``dgl.data.sbm.sbm``
        return importlib.import_module(f"{__name__}.sbm")
    if name == "citegrh":        # legacy alias of the citation-graph module (reference data/__init__.py)
        return importlib.import_module(f"{__name__}.citation_graph")
    if name in ("dgl_binary", "graph_serialize", "compat", "more", "builtin", "extra", "utils", "dgl_dataset", "csv_dataset", "raw_readers"):
        return importlib.import_module(f"{__name__}.{name}")
    for modname in ("dgl_dataset", "graph_serialize", "more", "builtin", "csv_dataset", "utils", "extra", "compat"):
        mod = importlib.import_module(f"{__name__}.{modname}")
        if hasattr(mod, name):
            return getattr(mod, name)
    raise AttributeError(name)
