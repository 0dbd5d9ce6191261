"""Reference module path ``dgl.distgnn.tools.tools``: the names below are implemented in
dgl_b200.distgnn, dgl_b200.partition, dgl_b200 ... and re-exported lazily."""
from ..._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.distgnn', 'dgl_b200.partition', 'dgl_b200'],
                                          ['download_proteins', 'load_proteins', 'proteins_mtx2dgl', 'rep_per_node', 'save'])
