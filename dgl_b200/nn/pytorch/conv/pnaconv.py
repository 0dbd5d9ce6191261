"""Reference module path ``dgl.nn.pytorch.conv.pnaconv``: the names below are implemented in
dgl_b200.nn.pytorch.conv, dgl_b200.nn.pytorch, dgl_b200.nn ... and re-exported lazily."""
from ...._pathshim import reexport

__getattr__, __dir__, __all__ = reexport(__name__, ['dgl_b200.nn.pytorch.conv', 'dgl_b200.nn.pytorch', 'dgl_b200.nn', 'dgl_b200.nn.conv', 'dgl_b200'],
                                          ['PNAConv', 'PNAConvTower', 'aggregate_max', 'aggregate_mean', 'aggregate_min', 'aggregate_moment_3', 'aggregate_moment_4', 'aggregate_moment_5', 'aggregate_std', 'aggregate_sum', 'aggregate_var', 'scale_amplification', 'scale_attenuation', 'scale_identity'])
