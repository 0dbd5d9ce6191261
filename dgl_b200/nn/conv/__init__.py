"""Graph convolution layers (parity: /root/reference/python/dgl/nn/pytorch/conv/__init__.py)."""
from .attention import AGNNConv, DotGatConv, EdgeGATConv, EGATConv, GATConv, GATv2Conv
from .graphconv import EdgeWeightNorm, GraphConv
from .misc import AtomicConv, DenseChebConv, DenseGraphConv, DenseSAGEConv, DGNConv, PNAConv
from .relational import HeteroGraphConv, HGTConv, RelGraphConv
from .sageconv import SAGEConv
from .spectral import (APPNPConv, CFConv, ChebConv, EdgeConv, EGNNConv, GatedGCNConv, GatedGraphConv, GCN2Conv,
                       GINConv, GINEConv, GMMConv, NNConv, SGConv, TAGConv, TWIRLSConv)

# cuGraph-ops backed layers of the reference map onto the native sm_100a paths here
CuGraphSAGEConv = SAGEConv
CuGraphGATConv = GATConv
CuGraphRelGraphConv = RelGraphConv

from .extra import GroupRevRes, TWIRLSUnfoldingAndAttention  # noqa: E402
from .parts import (AX, MLP, Attention, CuGraphBaseConv, D_power_bias_X, D_power_X, InvertibleCheckpoint,  # noqa: E402,F401
                    Propagate, PropagateNoPrecond, normalized_AX)
from .misc import (DGNConvTower, PNAConvTower, RadialPooling, msg_func, reduce_func, aggregate_dir_av, aggregate_dir_dx, aggregate_max,  # noqa: E402,F401
                   aggregate_mean, aggregate_min, aggregate_moment_3, aggregate_moment_4, aggregate_moment_5,
                   aggregate_std, aggregate_sum, aggregate_var, scale_amplification, scale_attenuation,
                   scale_identity)
from .spectral import ShiftedSoftplus  # noqa: E402,F401

__all__ = ["GroupRevRes", "TWIRLSUnfoldingAndAttention", "GraphConv", "EdgeWeightNorm", "SAGEConv", "GATConv", "GATv2Conv", "DotGatConv", "AGNNConv", "EGATConv",
           "EdgeGATConv", "GINConv", "GINEConv", "SGConv", "TAGConv", "APPNPConv", "ChebConv", "GCN2Conv",
           "GatedGraphConv", "GatedGCNConv", "CFConv", "NNConv", "GMMConv", "EdgeConv", "EGNNConv", "TWIRLSConv",
           "RelGraphConv", "HeteroGraphConv", "HGTConv", "DenseGraphConv", "DenseSAGEConv", "DenseChebConv", "PNAConv",
           "DGNConv", "AtomicConv", "CuGraphSAGEConv", "CuGraphGATConv", "CuGraphRelGraphConv"]
