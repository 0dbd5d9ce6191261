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

__all__ = ["GroupRevRes", "TWIRLSUnfoldingAndAttention", "GraphConv", "EdgeWeightNorm", "SAGEConv", "GATConv", "GATv2Conv", "DotGatConv", "AGNNConv", "EGATConv",
           "EdgeGATConv", "GINConv", "GINEConv", "SGConv", "TAGConv", "APPNPConv", "ChebConv", "GCN2Conv",
           "GatedGraphConv", "GatedGCNConv", "CFConv", "NNConv", "GMMConv", "EdgeConv", "EGNNConv", "TWIRLSConv",
           "RelGraphConv", "HeteroGraphConv", "HGTConv", "DenseGraphConv", "DenseSAGEConv", "DenseChebConv", "PNAConv",
           "DGNConv", "AtomicConv", "CuGraphSAGEConv", "CuGraphGATConv", "CuGraphRelGraphConv"]
