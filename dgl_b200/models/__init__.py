"""Model zoo + fused training engines."""
from .graphsage import FusedSAGETrainer, GraphSAGE  # noqa: F401
from .gcn import GCN  # noqa: F401
from .gat import GAT  # noqa: F401
from .rgcn import RGCN  # noqa: F401
from .zoo import (APPNP, DGI, GATv2, GCNII, GIN, HGT, SGC, TAGCN, ChebNet, HeteroRGCN, JKNet, LabelPropagationModel,  # noqa: F401
                  LinkPredictor)
from .zoo_extra import (ARMA, DAGNN, DGCNN, GAE, GCMC, GRACE, GRAND, HAN, LGNN, MPNN, SIGN, STGCN, CompGCN,  # noqa: F401
                        CorrectAndSmooth, DeeperGCN, EdgeConvNet, GNNFiLM, MixHop, MoNet, Node2VecModel, PinSAGEModel,
                        SAGPoolNet, drnl_node_labeling, sym_norm_propagate)
