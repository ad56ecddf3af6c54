"""acm_gnn_b200 -- Blackwell-native ACM graph-convolution layer (drop-in for the
``GraphConvolution`` module of SitaoLuan/ACM-GNN).  Hot path = hand-written sm_100a CUDA
kernels in libacm_b200.so behind a C ABI (include/acm_b200.h); this package is the host
side that mirrors the reference's Python module interface."""
from . import _lib  # noqa: F401
from .functional import AcmLayerFunction, LayerConfig, StagedInput, padded_width, stage_input  # noqa: F401
from .layers import MLP, GraphConvolution  # noqa: F401
from .models import GCN  # noqa: F401
from .operator import AcmOperator, CsrMatrix, cached_operator  # noqa: F401
from .graphed import GraphedForward, GraphedTrainStep  # noqa: F401

__version__ = "0.1.1"
