"""topicgcn_b200 — B200-native graph-convolution hot path of TopicGCN (drop-in for the reference's layer.py).

Importable as `topicgcn_b200` (the repo-root shim package points here; this directory's own name contains hyphens).
The compute backend is libtopicgcn.so (hand-written sm_100a CUDA behind the C-ABI of include/topicgcn.h); there is
no CPU or eager fallback.
"""
from . import _native
from ._native import TopicGCNError, LIB_PATH
from .csr import DeviceCSR, cached_csr
from .layer import GCN, GraphConvolution, Featureless
from .graph import CapturedTrainStep
from .ops import masked_cross_entropy, spmm
from . import optim
from . import ingest

__all__ = ["GCN", "GraphConvolution", "Featureless", "DeviceCSR", "cached_csr", "masked_cross_entropy", "spmm", "CapturedTrainStep",
           "optim", "ingest", "TopicGCNError", "LIB_PATH"]
__version__ = "0.1.0"
