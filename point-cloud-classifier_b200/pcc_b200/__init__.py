"""pcc_b200 — B200-native (sm_100a) point-set encoder hot path.

Host-side mirror of the reference's model interface (models/deep_sets.py,
models/graph_net.py) over the C-ABI library lib/libpcc.so.  See DESIGN.md.
"""
from .deep_sets import DeepSets, ResidualBlock
from .graph_net import GraphNet, GraphConv, KnnGraphNet, knn_graph, gaussian_edge_weights
from . import functional
from .optim import FusedAdam
from . import _lib

__all__ = ["DeepSets", "ResidualBlock", "GraphNet", "GraphConv", "KnnGraphNet", "knn_graph", "gaussian_edge_weights", "FusedAdam", "functional"]

import os as _os

if _os.environ.get("PCC_DENSE", "").lower() == "tf32":   # single-TF32 dense layers (functional.set_dense_precision)
    try:
        functional.set_dense_precision("tf32")
    except Exception:  # library not built yet: the first real call reports it
        pass
