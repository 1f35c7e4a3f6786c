"""pcc_b200 — B200-native (sm_100a) point-set encoder hot path.

Host-side mirror of the reference's model interface (models/deep_sets.py,
models/graph_net.py) over the C-ABI library lib/libpcc.so.  See DESIGN.md.
"""
from .deep_sets import DeepSets, ResidualBlock
from .graph_net import GraphNet, GraphConv, knn_graph
from . import functional
from . import _lib

__all__ = ["DeepSets", "ResidualBlock", "GraphNet", "GraphConv", "knn_graph", "functional"]
