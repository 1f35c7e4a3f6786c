"""Drop-in for /root/reference/models/graph_net.py (imported at train.py:8)."""
from pcc_b200.graph_net import GraphNet, GraphConv  # noqa: F401
