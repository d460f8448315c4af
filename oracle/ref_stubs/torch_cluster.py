"""Stub of torch_cluster (reference import: GNN_Models.py:5)."""
from torch_geometric.nn import radius_graph  # noqa: F401
