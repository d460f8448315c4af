"""Stand-in for torch_geometric (absent in this image). TEST INFRASTRUCTURE ONLY.

Provides exactly the names the reference's MachineLearning/GNN_Models.py:5-12 and
GNN_Layers.py:5-10 import, so those files import UNMODIFIED from /root/reference.
Only `MessagePassing.propagate` and `radius_graph` carry semantics.
"""
from . import nn, transforms, utils, data  # noqa: F401
