"""Stub of torch_sparse (imported by GNN_Models.py:12 / GNN_Layers.py:10, unused on the hot path)."""


class SparseTensor:  # pragma: no cover
    pass
