"""Stub of torch_scatter (imported by GNN_Models.py:11, unused on the hot path)."""


def scatter(*a, **k):
    raise NotImplementedError
