"""Names only (never called on the hot path)."""


def add_self_loops(*a, **k):
    raise NotImplementedError


def degree(*a, **k):
    raise NotImplementedError
