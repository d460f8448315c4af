"""Stub of torch_geometric.nn: MessagePassing (aggr='add', flow source->target) and radius_graph.

Upstream semantics restated (PyG 2.5.1 `MessagePassing.propagate`, not vendored in the reference):
  *_j <- kwarg[edge_index[0]], *_i <- kwarg[edge_index[1]], other kwargs passed through to
  `message`, output = scatter-add of the messages onto edge_index[1] with dim_size = x.size(0).
On CPU `index_add` accumulates sequentially in edge order (the order parity relies on).
"""
import inspect
import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", **kw):
        super().__init__()
        assert aggr == "add"

    def jittable(self):
        return self

    def propagate(self, edge_index, size=None, **kw):
        names = list(inspect.signature(self.message).parameters)
        args = {}
        for n in names:
            if n.endswith("_i"):
                args[n] = kw[n[:-2]].index_select(0, edge_index[1])
            elif n.endswith("_j"):
                args[n] = kw[n[:-2]].index_select(0, edge_index[0])
            else:
                args[n] = kw[n]
        msg = self.message(**args)
        ref = kw["x"] if "x" in kw else next(iter(kw.values()))
        out = torch.zeros((ref.size(0),) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        return out.index_add(0, edge_index[1], msg)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target", **kw):
    """Dense restatement of torch_cluster 1.6.3 radius_graph: strict `<` on the squared fp32
    distance, same-batch pairs only, ordered by target then source; row0 = source, row1 = target."""
    n = x.size(0)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long, device=x.device)
    d = x.unsqueeze(1) - x.unsqueeze(0)          # d[i, j] = x_i - x_j
    d2 = (d * d).sum(-1)
    mask = (d2 < r * r) & (batch.unsqueeze(1) == batch.unsqueeze(0))
    if not loop:
        mask &= ~torch.eye(n, dtype=torch.bool, device=x.device)
    tgt, src = mask.nonzero(as_tuple=True)       # row-major => sorted by target then source
    assert int(mask.sum(1).max()) <= max_num_neighbors
    return torch.stack([src, tgt])
