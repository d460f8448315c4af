"""Stub of torch_geometric.transforms.RadiusGraph (reference use: GNN_Models.py:73-75, 267-269)."""
from .nn import radius_graph


class RadiusGraph:
    def __init__(self, r, loop=False, max_num_neighbors=32, flow="source_to_target"):
        self.r, self.loop, self.max_num_neighbors = r, loop, max_num_neighbors

    def __call__(self, data):
        batch = getattr(data, "batch", None)
        data.edge_index = radius_graph(data.pos, self.r, batch, self.loop, self.max_num_neighbors)
        return data
