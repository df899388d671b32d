import torch

from gnnb200.utils import (batched_negative_sampling, coalesce, negative_sampling, subgraph,  # noqa: F401
                           to_undirected)


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], None


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    return torch.bincount(index, minlength=n).to(dtype or torch.float32)


def to_networkx(*args, **kwargs):
    raise NotImplementedError('to_networkx is offline preprocessing, outside the hot path')
