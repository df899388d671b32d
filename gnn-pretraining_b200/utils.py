"""Graph utilities with the PyG names the reference imports from ``torch_geometric.utils``
(reference src/pretrain/tasks.py:10, src/pretrain/augmentations.py:5): the device work
(symmetrise + coalesce) runs on the gnnb200 kernels; the samplers that draw from Python's
``random`` stay on the host so their streams match the reference's bit for bit."""
import random
from typing import Optional

import numpy as np
import torch
from torch import Tensor

from . import ops


def _num_nodes(edge_index: Tensor, num_nodes: Optional[int]) -> int:
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max()) + 1 if edge_index.numel() else 0


def coalesce(edge_index: Tensor, num_nodes: Optional[int] = None) -> Tensor:
    """Sort columns by row*N+col and drop duplicates (device radix sort + compaction)."""
    n = _num_nodes(edge_index, num_nodes)
    if edge_index.size(1) == 0:
        return edge_index
    out, count = ops.coalesce(edge_index, n)
    return out[:, :int(count)]        # one sync: the output width is data dependent (as upstream)


def to_undirected(edge_index: Tensor, edge_attr=None, num_nodes: Optional[int] = None, reduce: str = 'add') -> Tensor:
    """coalesce(cat([ei, ei.flip(0)]))."""
    n = _num_nodes(edge_index, num_nodes)
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    return coalesce(both, n)


def subgraph(subset: Tensor, edge_index: Tensor, edge_attr=None, relabel_nodes: bool = False,
             num_nodes: Optional[int] = None):
    """SURVEY.md App. A.8 (host-side augmentation helper; order preserving)."""
    n = _num_nodes(edge_index, num_nodes)
    keep_node = torch.zeros(n, dtype=torch.bool, device=edge_index.device)
    keep_node[subset] = True
    ei = edge_index[:, keep_node[edge_index[0]] & keep_node[edge_index[1]]]
    if relabel_nodes:
        new_id = torch.full((n,), -1, dtype=torch.long, device=edge_index.device)
        new_id[subset] = torch.arange(subset.numel(), device=edge_index.device)
        ei = new_id[ei]
    return ei, None


def _draw(population: int, k: int) -> Tensor:
    if population <= k:
        return torch.arange(population)
    return torch.tensor(random.sample(range(population), k))


def negative_sampling(edge_index: Tensor, num_nodes: Optional[int] = None, num_neg_samples: Optional[int] = None,
                      method: str = 'sparse', force_undirected: bool = False) -> Tensor:
    """SURVEY.md App. A.5, the 'sparse' directed branch the reference uses.  Small graphs take the
    deterministic arange branch (all non-edges in ascending order); large ones consume Python's
    `random` exactly like upstream."""
    if method != 'sparse' or force_undirected:
        raise NotImplementedError('only the branch used by the reference is provided')
    n = _num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    off_diag = row != col
    row, col = row[off_diag], col[off_diag]
    code = row * (n - 1) + torch.where(row < col, col - 1, col)
    population = n * n - n
    if code.numel() >= population:
        return edge_index.new_empty((2, 0))
    want = edge_index.size(1) if num_neg_samples is None else num_neg_samples
    p_neg = 1.0 - code.numel() / population
    k = int(1.1 * want / p_neg)
    taken = code.cpu().numpy()
    found = None
    for _ in range(3):
        cand = _draw(population, k)
        reject = np.isin(cand.numpy(), taken)
        if found is not None:
            reject |= np.isin(cand.numpy(), found.cpu().numpy())
        cand = cand[torch.from_numpy(~reject)].to(edge_index.device)
        found = cand if found is None else torch.cat([found, cand])
        if found.numel() >= want:
            found = found[:want]
            break
    r = found.div(n - 1, rounding_mode='floor')
    c = found % (n - 1)
    c = torch.where(r <= c, c + 1, c)
    return torch.stack([r, c], dim=0)


def batched_negative_sampling(edge_index: Tensor, batch: Tensor, num_neg_samples: Optional[int] = None,
                              method: str = 'sparse', force_undirected: bool = False) -> Tensor:
    """per graph, every graph with the same quota.  The per-graph splits are
    computed once on the host (one transfer) instead of one sync per graph."""
    counts = torch.bincount(batch)
    sizes = counts.tolist()
    starts = (counts.cumsum(0) - counts).tolist()
    per_graph = torch.bincount(batch[edge_index[0]]).tolist() if edge_index.size(1) else []
    ei_host = edge_index.cpu()
    out, at = [], 0
    for g, e_g in enumerate(per_graph):
        piece = ei_host[:, at:at + e_g] - starts[g]
        at += e_g
        out.append(negative_sampling(piece, sizes[g], num_neg_samples, method, force_undirected) + starts[g])
    if not out:
        return edge_index.new_empty((2, 0))
    return torch.cat(out, dim=1).to(edge_index.device)
