"""ORACLE (test infrastructure, never shipped): torch_geometric.utils restated in pure PyTorch.

PyG (`torch-geometric>=2.3.0`, unpinned in /root/reference/requirements.txt:3; the
authors' VMs ran 2.6.x, vm_execution_scripts/vm_setup.md:244-247) is NOT vendored in
the reference and is not installable here.  Every function below restates the
published upstream algorithm; parity against real
PyG is therefore UNPINNED.  Call sites that anchor the behaviour:
  to_undirected / batched_negative_sampling  src/pretrain/tasks.py:10,107-111
  subgraph                                   src/pretrain/augmentations.py:5,56
  negative_sampling                          src/data/data_setup.py:12,137
"""
import random
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor


def maybe_num_nodes(edge_index: Tensor, num_nodes: Optional[int] = None) -> int:
    if num_nodes is not None:
        return num_nodes
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0


def _broadcast(index: Tensor, ref: Tensor, dim: int) -> Tensor:
    shape = [1] * ref.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(ref)


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
            reduce: str = 'sum') -> Tensor:
    """App. A.2.  sum/mean via scatter_add_ (edge-order sequential on CPU), max via the
    native scatter_reduce_('amax', include_self=False) on a ZERO-initialised output."""
    if index.dim() != 1:
        raise ValueError("index must be one-dimensional")
    dim = src.dim() + dim if dim < 0 else dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0   # device sync, as upstream
    size = list(src.size())
    size[dim] = dim_size
    if reduce in ('sum', 'add'):
        return src.new_zeros(size).scatter_add_(dim, _broadcast(index, src, dim), src)
    if reduce == 'mean':
        count = src.new_zeros(dim_size)
        count.scatter_add_(0, index, src.new_ones(src.size(dim)))
        count = count.clamp(min=1)
        out = src.new_zeros(size).scatter_add_(dim, _broadcast(index, src, dim), src)
        return out / _broadcast(count, out, dim)
    if reduce in ('max', 'amax'):
        return src.new_zeros(size).scatter_reduce_(
            dim, _broadcast(index, src, dim), src, reduce='amax', include_self=False)
    raise ValueError(f"unsupported reduce {reduce!r}")


def degree(index: Tensor, num_nodes: Optional[int] = None, dtype=None) -> Tensor:
    n = maybe_num_nodes(index, num_nodes)
    out = torch.zeros((n,), dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, out.new_ones(index.size(0)))


def cumsum(x: Tensor) -> Tensor:
    out = x.new_zeros(x.size(0) + 1)
    torch.cumsum(x, 0, out=out[1:])
    return out


def coalesce(edge_index: Tensor, num_nodes: Optional[int] = None) -> Tensor:
    """App. A.4: sort by row*N+col, drop duplicates."""
    nnz = edge_index.size(1)
    n = maybe_num_nodes(edge_index, num_nodes)
    key = edge_index.new_empty(nnz + 1)
    key[0] = -1
    key[1:] = edge_index[0] * n + edge_index[1]
    if nnz > 1 and not bool((key[2:] >= key[1:-1]).all()):
        key[1:], perm = key[1:].sort(stable=True)
        edge_index = edge_index[:, perm]
    keep = key[1:] > key[:-1]
    if bool(keep.all()):
        return edge_index
    return edge_index[:, keep]


def to_undirected(edge_index: Tensor, edge_attr=None, num_nodes: Optional[int] = None,
                  reduce: str = 'add') -> Tensor:
    row, col = edge_index[0], edge_index[1]
    both = torch.stack([torch.cat([row, col]), torch.cat([col, row])], dim=0)
    return coalesce(both, num_nodes)


def remove_self_loops(edge_index: Tensor, edge_attr=None) -> Tuple[Tensor, None]:
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], None


def subgraph(subset: Tensor, edge_index: Tensor, edge_attr=None, relabel_nodes: bool = False,
             num_nodes: Optional[int] = None):
    """App. A.8: keep edges with both ends in `subset` (order preserved); relabel to the
    position inside `subset`."""
    n = maybe_num_nodes(edge_index, num_nodes)
    if subset.dtype == torch.bool:
        node_mask = subset
        subset = node_mask.nonzero().view(-1)
    else:
        node_mask = torch.zeros(n, dtype=torch.bool, device=edge_index.device)
        node_mask[subset] = True
    edge_mask = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    edge_index = edge_index[:, edge_mask]
    if relabel_nodes:
        remap = torch.full((n,), -1, dtype=torch.long, device=edge_index.device)
        remap[subset] = torch.arange(subset.numel(), device=edge_index.device)
        edge_index = remap[edge_index]
    return edge_index, None


def _sample(population: int, k: int) -> Tensor:
    if population <= k:
        return torch.arange(population)
    return torch.tensor(random.sample(range(population), k))


def negative_sampling(edge_index: Tensor, num_nodes: Optional[int] = None,
                      num_neg_samples: Optional[int] = None, method: str = 'sparse',
                      force_undirected: bool = False) -> Tensor:
    """App. A.5 ('sparse' method, directed).  Uses Python `random` exactly as upstream."""
    if force_undirected or method != 'sparse':
        raise NotImplementedError("only the branch the reference uses is restated")
    n = maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    row, col = row[keep].clone(), col[keep].clone()
    col[row < col] -= 1
    idx = row * (n - 1) + col
    population = n * n - n
    if idx.numel() >= population:
        return edge_index.new_empty((2, 0))
    if num_neg_samples is None:
        num_neg_samples = edge_index.size(1)
    prob = 1.0 - idx.numel() / population
    sample_size = int(1.1 * num_neg_samples / prob)
    neg_idx = None
    idx_np = idx.cpu().numpy()
    for _ in range(3):
        rnd = _sample(population, sample_size)
        bad = np.isin(rnd.numpy(), idx_np)
        if neg_idx is not None:
            bad |= np.isin(rnd.numpy(), neg_idx.cpu().numpy())
        rnd = rnd[~torch.from_numpy(bad).to(torch.bool)].to(edge_index.device)
        neg_idx = rnd if neg_idx is None else torch.cat([neg_idx, rnd])
        if neg_idx.numel() >= num_neg_samples:
            neg_idx = neg_idx[:num_neg_samples]
            break
    r = neg_idx.div(n - 1, rounding_mode='floor')
    c = neg_idx % (n - 1)
    c[r <= c] += 1
    return torch.stack([r, c], dim=0)


def batched_negative_sampling(edge_index: Tensor, batch: Tensor,
                              num_neg_samples: Optional[int] = None, method: str = 'sparse',
                              force_undirected: bool = False) -> Tensor:
    """App. A.5: per-graph negative_sampling with the SAME num_neg_samples for every graph."""
    split = degree(batch[edge_index[0]], dtype=torch.long).tolist()
    pieces = torch.split(edge_index, split, dim=1)
    num_src = degree(batch, dtype=torch.long)
    ptr = cumsum(num_src)[:-1]
    sizes = num_src.tolist()
    out = []
    for i, ei in enumerate(pieces):
        ei = ei - ptr[i]
        neg = negative_sampling(ei, sizes[i], num_neg_samples, method, force_undirected)
        neg = neg + ptr[i]
        out.append(neg)
    return torch.cat(out, dim=1)


def to_networkx(*args, **kwargs):  # import-time stub (src/data/graph_properties.py:11)
    raise NotImplementedError("to_networkx is outside the hot path")
