"""ORACLE (test infrastructure, never shipped): torch_geometric.nn primitives restated.

Upstream PyG is absent (see utils/__init__.py header) -> parity vs real PyG UNPINNED.
Anchoring call sites in the reference:
  GINConv            src/models/gnn.py:4,29-37,41
  global_mean_pool   src/models/finetune_model.py:7,75 ; src/pretrain/tasks.py:9,241,245,299,331
  global_max_pool    src/pretrain/tasks.py:9,242,246
"""
from typing import Optional

import torch
from torch import Tensor

from torch_geometric.utils import scatter


class SumAggregation(torch.nn.Module):
    """Parameter-free child visible in reference checkpoints as `gin_conv.aggr_module`."""

    def forward(self, x: Tensor, index: Tensor, dim_size: int) -> Tensor:
        return scatter(x, index, dim=0, dim_size=dim_size, reduce='sum')


def _reset(module: torch.nn.Module) -> None:
    if hasattr(module, 'reset_parameters'):
        module.reset_parameters()
        return
    for child in module.children():
        _reset(child)


class GINConv(torch.nn.Module):
    """App. A.1: out = nn( sum_{j->i} x_j + (1+eps) * x_i ), flow source_to_target."""

    def __init__(self, nn: torch.nn.Module, eps: float = 0.0, train_eps: bool = False, **kwargs):
        super().__init__()
        self.aggr_module = SumAggregation()
        self.nn = nn
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.empty(1))
        else:
            self.register_buffer('eps', torch.empty(1))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        _reset(self.nn)
        self.eps.data.fill_(self.initial_eps)

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        msg = x.index_select(0, edge_index[0])
        out = self.aggr_module(msg, edge_index[1], dim_size=x.size(0))
        out = out + (1 + self.eps) * x
        return self.nn(out)


def global_mean_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    if batch is None:
        return x.mean(dim=-2, keepdim=x.dim() <= 2)
    return scatter(x, batch, dim=-2, dim_size=size, reduce='mean')


def global_max_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    if batch is None:
        return x.max(dim=-2, keepdim=x.dim() <= 2)[0]
    return scatter(x, batch, dim=-2, dim_size=size, reduce='max')


def global_add_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    if batch is None:
        return x.sum(dim=-2, keepdim=x.dim() <= 2)
    return scatter(x, batch, dim=-2, dim_size=size, reduce='sum')
