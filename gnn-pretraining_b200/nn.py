"""Drop-in message-passing primitives with the call signatures of the PyG names the reference
imports (`from torch_geometric.nn import GINConv, global_mean_pool, global_max_pool`,
reference src/models/gnn.py:4, src/models/finetune_model.py:7, src/pretrain/tasks.py:9), backed by
the gnnb200 CUDA kernels."""
from typing import Optional

import torch
from torch import Tensor

from . import _lib as L
from . import ops
from .graph import graph_of, segment_ptr_of

# Default arithmetic of the dense transforms: 'tf32_fwd3' = tcgen05 tensor cores (kind::tf32, fp32 accumulate) wherever
# TMA's layout rules hold, FFMA elsewhere; the forward GEMM of every Linear error-compensated (3xTF32 on pre-split weights,
# fp32-class activations), backward GEMMs plain tf32: end-to-end gradients inside the north star's 2e-2 class
# (ops.PRECISIONS).  'tf32' = plain tf32 everywhere; 'tf32x3' = compensated everywhere; 'f32' = FFMA everywhere (1e-5 class).
_default_precision = 'tf32_fwd3'
# node-partitioned execution (gnnb200.partition.partition_scope sets this to the rank's PartitionedGraph)
_partition = None


def set_default_precision(name: str) -> None:
    global _default_precision
    if name not in ops.PRECISIONS:
        raise ValueError(f'unknown precision {name!r}')
    _default_precision = name


def default_precision() -> str:
    return _default_precision


class Linear(torch.nn.Linear):
    """nn.Linear whose forward/backward GEMMs run on the gnnb200 kernels (same parameters/keys).
    ``residual`` (same shape as the output) is added in the GEMM epilogue."""

    precision: Optional[str] = None

    def forward(self, x: Tensor, residual: Optional[Tensor] = None, bias_feeds_norm: bool = False,
                return_stats: bool = False):
        """return_stats=True (2-D input): also returns (column sums, centred second moments) of the output, computed in
        the GEMM epilogue, for the BatchNormAct that follows (pass them as ``stats=``)."""
        prec = ops.PRECISIONS[self.precision or _default_precision]
        lead = x.shape[:-1]
        res = None if residual is None else residual.reshape(-1, self.out_features)
        # a 2-D input is passed as the caller's own tensor object: the re-pitched copy of a ragged-width feature matrix
        # (ops._tma_rows) is cached on that object and survives from step to step
        x2 = x if x.dim() == 2 else x.reshape(-1, x.size(-1))
        if return_stats:
            y, s, m2 = ops.linear_stats(x2, self.weight, self.bias, prec, res, bias_feeds_norm and self.training)
            return y.view(*lead, self.out_features), (s.detach(), m2.detach())
        y = ops.linear(x2, self.weight, self.bias, prec, res, bias_feeds_norm and self.training)
        return y.view(*lead, self.out_features)


def _dropout_seed() -> int:
    """Per-call Philox key drawn from torch's CPU generator: reproducible under torch.manual_seed,
    no device sync."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64))


class BatchNormAct(torch.nn.BatchNorm1d):
    """BatchNorm1d fused with the ReLU (and optionally the dropout) that follows it in the reference's
    InputEncoder / GINLayer (src/models/gnn.py:19-22,31-32,41-43).  Same parameters, buffers and
    state-dict keys as nn.BatchNorm1d; one read + one write of the activation in the forward pass and
    only the pre-BN activation saved for the backward pass."""

    def __init__(self, num_features: int, relu: bool = True, **kwargs):
        super().__init__(num_features, **kwargs)
        self.fused_relu = relu

    def forward(self, x: Tensor, drop_p: float = 0.0, stats=None) -> Tensor:
        """stats = (column sums, centred second moments) of x when a producer already computed them."""
        use_batch_stats = self.training or self.running_mean is None
        p = float(drop_p) if self.training else 0.0
        fusable = (x.dim() == 2 and x.dtype == torch.float32 and x.size(1) % 4 == 0
                   and self.affine and (self.momentum is not None or not use_batch_stats) and x.size(0) > 0)
        if not ops.on_device(x):
            raise L.Gnnb200Error('gnnb200 modules run on CUDA tensors only (no CPU fallback)')
        if not fusable:      # CUDA layouts outside the hot path (never hit by the reference's 256/512-wide layers)
            y = super().forward(x)
            y = torch.relu(y) if self.fused_relu else y
            return torch.nn.functional.dropout(y, p, self.training) if p > 0 else y
        if use_batch_stats and self.training and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        seed = _dropout_seed() if p > 0.0 else 0
        if use_batch_stats and _partition is not None:
            from . import partition as part        # x is this rank's row shard: reduce the statistics over ranks
            upd = self.training and self.track_running_stats
            seed ^= (_partition.rank * 0x9E3779B97F4A7C15) & (2 ** 62 - 1)      # independent masks per shard
            mean, invstd = part.synced_batch_stats(x, _partition.num_nodes, ops.SYNC_GROUP,
                                                   self.running_mean if upd else None, self.running_var if upd else None,
                                                   float(self.momentum if self.momentum is not None else 0.0), float(self.eps),
                                                   stats)
            return ops.bn_act(x, mean, invstd, self.weight, self.bias, self.fused_relu, p, seed, True,
                              _partition.num_nodes)
        if use_batch_stats and stats is not None:
            upd = self.training and self.track_running_stats
            mean, invstd = ops.bn_stats_finalize(stats[0], stats[1], x.size(0), self.running_mean if upd else None,
                                                 self.running_var if upd else None,
                                                 float(self.momentum if self.momentum is not None else 0.0), float(self.eps))
        elif use_batch_stats:
            upd = self.training and self.track_running_stats
            mean, invstd = ops.bn_batch_stats(x.detach(), self.running_mean if upd else None,
                                              self.running_var if upd else None,
                                              float(self.momentum if self.momentum is not None else 0.0), float(self.eps))
        else:
            mean, invstd = self.running_mean, torch.rsqrt(self.running_var + self.eps)
        return ops.bn_act(x, mean, invstd, self.weight, self.bias, self.fused_relu, p, seed, use_batch_stats)


class FusedAwayReLU(torch.nn.ReLU):
    """Placeholder keeping the reference's nn.Sequential indices (gin_conv.nn.{0,1,2,3}): the ReLU itself
    runs inside the preceding BatchNormAct, so this module is the identity."""

    def forward(self, x: Tensor) -> Tensor:
        return x


class SumAggregation(torch.nn.Module):
    """Parameter-free child kept so that reference checkpoints (which list `gin_conv.aggr_module`)
    load with strict key matching."""

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        g = graph_of(edge_index, x.size(0))
        return ops.aggregate(x, g.rowptr, g.col, L.AGG_SUM)


def _reset(module: torch.nn.Module) -> None:
    if hasattr(module, 'reset_parameters'):
        module.reset_parameters()
    else:
        for child in module.children():
            _reset(child)


class GINConv(torch.nn.Module):
    """out = nn( sum_{j->i} x_j + (1 + eps) * x_i )  — PyG GINConv(nn, eps, train_eps) semantics
    (SURVEY.md App. A.1), one fused CSR gather kernel for the sum and the self term."""

    def __init__(self, nn: torch.nn.Module, eps: float = 0.0, train_eps: bool = False, **kwargs):
        super().__init__()
        self.aggr_module = SumAggregation()
        self.nn = nn
        self.initial_eps = float(eps)
        if train_eps:
            self.eps = torch.nn.Parameter(torch.empty(1))
        else:
            self.register_buffer('eps', torch.empty(1))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        _reset(self.nn)
        self.eps.data.fill_(self.initial_eps)

    def aggregate(self, x: Tensor, edge_index) -> Tensor:
        if not isinstance(edge_index, Tensor):     # a gnnb200.partition.PartitionedGraph
            from .partition import partitioned_gin_aggregate
            return partitioned_gin_aggregate(x, self.eps, edge_index)
        g = graph_of(edge_index, x.size(0))
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or self.eps.requires_grad)
        if needs_grad:
            rowptr_t, col_t = g.rowptr_t, g.col_t
        else:
            rowptr_t = col_t = g.rowptr.new_empty(0)
        return ops.gin_aggregate(x, self.eps, g.rowptr, g.col, rowptr_t, col_t)

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        return self.nn(self.aggregate(x, edge_index))


def _pool(x: Tensor, batch: Optional[Tensor], size: Optional[int], mode: int) -> Tensor:
    if batch is None:
        ptr = torch.tensor([0, x.size(0)], dtype=torch.int32, device=x.device)
    else:
        ptr = segment_ptr_of(batch, size)
    return ops.segment_pool(x, ptr, mode)


def global_mean_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    """Per-graph mean (empty graphs -> 0).  ``batch`` must be sorted (PyG Batch guarantees it)."""
    return _pool(x, batch, size, L.POOL_MEAN)


def global_max_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    """Per-graph channel-wise max with torch's native amax gradient rule."""
    return _pool(x, batch, size, L.POOL_MAX)


def global_add_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    return _pool(x, batch, size, L.POOL_SUM)
