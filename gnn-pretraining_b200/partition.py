"""Node-partitioned execution of the backbone on one large graph over N GPUs (BASELINE config 5;
SURVEY.md §5.8b / §8e — a capability the reference does not have: it is single-device, full-graph).

Rank r owns the contiguous destination-row range [lo_r, hi_r).  Per layer:
  forward : h_full = all_gather(h_local)                  (NCCL over NVLink; every row crosses once)
            z_local = CSR_r gather over h_full + (1+eps) h_local      (rows of this rank only)
            dense transforms on local rows; BatchNorm statistics reduced over all ranks (Chan merge of the
            per-rank (n, sum, m2) in rank order -> every rank holds bit-identical statistics)
  backward: g_full = all_gather(g_local);  dh_local = CSC_r gather over g_full + (1+eps) g_local
            (the transposed partition: no reduce-scatter of float partial sums, the order stays fixed)
            BatchNorm dgamma/dbeta all-reduced before the dx pass; d(eps) and the weight gradients are
            per-rank partial sums added by one flat all-reduce before the optimizer step.
With world size 1 every collective is skipped and the result equals the single-device path.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib as L
from . import nn as gnn
from . import ops


def shard_bounds(num_nodes: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(lo, hi, rows_per_rank) of the contiguous equal split; the last rank may own fewer rows."""
    per = (num_nodes + world - 1) // world
    lo = min(num_nodes, rank * per)
    hi = min(num_nodes, lo + per)
    return lo, hi, per


class PartitionedGraph:
    """This rank's slice of the graph: CSR over its destination rows (columns = global source ids) and
    CSC over its source rows (columns = global destination ids).  Passed wherever the backbone expects
    ``edge_index``; GINConv recognises it and takes the partitioned aggregation."""

    def __init__(self, edge_index: Tensor, num_nodes: int, rank: int, world: int, group=None):
        self.num_nodes, self.rank, self.world, self.group = int(num_nodes), rank, world, group
        self.lo, self.hi, self.per = shard_bounds(self.num_nodes, rank, world)
        self.n_local = self.hi - self.lo
        src, dst = edge_index[0], edge_index[1]
        own_dst = (dst >= self.lo) & (dst < self.hi)
        fwd = torch.stack([src[own_dst], dst[own_dst] - self.lo], dim=0)
        self.rowptr, self.col, _ = ops.csr_build(fwd, max(self.n_local, 1), False)      # key = local dst
        own_src = (src >= self.lo) & (src < self.hi)
        bwd = torch.stack([src[own_src] - self.lo, dst[own_src]], dim=0)
        self.rowptr_t, self.col_t, _ = ops.csr_build(bwd, max(self.n_local, 1), True)   # key = local src
        if self.n_local == 0:
            self.rowptr, self.rowptr_t = self.rowptr[:1], self.rowptr_t[:1]
        self.local_edges = int(fwd.size(1))

    def all_gather_rows(self, x_local: Tensor) -> Tensor:
        """[N, F] from every rank's [n_local, F] shard (equal-size padded shards on the wire)."""
        if self.world == 1:
            return x_local
        f = x_local.size(1)
        if x_local.size(0) == self.per:
            send = x_local.contiguous()
        else:
            send = x_local.new_zeros(self.per, f)
            send[: x_local.size(0)] = x_local
        full = x_local.new_empty(self.per * self.world, f)
        dist.all_gather_into_tensor(full, send, group=self.group)
        return full[: self.num_nodes]


class _PartitionedGINAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h_local: Tensor, eps: Tensor, graph: PartitionedGraph) -> Tensor:
        h_local = h_local.contiguous()
        h_full = graph.all_gather_rows(h_local)
        ctx.graph = graph
        ctx.save_for_backward(h_local, eps)
        return ops._aggregate_raw(h_full, graph.rowptr, graph.col, L.AGG_SUM, h_local, eps, None)

    @staticmethod
    def backward(ctx, g_local: Tensor):
        graph = ctx.graph
        h_local, eps = ctx.saved_tensors
        g_local = g_local.contiguous()
        g_full = graph.all_gather_rows(g_local)
        gh = ops._aggregate_raw(g_full, graph.rowptr_t, graph.col_t, L.AGG_SUM, g_local, eps, None)
        geps = ops.dot(g_local, h_local) if ctx.needs_input_grad[1] else None      # per-rank partial sum
        return gh, geps, None


def partitioned_gin_aggregate(h_local: Tensor, eps: Tensor, graph: PartitionedGraph) -> Tensor:
    return _PartitionedGINAggregate.apply(h_local, eps, graph)


def merge_moments(per_rank: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Chan's parallel-variance merge of per-rank moments [P, 3, C] = (n, sum, centred m2) ->
    (total n, total sum, total m2), each [C].  Same inputs in the same rank order on every rank -> same bits."""
    cnt, sums, m2s = per_rank[:, 0], per_rank[:, 1], per_rank[:, 2]
    total_n = cnt.sum(0)
    total_sum = sums.sum(0)
    mean = total_sum / total_n.clamp(min=1.0)
    local_mean = sums / cnt.clamp(min=1.0)
    m2_tot = m2s.sum(0) + (cnt * (local_mean - mean) ** 2).sum(0)
    return total_n, total_sum, m2_tot


def gather_moments(n_rows: int, s: Tensor, m2: Tensor, group) -> Tensor:
    """All-gather this rank's (n, sum, m2) -> [P, 3, C]."""
    mine = torch.stack([torch.full_like(s, float(n_rows)), s, m2])
    world = dist.get_world_size(group)
    flat = mine.new_empty(world * 3, s.numel())            # concatenation along dim 0 (gloo and nccl agree on this form)
    dist.all_gather_into_tensor(flat, mine, group=group)
    return flat.view(world, 3, s.numel())


def synced_batch_stats(x_local: Tensor, graph_rows_total: int, group, running_mean: Optional[Tensor],
                       running_var: Optional[Tensor], momentum: float, eps: float) -> Tuple[Tensor, Tensor]:
    """Global (mean, invstd) of a row-partitioned activation: per-rank (n, sum, m2) gathered and merged with
    Chan's formula in rank order; running buffers updated from the global moments (torch's rule)."""
    s, m2 = ops.colstats(x_local.detach())
    _, gsum, m2_tot = merge_moments(gather_moments(x_local.size(0), s, m2, group))
    mean_out = torch.empty_like(s)
    invstd = torch.empty_like(s)
    L.check(ops._invoke('gnnb200_bn_finalize_f32', gsum.data_ptr(), m2_tot.data_ptr(), graph_rows_total, s.numel(),
                        eps, momentum, ops._ptr(running_mean), ops._ptr(running_var), mean_out.data_ptr(),
                        invstd.data_ptr(), ops._stream(s)), 'bn_finalize (synced)')
    return mean_out, invstd


class partition_scope:
    """Context manager: inside it BatchNormAct layers treat their input as this rank's row shard."""

    def __init__(self, graph: PartitionedGraph):
        self.graph = graph

    def __enter__(self):
        self._old = (gnn._partition, ops.SYNC_GROUP)
        gnn._partition = self.graph if self.graph.world > 1 else None
        ops.SYNC_GROUP = (self.graph.group or dist.group.WORLD) if self.graph.world > 1 else None
        return self.graph

    def __exit__(self, *exc):
        gnn._partition, ops.SYNC_GROUP = self._old


def allreduce_gradients(module: torch.nn.Module, group=None) -> None:
    """One flat all-reduce (sum) over every parameter gradient (C1 of SURVEY §2.4)."""
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


class PartitionedBackboneStep:
    """The bench workload's training step (InputEncoder + GINBackbone, loss = sum(h), AdamW) on this
    rank's shard.  Every rank builds the same model from the same seed."""

    def __init__(self, models, device, feat_in: int, hidden: int, layers: int, num_nodes: int, rank: int,
                 world: int, group=None, seed: int = 0, lr: float = 1e-4):
        torch.manual_seed(seed)
        self.model = torch.nn.ModuleDict({'input_encoder': models.InputEncoder(feat_in, hidden),
                                          'gnn_backbone': models.GINBackbone(layers, hidden)}).to(device)
        self.model.train()
        self.opt = torch.optim.AdamW(self.model.parameters(), lr=lr)
        self.num_nodes, self.rank, self.world, self.group = num_nodes, rank, world, group

    def step(self, x: Tensor, edge_index: Tensor) -> Tensor:
        graph = PartitionedGraph(edge_index, self.num_nodes, self.rank, self.world, self.group)
        x_local = x[graph.lo:graph.hi]
        self.opt.zero_grad(set_to_none=True)
        with partition_scope(graph):
            h = self.model['gnn_backbone'](self.model['input_encoder'](x_local), graph)
            loss = h.sum()
            loss.backward()
        if self.world > 1:
            allreduce_gradients(self.model, self.group)
        self.opt.step()
        return loss
