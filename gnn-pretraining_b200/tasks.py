"""Multi-task pre-training heads/losses with the reference's call surface
(`BasePretrainTask.compute_loss(domain_batches, generator) -> (loss, per_domain)`,
reference src/pretrain/tasks.py:61-343; called from src/pretrain/pretrain.py:124-129,220).
Backbone passes, pooling, decoder features and the NT-Xent loss run on the gnnb200 kernels;
loss normalisation (sum / integer count, divided once) follows the reference."""
import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .augment import GraphAugmentor
from .data import host_mirror
from .models import GRAPH_PROPERTY_DIM, PretrainableGNN
from .nn import global_max_pool, global_mean_pool
from .utils import batched_negative_sampling, to_undirected

# reference src/pretrain/schedulers.py:3-7
FINAL_TEMP = 0.2
GAMMA = 10.0
INITIAL_TEMP = 0.5
MAX_LAMBDA = 0.01
START_ADVERSARIAL_EPOCH_FRACTION = 0.4
# reference src/pretrain/pretrain.py:43-52
ACTIVE_TASKS = {
    'b2': ['node_feat_mask'],
    'b3': ['node_contrast'],
    'b4': ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop'],
    's1': ['node_feat_mask', 'link_pred'],
    's2': ['node_contrast', 'graph_contrast'],
    's3': ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast'],
    's4': ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop'],
    's5': ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop', 'domain_adv'],
}


class TemperatureScheduler:
    """reference src/pretrain/schedulers.py:10-21."""

    def __init__(self, total_steps: int):
        self.total_steps = total_steps
        self.current_step = 0

    def __call__(self) -> float:
        t = min(1.0, self.current_step / self.total_steps)
        return float(INITIAL_TEMP * (FINAL_TEMP / INITIAL_TEMP) ** t)

    def step(self):
        self.current_step += 1


class GRLScheduler:
    """reference src/pretrain/schedulers.py:24-45."""

    def __init__(self, total_epochs: int, steps_per_epoch: int):
        self.total_steps = total_epochs * steps_per_epoch
        self.start_steps = START_ADVERSARIAL_EPOCH_FRACTION * total_epochs * steps_per_epoch
        self.current_step = 0

    def __call__(self) -> float:
        if self.current_step < self.start_steps:
            return 0.0
        p = float(self.current_step - self.start_steps) / float(self.total_steps - self.start_steps)
        return float((2.0 / (1.0 + math.exp(-GAMMA * p)) - 1.0) * MAX_LAMBDA)

    def step(self):
        self.current_step += 1


# 2M at or above this takes the tensor-core path: sim = zn zn^T on tcgen05 (materialised once, overwritten by its own
# gradient) instead of the fused FFMA kernel that never materialises it.  The FFMA kernel is exact-fp32 and memory-free
# but O((2M)^2 D) scalar work; the crossover on B200 is around a few thousand rows.
NTXENT_TENSOR_CORE_ROWS = 2048
NTXENT_TENSOR_CORE_MAX_ROWS = 46000          # 8.5 GB of fp32 similarities


def nt_xent(z1: Tensor, z2: Tensor, temperature: float) -> Tuple[Tensor, Tensor]:
    """reference tasks.py:192-213 / :265-287: normalise, similarity / T with the diagonal excluded, CE(sum) against
    the paired view.  Small batches: one fused kernel pair (tiled similarity + online log-sum-exp, fp32 FFMA); large
    batches: similarity and its two backward contractions as tcgen05 GEMMs."""
    from . import nn as _gnn
    m = z1.size(0)
    z = torch.cat([z1, z2], dim=0)
    size = torch.tensor(2 * m, device=z1.device, dtype=torch.long)
    prec = _gnn.default_precision()
    if NTXENT_TENSOR_CORE_ROWS <= 2 * m <= NTXENT_TENSOR_CORE_MAX_ROWS and prec != 'f32' and z.size(1) % 4 == 0:
        loss = ops.ntxent_tensor_core(z, float(temperature), ops.PRECISIONS[prec])
        return loss.squeeze(0), size
    loss, _, _, _ = ops.ntxent_fwd(z, float(temperature))
    return loss.squeeze(0), size


def _num_graphs(batch) -> int:
    return batch.num_graphs


class BasePretrainTask:
    """reference tasks.py:61-67."""

    def __init__(self, model: PretrainableGNN) -> None:
        self.model = model

    def compute_loss(self, domain_batches: Dict[str, object], generator: torch.Generator
                     ) -> Tuple[Tensor, Dict[str, Tensor]]:
        raise NotImplementedError

    def _start(self):
        dev = self.model.device
        return torch.tensor(0.0, device=dev), 0, {}

    @staticmethod
    def _finish(total: Tensor, count: int, always_divide: bool) -> Tensor:
        if always_divide or count > 0:
            total = total / count
        return total


class NodeFeatureMaskingTask(BasePretrainTask):
    """reference tasks.py:70-94."""

    def compute_loss(self, domain_batches, generator):
        total, count, per_domain = self._start()
        for name, batch in domain_batches.items():
            masked_h0, idx, target = self.model.apply_node_masking(batch, name, generator)
            if idx.size(0) == 0:
                per_domain[name] = torch.tensor(0.0, device=self.model.device)
                continue
            h = self.model.forward_with_h0(masked_h0, batch.edge_index)
            recon = self.model.get_head('node_feat_mask', name)(ops.rows_gather(h, idx))
            loss = ops.mse_sum(recon, target)
            size = idx.size(0) * masked_h0.size(1)
            total = total + loss
            count += size
            per_domain[name] = loss / size
        return self._finish(total, count, False), per_domain


class LinkPredictionTask(BasePretrainTask):
    """reference tasks.py:97-127."""

    @staticmethod
    def _negatives(batch, pos: Tensor) -> Tensor:
        """tasks.py:107-111: batched_negative_sampling(to_undirected(edge_index), batch, E).  Batches cut by
        gnnb200.loader carry the structure on the host: symmetrise + coalesce + sampling then run there and only the
        result is uploaded (no coalesce kernels, no count read-back, no edge-list download)."""
        ei_host, ptr_host = host_mirror(batch, '_edge_index_host'), host_mirror(batch, '_ptr_host')
        if ei_host is None or ptr_host is None:
            return batched_negative_sampling(edge_index=to_undirected(pos), batch=batch.batch, num_neg_samples=pos.size(1))
        if pos.size(1) == 0:
            return pos.new_empty((2, 0))
        from .utils import batched_negative_sampling_host, to_undirected_host
        # like the device path, the node count of coalesce() is max index + 1 and graphs are read off `batch`
        und = to_undirected_host(ei_host, int(ei_host.max()) + 1)
        neg = batched_negative_sampling_host(und, np.diff(np.asarray(ptr_host, dtype=np.int64)), pos.size(1))
        if neg is None:
            return pos.new_empty((2, 0))
        return torch.from_numpy(neg).to(pos.device)

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._start()
        decoder = self.model.get_head('link_pred')
        for name, batch in domain_batches.items():
            pos = batch.edge_index
            neg = self._negatives(batch, pos)
            edges = torch.cat([pos, neg], dim=1)
            labels = torch.cat([torch.ones(pos.size(1), device=dev), torch.zeros(neg.size(1), device=dev)])
            _, loss = decoder.loss(self.model(batch, name), edges, labels)      # sigmoid + BCE(sum) fused
            size = labels.size(0)
            total = total + loss
            count += size
            per_domain[name] = loss / size
        return self._finish(total, count, True), per_domain


class NodeContrastiveTask(BasePretrainTask):
    """reference tasks.py:130-213.  The per-graph boolean-mask loop of :153-164 becomes one index
    vector per view (same rows, same order) and a single row-gather kernel."""

    def __init__(self, model, temperature_scheduler: TemperatureScheduler):
        super().__init__(model)
        self.temperature_scheduler = temperature_scheduler

    @staticmethod
    def _common_rows(view, masks) -> Tensor:
        planned = host_mirror(view, '_common_rows_host')
        if planned is not None and planned[0] is masks:      # views made by gnnb200.augment carry the answer (host side)
            return torch.from_numpy(planned[1])
        starts = host_mirror(view, '_ptr_host') or view.ptr.tolist()
        rows = [torch.nonzero(m, as_tuple=False).view(-1) + starts[g] for g, m in enumerate(masks)]
        return torch.cat(rows) if rows else torch.empty(0, dtype=torch.long, device=view.x.device)

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._start()
        temperature = self.temperature_scheduler()
        for name, batch in domain_batches.items():
            v1, v2, m1, m2 = GraphAugmentor.create_two_views(batch, generator)
            h1 = self.model(v1, name)
            h2 = self.model(v2, name)
            r1, r2 = self._common_rows(v1, m1), self._common_rows(v2, m2)
            if r1.numel() < 2 or r2.numel() < 2:
                per_domain[name] = torch.tensor(0.0, device=dev)
                continue
            proj = self.model.get_head('node_contrast', name)
            z1 = proj(ops.rows_gather(h1, r1.to(dev)))
            z2 = proj(ops.rows_gather(h2, r2.to(dev)))
            loss, size = self._simclr_nt_xent(z1, z2, temperature)
            total = total + loss
            count += int(2 * z1.size(0))
            per_domain[name] = loss / size
        return self._finish(total, count, False), per_domain

    def _simclr_nt_xent(self, z1, z2, temperature):
        return nt_xent(z1, z2, temperature)


class GraphContrastiveTask(BasePretrainTask):
    """reference tasks.py:216-287."""

    def __init__(self, model, temperature_scheduler: TemperatureScheduler = None):
        super().__init__(model)
        self.temperature_scheduler = temperature_scheduler

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._start()
        temperature = self.temperature_scheduler()
        for name, batch in domain_batches.items():
            if _num_graphs(batch) < 2:
                per_domain[name] = torch.tensor(0.0, device=dev)
                continue
            v1, v2, _, _ = GraphAugmentor.create_two_views(batch, generator)
            s = []
            for view in (v1, v2):
                h = self.model(view, name)
                b = _num_graphs(view)
                s.append(torch.cat([global_mean_pool(h, view.batch, b), global_max_pool(h, view.batch, b)], dim=1))
            proj = self.model.get_head('graph_contrast', name)
            loss, size = self._graph_contrastive_loss(proj(s[0]), proj(s[1]), temperature)
            total = total + loss
            count += int(2 * s[0].size(0))
            per_domain[name] = loss / size
        return self._finish(total, count, False), per_domain

    def _graph_contrastive_loss(self, z1, z2, temperature):
        return nt_xent(z1, z2, temperature)


class GraphPropertyPredictionTask(BasePretrainTask):
    """reference tasks.py:290-312."""

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._start()
        for name, batch in domain_batches.items():
            emb = global_mean_pool(self.model(batch, name), batch.batch, _num_graphs(batch))
            pred = self.model.get_head('graph_prop', name)(emb)
            labels = batch.graph_properties.to(torch.float32).to(dev).view(emb.size(0), GRAPH_PROPERTY_DIM)
            loss = ops.mse_sum(pred, labels)
            size = emb.size(0) * GRAPH_PROPERTY_DIM
            total = total + loss
            count += size
            per_domain[name] = loss / size
        return self._finish(total, count, True), per_domain


class DomainAdversarialTask(BasePretrainTask):
    """reference tasks.py:315-343."""

    def __init__(self, model: PretrainableGNN, grl_scheduler: GRLScheduler = None) -> None:
        super().__init__(model)
        self.domain_to_idx = {n: i for i, n in enumerate(self.model.input_encoders.keys())}
        self.grl_scheduler = grl_scheduler

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._start()
        lam = self.grl_scheduler() if self.grl_scheduler is not None else 0.0
        for name, batch in domain_batches.items():
            emb = global_mean_pool(self.model(batch, name), batch.batch, _num_graphs(batch))
            logits = self.model.get_head('domain_adv')(emb, lam)
            labels = torch.full((emb.size(0),), self.domain_to_idx[name], device=dev, dtype=torch.long)
            loss, _ = ops.cross_entropy_sum(logits, labels)
            size = labels.size(0)
            total = total + loss
            count += size
            per_domain[name] = loss / size
        return self._finish(total, count, True), per_domain


def instantiate_tasks(model, active_tasks, grl_scheduler, temperature_scheduler):
    """reference src/pretrain/pretrain.py:77-93."""
    make = {
        'node_feat_mask': lambda: NodeFeatureMaskingTask(model),
        'link_pred': lambda: LinkPredictionTask(model),
        'node_contrast': lambda: NodeContrastiveTask(model, temperature_scheduler),
        'graph_contrast': lambda: GraphContrastiveTask(model, temperature_scheduler),
        'graph_prop': lambda: GraphPropertyPredictionTask(model),
        'domain_adv': lambda: DomainAdversarialTask(model, grl_scheduler),
    }
    return {t: make[t]() for t in active_tasks if t in make}
