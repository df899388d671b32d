"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference's model
modules, heads, pre-training tasks, augmentations and schedulers on top of the pure-PyTorch
torch_geometric shim.  Class names, constructor arguments, forward signatures and
state-dict keys are the reference's; every class cites the file:line it follows
(paths relative to /root/reference).  Pinned bit-for-bit against the unmodified reference in
tests/test_oracle_reference.py and via tests/golden/.

`hidden_dim` / `num_layers` are constructor parameters with the reference's constants as
defaults (BASELINE config 1 asks for L=3, the reference hard-codes 5).
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from oracle import install_pyg_shim

install_pyg_shim()
from torch_geometric.data import Batch, Data  # noqa: E402
from torch_geometric.nn import GINConv, global_max_pool, global_mean_pool  # noqa: E402
from torch_geometric.utils import batched_negative_sampling, subgraph, to_undirected  # noqa: E402

# ---- constants -----------------------------------------------------------------------------
# src/models/gnn.py:6-8
DROPOUT_RATE = 0.2
GNN_HIDDEN_DIM = 256
GNN_NUM_LAYERS = 5
# src/models/heads.py:10-13
CONTRASTIVE_PROJ_DIM = 128
DOMAIN_CLASSIFIER_DROPOUT_RATE = 0.5
DOMAIN_CLASSIFIER_HIDDEN_DIM = 128
GRAPH_PROP_HIDDEN_DIM = 512
# src/data/graph_properties.py:13
GRAPH_PROPERTY_DIM = 12
# src/models/pretrain_model.py:18-20
MASK_TOKEN_INIT_STD = 0.1
NODE_FEATURE_MASKING_MASK_RATE = 0.15
NODE_FEATURE_MASKING_MIN_NUM_NODES = 3
# src/models/finetune_model.py:14-17
FINETUNE_HIDDEN_DIM = 128
LR_BACKBONE = 1e-4
LR_FINETUNE = 1e-3
# src/data/data_setup.py:24-59
PRETRAIN_TUDATASETS = ['MUTAG', 'PROTEINS', 'NCI1', 'ENZYMES']
DOMAIN_DIMENSIONS = {'MUTAG': 7, 'PROTEINS': 4, 'NCI1': 37, 'ENZYMES': 21, 'PTC_MR': 18,
                     'Cora_NC': 1433, 'CiteSeer_NC': 3703, 'Cora_LP': 1433, 'CiteSeer_LP': 3703}
NUM_CLASSES = {'ENZYMES': 6, 'PTC_MR': 2, 'Cora_NC': 7, 'CiteSeer_NC': 6, 'Cora_LP': 2, 'CiteSeer_LP': 2}
TASK_TYPES = {'ENZYMES': 'graph_classification', 'PTC_MR': 'graph_classification',
              'Cora_NC': 'node_classification', 'CiteSeer_NC': 'node_classification',
              'Cora_LP': 'link_prediction', 'CiteSeer_LP': 'link_prediction'}
# src/pretrain/augmentations.py:7-14
ATTR_MASK_MIN_NUM_FEATURES = 3
ATTR_MASK_PROB = 0.2
ATTR_MASK_RATE = 0.2
EDGE_DROP_MIN_NUM_EDGES = 3
EDGE_DROP_PROB = 0.2
EDGE_DROP_RATE = 0.2
NODE_DROP_MIN_NUM_NODES = 3
NODE_DROP_RATE = 0.2
# src/pretrain/schedulers.py:3-7
FINAL_TEMP = 0.2
GAMMA = 10.0
INITIAL_TEMP = 0.5
MAX_LAMBDA = 0.01
START_ADVERSARIAL_EPOCH_FRACTION = 0.4
# src/pretrain/pretrain.py:43-52
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


# ---- backbone (src/models/gnn.py) ---------------------------------------------------------------
class InputEncoder(nn.Module):
    """src/models/gnn.py:11-23 — Linear -> BatchNorm1d -> ReLU -> Dropout(0.2)."""

    def __init__(self, dim_in: int, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.linear = nn.Linear(dim_in, hidden_dim)
        self.batch_norm = nn.BatchNorm1d(hidden_dim)
        self.dropout = nn.Dropout(DROPOUT_RATE)

    def forward(self, x: Tensor) -> Tensor:
        return self.dropout(F.relu(self.batch_norm(self.linear(x))))


class GINLayer(nn.Module):
    """src/models/gnn.py:26-43 — GINConv(MLP H->2H->BN->ReLU->H, train_eps) + residual -> BN -> ReLU -> dropout."""

    def __init__(self, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        mlp = nn.Sequential(nn.Linear(hidden_dim, 2 * hidden_dim), nn.BatchNorm1d(2 * hidden_dim),
                            nn.ReLU(), nn.Linear(2 * hidden_dim, hidden_dim))
        self.gin_conv = GINConv(mlp, train_eps=True)
        self.batch_norm = nn.BatchNorm1d(hidden_dim)

    def forward(self, h: Tensor, edge_index: Tensor) -> Tensor:
        z = self.gin_conv(h, edge_index) + h
        z = F.relu(self.batch_norm(z))
        return F.dropout(z, p=DROPOUT_RATE, training=self.training)


class GINBackbone(nn.Module):
    """src/models/gnn.py:46-54."""

    def __init__(self, num_layers: int = GNN_NUM_LAYERS, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.layers = nn.ModuleList([GINLayer(hidden_dim) for _ in range(num_layers)])

    def forward(self, h: Tensor, edge_index: Tensor) -> Tensor:
        for layer in self.layers:
            h = layer(h, edge_index)
        return h


# ---- heads (src/models/heads.py) -------------------------------------------------------------
class GradientReversalFunction(torch.autograd.Function):
    """src/models/heads.py:16-24 — identity forward, -lambda * g backward."""

    @staticmethod
    def forward(ctx, x, lambda_val):
        ctx.lambda_val = lambda_val
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.neg() * ctx.lambda_val, None


class GradientReversalLayer(nn.Module):
    """src/models/heads.py:27-32."""

    def forward(self, x, lambda_val):
        return GradientReversalFunction.apply(x, lambda_val)


class MLPHead(nn.Module):
    """src/models/heads.py:35-50 — Linear(+ReLU+Dropout) stack, last layer linear."""

    def __init__(self, dims: List[int], dropout_rates: List[float] = None) -> None:
        super().__init__()
        seq = []
        last = len(dims) - 2
        for i in range(len(dims) - 1):
            seq.append(nn.Linear(dims[i], dims[i + 1]))
            if i < last:
                seq.append(nn.ReLU())
                seq.append(nn.Dropout(DROPOUT_RATE if dropout_rates is None else dropout_rates[i]))
        self.mlp = nn.Sequential(*seq)

    def forward(self, x: Tensor) -> Tensor:
        return self.mlp(x)


class MLPLinkPredictor(nn.Module):
    """src/models/heads.py:53-67 — sigmoid(MLP([hu+hv, hu*hv, abs(hu-hv)]))."""

    def __init__(self, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.predictor = MLPHead([3 * hidden_dim, hidden_dim, 1])

    def forward(self, h: Tensor, edge_index: Tensor) -> Tensor:
        hu, hv = h[edge_index[0]], h[edge_index[1]]
        feats = torch.cat([hu + hv, hu * hv, torch.abs(hu - hv)], dim=1)
        return torch.sigmoid(self.predictor(feats).squeeze(-1))


class DomainClassifierHead(nn.Module):
    """src/models/heads.py:70-82 — GRL -> MLP[H,128,4] (always 4 outputs, App. C.7)."""

    def __init__(self, hidden_dim: int = GNN_HIDDEN_DIM):
        super().__init__()
        self.grl = GradientReversalLayer()
        self.classifier = MLPHead([hidden_dim, DOMAIN_CLASSIFIER_HIDDEN_DIM, len(PRETRAIN_TUDATASETS)],
                                  dropout_rates=[DOMAIN_CLASSIFIER_DROPOUT_RATE])

    def forward(self, x: Tensor, lambda_val: float) -> Tensor:
        return self.classifier(self.grl(x, lambda_val))


# ---- models (src/models/pretrain_model.py, finetune_model.py) ----------------------------------
class PretrainableGNN(nn.Module):
    """src/models/pretrain_model.py:23-99."""

    def __init__(self, device: torch.device, domain_names: List[str], task_names: List[str],
                 num_layers: int = GNN_NUM_LAYERS, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.device = device
        H = hidden_dim
        self.input_encoders = nn.ModuleDict({d: InputEncoder(DOMAIN_DIMENSIONS[d], H) for d in domain_names})
        self.mask_token = nn.Parameter(torch.zeros(H))
        nn.init.normal_(self.mask_token, std=MASK_TOKEN_INIT_STD)
        self.gnn_backbone = GINBackbone(num_layers, H)
        per_domain = {
            'node_feat_mask': [H, H, H],
            'node_contrast': [H, H, CONTRASTIVE_PROJ_DIM],
            'graph_contrast': [2 * H, H, CONTRASTIVE_PROJ_DIM],
            'graph_prop': [H, GRAPH_PROP_HIDDEN_DIM, GRAPH_PROPERTY_DIM],
        }
        self.heads = nn.ModuleDict()
        for t in task_names:
            if t in per_domain:
                self.heads[t] = nn.ModuleDict({d: MLPHead(per_domain[t]) for d in domain_names})
            elif t == 'link_pred':
                self.heads[t] = MLPLinkPredictor(H)
            elif t == 'domain_adv':
                self.heads[t] = DomainClassifierHead(H)
        self.to(self.device)

    def apply_node_masking(self, batch: Batch, domain_name: str, generator: torch.Generator
                           ) -> Tuple[Tensor, Tensor, Tensor]:
        """pretrain_model.py:67-88 — encoder under no_grad (train-mode BN/dropout still active),
        per-graph randperm from the CPU generator, 15 % of nodes (>=1) for graphs with >=3 nodes."""
        with torch.no_grad():
            h0 = self.input_encoders[domain_name](batch.x)
        picked = []
        for g in range(batch.num_graphs):
            lo, hi = batch.ptr[g].item(), batch.ptr[g + 1].item()
            n = hi - lo
            if n >= NODE_FEATURE_MASKING_MIN_NUM_NODES:
                k = max(1, int(n * NODE_FEATURE_MASKING_MASK_RATE))
                picked.append(torch.randperm(n, generator=generator)[:k].to(self.device) + lo)
        if not picked:
            return (h0, torch.empty(0, dtype=torch.long, device=self.device),
                    torch.empty(0, h0.size(1), device=self.device))
        idx = torch.cat(picked)
        masked = h0.clone()
        masked[idx] = self.mask_token.expand(len(idx), -1)
        return masked, idx, h0[idx].detach()

    def forward(self, batch: Batch, domain_name: str) -> Tensor:
        return self.gnn_backbone(self.input_encoders[domain_name](batch.x), batch.edge_index)

    def forward_with_h0(self, h_0: Tensor, edge_index: Tensor) -> Tensor:
        return self.gnn_backbone(h_0, edge_index)

    def get_head(self, task_name: str, domain_name: Optional[str] = None) -> nn.Module:
        head = self.heads[task_name]
        return head if domain_name is None else head[domain_name]


class FinetuneGNN(nn.Module):
    """src/models/finetune_model.py:20-80 (the wandb artifact loader at :83-153 is out of scope)."""

    def __init__(self, device: torch.device, domain_name: str, finetune_strategy: str,
                 num_layers: int = GNN_NUM_LAYERS, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.device = device
        self.domain_name = domain_name
        H = hidden_dim
        self.input_encoder = InputEncoder(DOMAIN_DIMENSIONS[domain_name], H)
        self.gnn_backbone = GINBackbone(num_layers, H)
        kind = TASK_TYPES[domain_name]
        if kind == 'graph_classification':
            self.classification_head = MLPHead([H, FINETUNE_HIDDEN_DIM, NUM_CLASSES[domain_name]])
        elif kind == 'node_classification':
            self.classification_head = MLPHead([H, NUM_CLASSES[domain_name]])
        else:
            self.classification_head = MLPLinkPredictor(H)

        self.param_groups = []
        if domain_name == 'ENZYMES':                      # App. C.11: encoder frozen even for b1
            for p in self.input_encoder.parameters():
                p.requires_grad = False
        else:
            self.param_groups.append({'params': self.input_encoder.parameters(), 'lr': LR_FINETUNE, 'name': 'encoder'})
        if finetune_strategy == 'linear_probe':
            for p in self.gnn_backbone.parameters():
                p.requires_grad = False
        else:
            self.param_groups.append({'params': self.gnn_backbone.parameters(), 'lr': LR_BACKBONE, 'name': 'backbone'})
        self.param_groups.append({'params': self.classification_head.parameters(), 'lr': LR_FINETUNE, 'name': 'head'})
        self.to(self.device)

    def forward(self, batch: Batch, edge_index: Optional[Tensor] = None,
                message_passing_edges: Optional[Tensor] = None) -> Tensor:
        h0 = self.input_encoder(batch.x)
        mp = batch.edge_index if message_passing_edges is None else message_passing_edges
        h = self.gnn_backbone(h0, mp)
        kind = TASK_TYPES[self.domain_name]
        if kind == 'graph_classification':
            return self.classification_head(global_mean_pool(h, batch.batch))
        if kind == 'node_classification':
            return self.classification_head(h)
        return self.classification_head(h, edge_index)


# ---- schedulers (src/pretrain/schedulers.py) ----------------------------------------------------
class TemperatureScheduler:
    """schedulers.py:10-21 — geometric 0.5 -> 0.2."""

    def __init__(self, total_steps: int):
        self.total_steps = total_steps
        self.current_step = 0

    def __call__(self) -> float:
        frac = min(1.0, self.current_step / self.total_steps)
        return float(INITIAL_TEMP * (FINAL_TEMP / INITIAL_TEMP) ** frac)

    def step(self):
        self.current_step += 1


class GRLScheduler:
    """schedulers.py:24-45 — lambda = 0 before 40 % of training, then 0.01*(2/(1+exp(-10p))-1)."""

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


# ---- augmentations (src/pretrain/augmentations.py) -------------------------------------------
def _attribute_mask(data: Data, generator: torch.Generator) -> Data:
    """augmentations.py:17-27."""
    f = data.num_node_features
    if f < ATTR_MASK_MIN_NUM_FEATURES:
        return data
    k = max(1, int(f * ATTR_MASK_RATE))
    cols = torch.randperm(f, generator=generator)[:k].to(data.x.device)
    data.x[:, cols] = 0.0
    return data


def _edge_drop(data: Data, generator: torch.Generator) -> Data:
    """augmentations.py:30-42."""
    e = data.num_edges
    if e < EDGE_DROP_MIN_NUM_EDGES:
        return data
    keep_n = e - max(1, int(e * EDGE_DROP_RATE))
    keep = torch.randperm(e, generator=generator)[:keep_n].to(data.edge_index.device)
    data.edge_index = data.edge_index[:, keep]
    return data


def _node_drop(data: Data, generator: torch.Generator) -> Tuple[Data, Tensor]:
    """augmentations.py:45-60."""
    n = data.num_nodes
    if n < NODE_DROP_MIN_NUM_NODES:
        return data, torch.arange(n, device=data.x.device)
    keep_n = n - max(1, int(n * NODE_DROP_RATE))
    kept = torch.randperm(n, generator=generator)[:keep_n].to(data.x.device).sort()[0]
    ei, _ = subgraph(kept, data.edge_index, relabel_nodes=True, num_nodes=n)
    data.x = data.x[kept]
    data.edge_index = ei
    return data, kept


def _create_augmented_view(data: Data, generator: torch.Generator) -> Tuple[Data, Tensor]:
    """augmentations.py:63-74 — draw order: randperm(n), rand(1), [randperm(E')], rand(1), [randperm(F)]."""
    view, kept = _node_drop(data.clone(), generator)
    if torch.rand(1, generator=generator).item() < EDGE_DROP_PROB:
        view = _edge_drop(view, generator)
    if torch.rand(1, generator=generator).item() < ATTR_MASK_PROB:
        view = _attribute_mask(view, generator)
    return view, kept


def _find_common_nodes_for_contrastive_loss(kept_1: Tensor, kept_2: Tensor) -> Tuple[Tensor, Tensor]:
    """augmentations.py:77-85."""
    uniq, cnt = torch.cat((kept_1, kept_2)).unique(return_counts=True)
    common = uniq[cnt == 2]
    return torch.isin(kept_1, common), torch.isin(kept_2, common)


class GraphAugmentor:
    """augmentations.py:88-111."""

    @staticmethod
    def create_two_views(batch: Batch, generator: torch.Generator):
        v1, v2, m1, m2 = [], [], [], []
        for g in batch.to_data_list():
            a, ka = _create_augmented_view(g, generator)
            b, kb = _create_augmented_view(g, generator)
            ma, mb = _find_common_nodes_for_contrastive_loss(ka, kb)
            v1.append(a)
            v2.append(b)
            m1.append(ma)
            m2.append(mb)
        return Batch.from_data_list(v1), Batch.from_data_list(v2), m1, m2


# ---- pre-training tasks (src/pretrain/tasks.py) ------------------------------------------------
def nt_xent(z1: Tensor, z2: Tensor, temperature: float) -> Tuple[Tensor, Tensor]:
    """tasks.py:192-213 and :265-287 (the two copies are arithmetically identical):
    L2-normalise, sim = z z^T / T, diagonal -> -inf, CE(sum) with targets i <-> i+M."""
    m = z1.size(0)
    z = torch.cat([F.normalize(z1, dim=1), F.normalize(z2, dim=1)], dim=0)
    sim = (z @ z.T) / temperature
    sim = sim.masked_fill(torch.eye(2 * m, device=z.device, dtype=torch.bool), float('-inf'))
    target = torch.cat([torch.arange(m, 2 * m, device=z.device), torch.arange(0, m, device=z.device)])
    loss = F.cross_entropy(sim, target, reduction='sum')
    return loss, torch.tensor(2 * m, device=z.device, dtype=torch.long)


class BasePretrainTask:
    """tasks.py:61-67."""

    def __init__(self, model: PretrainableGNN) -> None:
        self.model = model

    def compute_loss(self, domain_batches: Dict[str, Batch], generator: torch.Generator):
        raise NotImplementedError

    def _acc(self):
        dev = self.model.device
        return torch.tensor(0.0, device=dev), torch.tensor(0, device=dev, dtype=torch.long), {}


class NodeFeatureMaskingTask(BasePretrainTask):
    """tasks.py:70-94."""

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._acc()
        for name, batch in domain_batches.items():
            masked_h0, idx, target = self.model.apply_node_masking(batch, name, generator)
            if idx.size(0) == 0:
                per_domain[name] = torch.tensor(0.0, device=dev)
                continue
            h = self.model.forward_with_h0(masked_h0, batch.edge_index)
            recon = self.model.get_head('node_feat_mask', name)(h[idx])
            loss = F.mse_loss(recon, target, reduction='sum')
            size = torch.tensor(idx.size(0) * masked_h0.size(1), device=dev, dtype=torch.long)
            total += loss
            count += size
            per_domain[name] = loss / size
        if count > 0:
            total /= count
        return total, per_domain


class LinkPredictionTask(BasePretrainTask):
    """tasks.py:97-127 — positives = batch.edge_index, negatives from batched_negative_sampling
    over the symmetrised edges with quota = E_batch per graph (App. A.5); BCE on probabilities."""

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._acc()
        decoder = self.model.get_head('link_pred')
        for name, batch in domain_batches.items():
            pos = batch.edge_index
            neg = batched_negative_sampling(edge_index=to_undirected(pos), batch=batch.batch,
                                            num_neg_samples=pos.size(1))
            edges = torch.cat([pos, neg], dim=1)
            labels = torch.cat([torch.ones(pos.size(1), device=dev, dtype=torch.float32),
                                torch.zeros(neg.size(1), device=dev, dtype=torch.float32)], dim=0)
            probs = decoder(self.model(batch, name), edges)
            loss = F.binary_cross_entropy(probs, labels, reduction='sum')
            size = torch.tensor(labels.size(0), device=dev, dtype=torch.long)
            total += loss
            count += size
            per_domain[name] = loss / size
        total /= count
        return total, per_domain


class NodeContrastiveTask(BasePretrainTask):
    """tasks.py:130-213."""

    def __init__(self, model, temperature_scheduler: TemperatureScheduler):
        super().__init__(model)
        self.temperature_scheduler = temperature_scheduler

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._acc()
        temp = self.temperature_scheduler()
        for name, batch in domain_batches.items():
            b1, b2, m1, m2 = GraphAugmentor.create_two_views(batch, generator)
            h1 = self.model(b1, name)
            h2 = self.model(b2, name)
            c1, c2 = [], []
            for g, (ma, mb) in enumerate(zip(m1, m2)):
                c1.append(h1[b1.batch == g][ma])
                c2.append(h2[b2.batch == g][mb])
            if not c1 or not c2:
                per_domain[name] = torch.tensor(0.0, device=dev)
                continue
            c1, c2 = torch.cat(c1, dim=0), torch.cat(c2, dim=0)
            if c1.size(0) < 2 or c2.size(0) < 2:
                per_domain[name] = torch.tensor(0.0, device=dev)
                continue
            proj = self.model.get_head('node_contrast', name)
            loss, size = self._simclr_nt_xent(proj(c1), proj(c2), temp)
            total += loss
            count += size
            per_domain[name] = loss / size
        if count > 0:
            total /= count
        return total, per_domain

    def _simclr_nt_xent(self, z1, z2, temperature):
        return nt_xent(z1, z2, temperature)


class GraphContrastiveTask(BasePretrainTask):
    """tasks.py:216-287 — skipped for domains with < 2 graphs (:231-234)."""

    def __init__(self, model, temperature_scheduler: TemperatureScheduler = None):
        super().__init__(model)
        self.temperature_scheduler = temperature_scheduler

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._acc()
        temp = self.temperature_scheduler()
        for name, batch in domain_batches.items():
            if len(torch.unique(batch.batch)) < 2:
                per_domain[name] = torch.tensor(0.0, device=dev)
                continue
            b1, b2, _, _ = GraphAugmentor.create_two_views(batch, generator)
            h1 = self.model(b1, name)
            h2 = self.model(b2, name)
            s1 = torch.cat([global_mean_pool(h1, b1.batch), global_max_pool(h1, b1.batch)], dim=1)
            s2 = torch.cat([global_mean_pool(h2, b2.batch), global_max_pool(h2, b2.batch)], dim=1)
            proj = self.model.get_head('graph_contrast', name)
            loss, size = self._graph_contrastive_loss(proj(s1), proj(s2), temp)
            total += loss
            count += size
            per_domain[name] = loss / size
        if count > 0:
            total /= count
        return total, per_domain

    def _graph_contrastive_loss(self, z1, z2, temperature):
        return nt_xent(z1, z2, temperature)


class GraphPropertyPredictionTask(BasePretrainTask):
    """tasks.py:290-312."""

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._acc()
        for name, batch in domain_batches.items():
            emb = global_mean_pool(self.model(batch, name), batch.batch)
            pred = self.model.get_head('graph_prop', name)(emb)
            labels = batch.graph_properties.to(torch.float32).to(dev).view(emb.size(0), GRAPH_PROPERTY_DIM)
            loss = F.mse_loss(pred, labels, reduction='sum')
            size = torch.tensor(emb.size(0) * GRAPH_PROPERTY_DIM, device=dev, dtype=torch.long)
            total += loss
            count += size
            per_domain[name] = loss / size
        total /= count
        return total, per_domain


class DomainAdversarialTask(BasePretrainTask):
    """tasks.py:315-343."""

    def __init__(self, model: PretrainableGNN, grl_scheduler: GRLScheduler = None) -> None:
        super().__init__(model)
        self.domain_to_idx = {n: i for i, n in enumerate(self.model.input_encoders.keys())}
        self.grl_scheduler = grl_scheduler

    def compute_loss(self, domain_batches, generator):
        dev = self.model.device
        total, count, per_domain = self._acc()
        lam = self.grl_scheduler() if self.grl_scheduler is not None else 0.0
        for name, batch in domain_batches.items():
            emb = global_mean_pool(self.model(batch, name), batch.batch)
            logits = self.model.get_head('domain_adv')(emb, lam)
            labels = torch.full((emb.size(0),), self.domain_to_idx[name], device=dev, dtype=torch.long)
            loss = F.cross_entropy(logits, labels, reduction='sum')
            size = torch.tensor(labels.size(0), device=dev, dtype=torch.long)
            total += loss
            count += size
            per_domain[name] = loss / size
        total /= count
        return total, per_domain


def instantiate_tasks(model, active_tasks, grl_scheduler, temperature_scheduler):
    """src/pretrain/pretrain.py:77-93."""
    table = {
        'node_feat_mask': lambda: NodeFeatureMaskingTask(model),
        'link_pred': lambda: LinkPredictionTask(model),
        'node_contrast': lambda: NodeContrastiveTask(model, temperature_scheduler),
        'graph_contrast': lambda: GraphContrastiveTask(model, temperature_scheduler),
        'graph_prop': lambda: GraphPropertyPredictionTask(model),
        'domain_adv': lambda: DomainAdversarialTask(model, grl_scheduler),
    }
    return {name: table[name]() for name in active_tasks if name in table}


# ---- gradient surgery (src/pretrain/gradient_surgery.py) ---------------------------------------------
class GradientSurgery:
    """gradient_surgery.py:7-103 — the reference's PCGrad variant: one backward per task, per-parameter-tensor
    projection against the ORIGINAL gradients of the tasks earlier in a random.shuffle order, mean written only for
    the parameters of the first shuffled task."""

    def __init__(self, device: torch.device):
        self.device = device

    def apply_gradient_surgery(self, model, task_losses, task_names):
        import random
        if len(task_losses) <= 1:
            return {}
        per_task = {}
        for name, loss in task_losses.items():
            model.zero_grad(set_to_none=True)
            loss.backward(retain_graph=True)
            per_task[name] = {k: p.grad.clone().to(self.device) for k, p in model.named_parameters() if p.grad is not None}
        order = list(task_names)
        random.shuffle(order)
        modified = {}
        conflicts = projections = 0
        for i, ti in enumerate(order):
            modified[ti] = per_task[ti].copy()
            for tj in order[:i]:
                for key in modified[ti].keys():
                    if key not in per_task[tj]:
                        continue
                    gi, gj = modified[ti][key].flatten(), per_task[tj][key].flatten()
                    if gi.norm() == 0 or gj.norm() == 0:
                        continue
                    projections += 1
                    d = torch.dot(gi, gj)
                    if d < 0:
                        conflicts += 1
                        modified[ti][key] = (gi - (d / (gj.norm() ** 2)) * gj).reshape(modified[ti][key].shape)
        final = {}
        for key in per_task[order[0]].keys():
            stack = [modified[t][key] for t in order if key in modified[t]]
            if stack:
                final[key] = torch.stack(stack).mean(dim=0)
        for key, p in model.named_parameters():
            if key in final:
                p.grad = final[key].to(p.device)
        return {'gradient_surgery/total_conflicts': conflicts, 'gradient_surgery/total_projections': projections,
                'gradient_surgery/conflict_ratio': conflicts / max(projections, 1)}
