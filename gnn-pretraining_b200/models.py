"""Drop-in backbone, heads and model wrappers: same class names, constructor arguments, forward
signatures and state-dict keys as the reference's ``src/models/{gnn,heads,pretrain_model,
finetune_model}.py`` (file:line cited per class), with the message passing, pooling, dense
transforms, row gathers and decoder features running on the gnnb200 CUDA kernels.

``hidden_dim`` / ``num_layers`` are constructor parameters defaulting to the reference constants
(BASELINE config 1 asks for 3 layers; the reference hard-codes 5).
"""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from . import nn as _gnn
from .fused import GINLayerFn
from .data import host_mirror
from .graph import graph_of
from .nn import BatchNormAct, FusedAwayReLU, GINConv, Linear, global_mean_pool

# constants of the reference (src/models/gnn.py:6-8, heads.py:10-13, pretrain_model.py:18-20,
# finetune_model.py:14-17, src/data/data_setup.py:24-59, src/data/graph_properties.py:13)
DROPOUT_RATE = 0.2
GNN_HIDDEN_DIM = 256
GNN_NUM_LAYERS = 5
CONTRASTIVE_PROJ_DIM = 128
DOMAIN_CLASSIFIER_DROPOUT_RATE = 0.5
DOMAIN_CLASSIFIER_HIDDEN_DIM = 128
GRAPH_PROP_HIDDEN_DIM = 512
GRAPH_PROPERTY_DIM = 12
MASK_TOKEN_INIT_STD = 0.1
NODE_FEATURE_MASKING_MASK_RATE = 0.15
NODE_FEATURE_MASKING_MIN_NUM_NODES = 3
FINETUNE_HIDDEN_DIM = 128
LR_BACKBONE = 1e-4
LR_FINETUNE = 1e-3
PRETRAIN_TUDATASETS = ['MUTAG', 'PROTEINS', 'NCI1', 'ENZYMES']
DOMAIN_DIMENSIONS = {'MUTAG': 7, 'PROTEINS': 4, 'NCI1': 37, 'ENZYMES': 21, 'PTC_MR': 18,
                     'Cora_NC': 1433, 'CiteSeer_NC': 3703, 'Cora_LP': 1433, 'CiteSeer_LP': 3703}
NUM_CLASSES = {'ENZYMES': 6, 'PTC_MR': 2, 'Cora_NC': 7, 'CiteSeer_NC': 6, 'Cora_LP': 2, 'CiteSeer_LP': 2}
TASK_TYPES = {'ENZYMES': 'graph_classification', 'PTC_MR': 'graph_classification',
              'Cora_NC': 'node_classification', 'CiteSeer_NC': 'node_classification',
              'Cora_LP': 'link_prediction', 'CiteSeer_LP': 'link_prediction'}


class InputEncoder(nn.Module):
    """reference src/models/gnn.py:11-23."""

    def __init__(self, dim_in: int, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.linear = Linear(dim_in, hidden_dim)
        self.batch_norm = BatchNormAct(hidden_dim, relu=True)
        self.dropout = nn.Dropout(DROPOUT_RATE)

    def forward(self, x: Tensor) -> Tensor:
        # Linear -> [BatchNorm + ReLU + Dropout in one pass]; p follows self.dropout.p like the reference
        return self.batch_norm(self.linear(x, bias_feeds_norm=True), drop_p=self.dropout.p)


class GINLayer(nn.Module):
    """reference src/models/gnn.py:26-43 — `forward(h, edge_index)`."""

    def __init__(self, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.gin_conv = GINConv(
            nn.Sequential(Linear(hidden_dim, 2 * hidden_dim), BatchNormAct(2 * hidden_dim, relu=True), FusedAwayReLU(),
                          Linear(2 * hidden_dim, hidden_dim)),
            train_eps=True)
        self.batch_norm = BatchNormAct(hidden_dim, relu=True)

    fused = True      # one autograd node per layer (fused.GINLayerFn); False = one node per kernel (ops.py)

    def forward(self, h: Tensor, edge_index: Tensor) -> Tensor:
        # 5 kernels per layer forward: gather(+self term) -> GEMM -> BN+ReLU -> GEMM(+residual h) -> BN+ReLU+dropout
        mlp = self.gin_conv.nn
        if (self.fused and isinstance(edge_index, Tensor) and ops.on_device(h) and h.dim() == 2 and h.size(1) % 4 == 0
                and mlp[1].momentum is not None and self.batch_norm.momentum is not None):
            graph = graph_of(edge_index, h.size(0))
            p = DROPOUT_RATE if self.training else 0.0
            seed = _gnn._dropout_seed() if p > 0.0 else 0
            prec = ops.PRECISIONS[mlp[0].precision or _gnn.default_precision()]
            return GINLayerFn.apply(h, self.gin_conv.eps, mlp[0].weight, mlp[0].bias, mlp[1].weight, mlp[1].bias,
                                    mlp[3].weight, mlp[3].bias, self.batch_norm.weight, self.batch_norm.bias, graph,
                                    mlp[1], self.batch_norm, self.training, p, seed, prec)
        z = self.gin_conv.aggregate(h, edge_index)
        # (Linear(..., return_stats=True) can hand the BatchNorm statistics over from the GEMM epilogue; measured on
        # C5 it is slower — 265 vs 247 ms/step — because the K=256 GEMMs are epilogue-bound and 76k per-32-row partials
        # per column have to be merged, so the separate column-statistics pass stays.)
        z = mlp[2](mlp[1](mlp[0](z, bias_feeds_norm=True)))
        z = mlp[3](z, residual=h, bias_feeds_norm=True)
        return self.batch_norm(z, drop_p=DROPOUT_RATE)


class GINBackbone(nn.Module):
    """reference src/models/gnn.py:46-54 — one CSR build per edge_index tensor, shared by all layers."""

    def __init__(self, num_layers: int = GNN_NUM_LAYERS, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.layers = nn.ModuleList([GINLayer(hidden_dim) for _ in range(num_layers)])

    def forward(self, h: Tensor, edge_index: Tensor) -> Tensor:
        for layer in self.layers:
            h = layer(h, edge_index)
        return h


class GradientReversalLayer(nn.Module):
    """reference src/models/heads.py:16-32: identity forward, gradient times -lambda backward (one scale kernel)."""

    def forward(self, x, lambda_val):
        if not ops.on_device(x):
            raise ops.L.Gnnb200Error('gnnb200 modules run on CUDA tensors only (no CPU fallback)')
        return ops.gradient_reversal(x, lambda_val)


class MLPHead(nn.Module):
    """reference src/models/heads.py:35-50."""

    def __init__(self, dims: List[int], dropout_rates: List[float] = None) -> None:
        super().__init__()
        stack = []
        n_lin = len(dims) - 1
        for i in range(n_lin):
            stack.append(Linear(dims[i], dims[i + 1]))
            if i + 1 < n_lin:
                stack += [nn.ReLU(), nn.Dropout(dropout_rates[i] if dropout_rates is not None else DROPOUT_RATE)]
        self.mlp = nn.Sequential(*stack)

    def forward(self, x: Tensor) -> Tensor:
        # same modules / state-dict keys as the reference's nn.Sequential; executed as one fused step per hidden layer:
        # Linear (+ ReLU in the GEMM epilogue) (+ one ReLU/dropout launch in training mode) instead of three eager kernels
        mods = list(self.mlp)
        if x.dim() != 2 or not ops.on_device(x):
            return self.mlp(x)
        i = 0
        while i < len(mods):
            lin = mods[i]
            if i + 2 < len(mods) and isinstance(mods[i + 1], nn.ReLU) and isinstance(mods[i + 2], nn.Dropout):
                p = float(mods[i + 2].p) if self.training else 0.0
                prec = ops.PRECISIONS[lin.precision or _gnn.default_precision()]
                x = ops.linear_act(x, lin.weight, lin.bias, prec, p, _gnn._dropout_seed() if p > 0.0 else 0)
                i += 3
            else:
                x = lin(x)
                i += 1
        return x


class MLPLinkPredictor(nn.Module):
    """reference src/models/heads.py:53-67 — the [E, 3H] decoder input comes from one fused
    gather kernel (ops.lp_features) instead of two index_selects + add/mul/abs + cat."""

    def __init__(self, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.predictor = MLPHead([3 * hidden_dim, hidden_dim, 1])

    def logits(self, h: Tensor, edge_index: Tensor) -> Tensor:
        return self.predictor(ops.lp_features(h, edge_index)).squeeze(-1)

    def forward(self, h: Tensor, edge_index: Tensor) -> Tensor:
        return torch.sigmoid(self.logits(h, edge_index))

    def loss(self, h: Tensor, edge_index: Tensor, labels: Tensor):
        """(probs, F.binary_cross_entropy(probs, labels, reduction='sum')): sigmoid and the loss sum in one launch
        (reference src/models/heads.py:67 + src/pretrain/tasks.py:120)."""
        return ops.sigmoid_bce_sum(self.logits(h, edge_index), labels)


class DomainClassifierHead(nn.Module):
    """reference src/models/heads.py:70-82."""

    def __init__(self, hidden_dim: int = GNN_HIDDEN_DIM):
        super().__init__()
        self.grl = GradientReversalLayer()
        self.classifier = MLPHead([hidden_dim, DOMAIN_CLASSIFIER_HIDDEN_DIM, len(PRETRAIN_TUDATASETS)],
                                  dropout_rates=[DOMAIN_CLASSIFIER_DROPOUT_RATE])

    def forward(self, x: Tensor, lambda_val: float) -> Tensor:
        return self.classifier(self.grl(x, lambda_val))


_PER_DOMAIN_HEADS = {
    'node_feat_mask': lambda H: [H, H, H],
    'node_contrast': lambda H: [H, H, CONTRASTIVE_PROJ_DIM],
    'graph_contrast': lambda H: [2 * H, H, CONTRASTIVE_PROJ_DIM],
    'graph_prop': lambda H: [H, GRAPH_PROP_HIDDEN_DIM, GRAPH_PROPERTY_DIM],
}


class PretrainableGNN(nn.Module):
    """reference src/models/pretrain_model.py:23-99."""

    def __init__(self, device: torch.device, domain_names: List[str], task_names: List[str],
                 num_layers: int = GNN_NUM_LAYERS, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.device = device
        self.hidden_dim = hidden_dim
        self.input_encoders = nn.ModuleDict(
            {name: InputEncoder(DOMAIN_DIMENSIONS[name], hidden_dim) for name in domain_names})
        self.mask_token = nn.Parameter(torch.zeros(hidden_dim))
        nn.init.normal_(self.mask_token, std=MASK_TOKEN_INIT_STD)
        self.gnn_backbone = GINBackbone(num_layers, hidden_dim)
        self.heads = nn.ModuleDict()
        for task in task_names:
            if task in _PER_DOMAIN_HEADS:
                dims = _PER_DOMAIN_HEADS[task](hidden_dim)
                self.heads[task] = nn.ModuleDict({name: MLPHead(dims) for name in domain_names})
            elif task == 'link_pred':
                self.heads[task] = MLPLinkPredictor(hidden_dim)
            elif task == 'domain_adv':
                self.heads[task] = DomainClassifierHead(hidden_dim)
        self.to(self.device)

    def apply_node_masking(self, batch, domain_name: str, generator: torch.Generator
                           ) -> Tuple[Tensor, Tensor, Tensor]:
        """reference pretrain_model.py:67-88.  Same CPU-generator draws in the same order; the
        per-graph sizes come from one ptr.tolist() instead of 2 .item() syncs per graph, and the
        mask-token write / target gather are one row-scatter and one row-gather kernel."""
        with torch.no_grad():
            h0 = self.input_encoders[domain_name](batch.x)
        bounds = host_mirror(batch, '_ptr_host') or batch.ptr.tolist()      # gnnb200.loader batches carry a host mirror
        chosen = []
        for g in range(batch.num_graphs):
            lo, n = bounds[g], bounds[g + 1] - bounds[g]
            if n >= NODE_FEATURE_MASKING_MIN_NUM_NODES:
                k = max(1, int(n * NODE_FEATURE_MASKING_MASK_RATE))
                chosen.append(torch.randperm(n, generator=generator)[:k] + lo)
        if not chosen:
            return (h0, torch.empty(0, dtype=torch.long, device=self.device),
                    torch.empty(0, h0.size(1), device=self.device))
        idx = torch.cat(chosen).to(self.device)
        target = ops.rows_gather(h0, idx)
        masked = ops.rows_scatter(h0, self.mask_token, idx)
        return masked, idx, target

    def forward(self, batch, domain_name: str) -> Tensor:
        return self.gnn_backbone(self.input_encoders[domain_name](batch.x), batch.edge_index)

    def forward_with_h0(self, h_0: Tensor, edge_index: Tensor) -> Tensor:
        return self.gnn_backbone(h_0, edge_index)

    def get_head(self, task_name: str, domain_name: Optional[str] = None) -> nn.Module:
        head = self.heads[task_name]
        return head if domain_name is None else head[domain_name]


class FinetuneGNN(nn.Module):
    """reference src/models/finetune_model.py:20-80 (the wandb artifact loader, :83-153, is out of
    scope; `load_backbone_state` below covers its key-prefix matching, :128-146)."""

    def __init__(self, device: torch.device, domain_name: str, finetune_strategy: str,
                 num_layers: int = GNN_NUM_LAYERS, hidden_dim: int = GNN_HIDDEN_DIM) -> None:
        super().__init__()
        self.device = device
        self.domain_name = domain_name
        self.input_encoder = InputEncoder(DOMAIN_DIMENSIONS[domain_name], hidden_dim)
        self.gnn_backbone = GINBackbone(num_layers, hidden_dim)
        kind = TASK_TYPES[domain_name]
        if kind == 'graph_classification':
            self.classification_head = MLPHead([hidden_dim, FINETUNE_HIDDEN_DIM, NUM_CLASSES[domain_name]])
        elif kind == 'node_classification':
            self.classification_head = MLPHead([hidden_dim, NUM_CLASSES[domain_name]])
        elif kind == 'link_prediction':
            self.classification_head = MLPLinkPredictor(hidden_dim)

        self.param_groups = []
        if domain_name == 'ENZYMES':
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

    def forward(self, batch, edge_index: Optional[Tensor] = None,
                message_passing_edges: Optional[Tensor] = None) -> Tensor:
        h0 = self.input_encoder(batch.x)
        mp_edges = batch.edge_index if message_passing_edges is None else message_passing_edges
        h = self.gnn_backbone(h0, mp_edges)
        kind = TASK_TYPES[self.domain_name]
        if kind == 'graph_classification':
            size = getattr(batch, 'num_graphs', None) if hasattr(batch, 'ptr') and batch.ptr is not None else None
            return self.classification_head(global_mean_pool(h, batch.batch, size))
        if kind == 'node_classification':
            return self.classification_head(h)
        return self.classification_head(h, edge_index)

    def load_backbone_state(self, state_dict: dict) -> None:
        """Copy `gnn_backbone.*` (and, for ENZYMES, `input_encoders.ENZYMES.*`) tensors out of a
        pre-training checkpoint's model_state_dict, as reference finetune_model.py:128-146 does."""
        own = self.state_dict()
        for key, value in state_dict.items():
            if key.startswith('gnn_backbone.') and key in own:
                own[key].copy_(value)
            elif self.domain_name == 'ENZYMES' and key.startswith('input_encoders.ENZYMES.'):
                tgt = key.replace('input_encoders.ENZYMES.', 'input_encoder.')
                if tgt in own:
                    own[tgt].copy_(value)
