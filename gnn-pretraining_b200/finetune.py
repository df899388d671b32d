"""The fine-tuning callers of the hot path (SURVEY.md §8f "next" #4): the link-prediction hard-negative miner
(reference src/finetune/finetune.py:45-106), the per-batch forward/loss dispatch `process_batch` (:136-206) and the
training iteration around it (:304-325).  Same signatures and results; metrics (sklearn) and wandb I/O stay outside.

The miner, B200-first.  The reference materialises, per training step, the [N, N] cosine-similarity matrix, an [N, N]
bool edge mask, its complement, the boolean-compressed score vector and the two int64 coordinate vectors of
`torch.where` (for a Cora-sized graph: 29 + 7 + 7 + 29 + 117 MB, five passes), and only then takes a top-k.  Here the
similarity is one tcgen05 GEMM over row-normalised embeddings (fp32-class 3xTF32, so the ranking is the fp32 ranking),
the forbidden cells (existing edges in both directions, the diagonal) are overwritten with -inf IN the similarity
matrix through their flat indices (O(E), not O(N^2)), and the top-k runs on the flat matrix; (row, col) are decoded from
the flat index.  Selecting the k largest of the admissible cells is the same set either way; only the order among
exactly tied scores (every unordered pair appears twice: sim[i, j] == sim[j, i]) is implementation-defined — it already
differs between the reference on CPU and the reference on CUDA.  The rarely taken random-fill branch (:83-101, graphs with
fewer admissible pairs than requested negatives) keeps the reference's `torch.where` row-major order and its
`torch.randperm(n, device=device)` draw, so on one device both consume the same generator stream.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import ops

# src/finetune/finetune.py:24-42
BATCH_SIZES = {'ENZYMES': 32, 'PTC_MR': 32, 'Cora_NC': -1, 'CiteSeer_NC': -1, 'Cora_LP': 256, 'CiteSeer_LP': 256}
EPOCHS = {'ENZYMES': 100, 'PTC_MR': 100, 'Cora_NC': 200, 'CiteSeer_NC': 200, 'Cora_LP': 300, 'CiteSeer_LP': 300}
HARD_NEGATIVE_RATIO = 0.3
MIN_HARD_NEGATIVES = 8
PATIENCE_FRACTION = 0.5
# src/data/data_setup.py:43-59
NUM_CLASSES = {'ENZYMES': 6, 'PTC_MR': 2, 'Cora_NC': 7, 'CiteSeer_NC': 6, 'Cora_LP': 2, 'CiteSeer_LP': 2}

_NEG_INF = float('-inf')


def cosine_similarity_matrix(node_embeddings: Tensor) -> Tensor:
    """[N, N] = normalize(h) normalize(h)^T (finetune.py:50-51): `normalize_rows` kernel + one tcgen05 GEMM in the
    error-compensated 3xTF32 mode (fp32-class; FFMA when the layout is not TMA-legal).  Returns a fresh tensor the
    caller may overwrite."""
    zn, _ = ops.normalize_rows.fn(node_embeddings.detach())
    return ops._gemm_raw(zn, False, zn, True, None, False, ops.PRECISIONS['tf32x3'])


class LinkPredictionHardNegativeMiner:
    """`mine_hard_negatives_for_edges(node_embeddings, positive_edges, num_negatives, existing_edges) -> [2, K]`
    (finetune.py:45-106).  `similarity` replaces the similarity kernel (tests pin the selection logic on CPU with the
    reference's own expression)."""

    def __init__(self, similarity: Optional[Callable[[Tensor], Tensor]] = None):
        self._similarity = similarity or cosine_similarity_matrix

    def mine_hard_negatives_for_edges(self, node_embeddings: Tensor, positive_edges: Tensor, num_negatives: int,
                                      existing_edges: Tensor) -> Tensor:
        device = node_embeddings.device
        n = node_embeddings.size(0)
        sim = self._similarity(node_embeddings)
        flat = sim.view(-1)
        forbidden = torch.arange(n, device=device) * (n + 1)                          # the diagonal (:59)
        if existing_edges.size(1) > 0:                                                # both directions (:55-57)
            u, v = existing_edges[0], existing_edges[1]
            forbidden = torch.cat([u * n + v, v * n + u, forbidden])
        flat.index_fill_(0, forbidden, _NEG_INF)
        num_potential = n * n - int(torch.unique(forbidden).numel())                  # = potential_negatives_mask.sum()
        empty = torch.empty(2, 0, dtype=torch.long, device=device)
        if num_potential == 0:
            return empty
        num_hard = max(MIN_HARD_NEGATIVES, int(num_potential * HARD_NEGATIVE_RATIO))
        num_hard = min(num_hard, num_potential, num_negatives)
        hard = empty
        if num_hard > 0:
            top = torch.topk(flat, num_hard, largest=True).indices
            hard_src, hard_dst = torch.div(top, n, rounding_mode='floor'), top % n
            hard = torch.stack([hard_src, hard_dst], dim=0)
        remaining = num_negatives - num_hard
        if remaining <= 0:
            return hard
        # random fill from the admissible cells not taken (in either direction) as hard negatives (:83-101)
        if num_hard > 0:
            flat.index_fill_(0, torch.cat([hard_src * n + hard_dst, hard_dst * n + hard_src]), _NEG_INF)
        available = torch.nonzero(sim != _NEG_INF)                                    # row-major, like torch.where
        if available.size(0) == 0:
            return hard
        take = min(remaining, available.size(0))
        pick = torch.randperm(available.size(0), device=device)[:take]
        rand = available[pick].t()
        return torch.cat([hard, rand], dim=1) if num_hard > 0 else rand


def _classification_outputs(logits: Tensor, targets: Tensor, domain_name: str) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """finetune.py:150-160 / :171-181: binary domains train the positive-class logit with BCE-with-logits."""
    if NUM_CLASSES[domain_name] == 2:
        loss = F.binary_cross_entropy_with_logits(logits[:, 1], targets.float())
    elif ops.on_device(logits) and logits.dim() == 2 and logits.size(0) > 0:
        loss = ops.cross_entropy_sum(logits, targets)[0] / logits.size(0)      # F.cross_entropy (mean) in one launch
    else:
        loss = F.cross_entropy(logits, targets)
    return loss, targets, torch.argmax(logits, dim=1), F.softmax(logits, dim=1)


def process_batch(model: torch.nn.Module, batch, device: torch.device, task_type: str, domain_name: str,
                  hard_negative_miner: Optional[LinkPredictionHardNegativeMiner],
                  train_edges_for_hard_mining: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(loss, targets, predictions, probabilities) of one batch for the three downstream task types
    (finetune.py:136-206)."""
    if task_type == 'graph_classification':
        batch = batch.to(device)
        return _classification_outputs(model(batch), batch.y, domain_name)
    if task_type == 'node_classification':
        data, node_indices, targets = batch
        data, node_indices, targets = data.to(device), node_indices.to(device), targets.to(device)
        logits = model(data, message_passing_edges=train_edges_for_hard_mining)[node_indices]
        return _classification_outputs(logits, targets, domain_name)
    if task_type == 'link_prediction':
        if model.training:
            data, pos_edges, _ = batch
            data, pos_edges = data.to(device), pos_edges.to(device)
            with torch.no_grad():       # the miner's embeddings come from a train-mode pass (BN stats move, dropout on)
                emb = model.gnn_backbone(model.input_encoder(data.x), train_edges_for_hard_mining)
            neg_edges = hard_negative_miner.mine_hard_negatives_for_edges(
                node_embeddings=emb, positive_edges=pos_edges, num_negatives=pos_edges.size(1),
                existing_edges=train_edges_for_hard_mining).to(device)
            all_edges = torch.cat([pos_edges, neg_edges], dim=1)
            edge_labels = torch.cat([torch.ones(pos_edges.size(1), device=device),
                                     torch.zeros(neg_edges.size(1), device=device)])
        else:
            data, all_edges, edge_labels = batch
            data, all_edges, edge_labels = data.to(device), all_edges.to(device), edge_labels.to(device)
        probs = model(data, edge_index=all_edges, message_passing_edges=train_edges_for_hard_mining)
        loss = F.binary_cross_entropy(probs, edge_labels)
        return loss, edge_labels.long(), (probs > 0.5).long(), torch.stack([1 - probs, probs], dim=1)
    raise ValueError(f'unknown task type {task_type!r}')


def train_step(model: torch.nn.Module, optimizer: torch.optim.Optimizer, batch, device: torch.device, task_type: str,
               domain_name: str, hard_negative_miner: Optional[LinkPredictionHardNegativeMiner] = None,
               train_edges: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """One iteration of the fine-tuning loop (finetune.py:304-325 without the metrics): returns what
    `compute_training_metrics` consumes."""
    if train_edges is not None:
        train_edges = train_edges.to(device)
    loss, targets, predictions, probabilities = process_batch(model, batch, device, task_type, domain_name,
                                                              hard_negative_miner, train_edges)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss, targets, predictions, probabilities
