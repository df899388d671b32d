"""Two-view graph augmentation (node drop / edge drop / attribute mask) — the host-side step
*before* the hot path (reference src/pretrain/augmentations.py:17-111; SURVEY.md §8f "next" #2).
All random draws come from the caller's CPU ``torch.Generator`` in the reference's order
(App. A.8) so the masks are bit-identical; only the resulting tensors reach the kernels."""
from typing import List, Tuple

import torch
from torch import Tensor

from .data import Batch, Data
from .utils import subgraph

ATTR_MASK_MIN_NUM_FEATURES = 3
ATTR_MASK_PROB = 0.2
ATTR_MASK_RATE = 0.2
EDGE_DROP_MIN_NUM_EDGES = 3
EDGE_DROP_PROB = 0.2
EDGE_DROP_RATE = 0.2
NODE_DROP_MIN_NUM_NODES = 3
NODE_DROP_RATE = 0.2


def _augmented_view(graph: Data, gen: torch.Generator) -> Tuple[Data, Tensor]:
    view = graph.clone()
    dev = view.x.device
    n = view.num_nodes
    # node drop (augmentations.py:45-60)
    if n >= NODE_DROP_MIN_NUM_NODES:
        keep = n - max(1, int(n * NODE_DROP_RATE))
        kept = torch.randperm(n, generator=gen)[:keep].to(dev).sort()[0]
        view.edge_index, _ = subgraph(kept, view.edge_index, relabel_nodes=True, num_nodes=n)
        view.x = view.x[kept]
    else:
        kept = torch.arange(n, device=dev)
    # edge drop with probability 0.2 (augmentations.py:30-42,68-69)
    if torch.rand(1, generator=gen).item() < EDGE_DROP_PROB:
        e = view.num_edges
        if e >= EDGE_DROP_MIN_NUM_EDGES:
            keep_e = e - max(1, int(e * EDGE_DROP_RATE))
            cols = torch.randperm(e, generator=gen)[:keep_e].to(dev)
            view.edge_index = view.edge_index[:, cols]
    # attribute mask with probability 0.2 (augmentations.py:17-27,71-72)
    if torch.rand(1, generator=gen).item() < ATTR_MASK_PROB:
        f = view.num_node_features
        if f >= ATTR_MASK_MIN_NUM_FEATURES:
            k = max(1, int(f * ATTR_MASK_RATE))
            cols = torch.randperm(f, generator=gen)[:k].to(dev)
            view.x[:, cols] = 0.0
    return view, kept


def _common_masks(kept_a: Tensor, kept_b: Tensor) -> Tuple[Tensor, Tensor]:
    """augmentations.py:77-85: nodes present in both views."""
    return torch.isin(kept_a, kept_b), torch.isin(kept_b, kept_a)


class GraphAugmentor:
    """reference src/pretrain/augmentations.py:88-111."""

    @staticmethod
    def create_two_views(batch: Batch, generator: torch.Generator) -> Tuple[Batch, Batch, List[Tensor], List[Tensor]]:
        first, second, masks_a, masks_b = [], [], [], []
        for graph in batch.to_data_list():
            va, ka = _augmented_view(graph, generator)
            vb, kb = _augmented_view(graph, generator)
            ma, mb = _common_masks(ka, kb)
            first.append(va)
            second.append(vb)
            masks_a.append(ma)
            masks_b.append(mb)
        return Batch.from_data_list(first), Batch.from_data_list(second), masks_a, masks_b
