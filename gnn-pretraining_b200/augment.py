"""Two-view graph augmentation (node drop / edge drop / attribute mask) — the step *before* the hot
path (reference src/pretrain/augmentations.py:17-111; SURVEY.md §8f "next" #2).

The reference clones every graph of the batch and runs ~15 tiny tensor ops per view on the device
(launch-bound: 64 views of a 32-graph batch are ~1,000 launches).  Here the *plan* is made on the host
from a CPU copy of the edge list — every random draw comes from the caller's CPU ``torch.Generator``
in the reference's exact order (App. A.8: per graph, view 1 then view 2: randperm(n), rand(1),
[randperm(E')], rand(1), [randperm(F)]) — and is then *applied* to the whole batch with a handful of
batched device ops (one row gather, one index_put for masked attributes, one upload of the relabelled
edge list).  Kept-node sets, edge order, masked columns and common-node masks are bit-identical to the
reference's (tests/test_augment_parity.py).
"""
from typing import List, Tuple

import numpy as np
import torch
from torch import Tensor

from .data import Batch

ATTR_MASK_MIN_NUM_FEATURES = 3
ATTR_MASK_PROB = 0.2
ATTR_MASK_RATE = 0.2
EDGE_DROP_MIN_NUM_EDGES = 3
EDGE_DROP_PROB = 0.2
EDGE_DROP_RATE = 0.2
NODE_DROP_MIN_NUM_NODES = 3
NODE_DROP_RATE = 0.2


class _ViewPlan:
    """Host-side description of one augmented view of the whole batch."""

    def __init__(self):
        self.total = 0                            # nodes of the view batch so far
        self.node_ids: List[np.ndarray] = []      # global ids of kept nodes, per graph (ascending)
        self.edges: List[np.ndarray] = []         # [2, e'] relabelled + offset into the view batch, per graph
        self.sizes: List[int] = []
        self.mask_rows: List[np.ndarray] = []     # flat indices into x_view.view(-1) to zero
        self.kept_local: List[np.ndarray] = []

    def add(self, start: int, kept: np.ndarray, edges: np.ndarray, feat_cols, num_feats: int):
        offset = self.total
        self.total += int(kept.size)
        self.node_ids.append(kept + start)
        self.edges.append(edges + offset)
        if feat_cols is not None and kept.size:
            rows = np.arange(offset, offset + kept.size, dtype=np.int64)
            self.mask_rows.append((rows[:, None] * num_feats + feat_cols[None, :]).reshape(-1))
        self.sizes.append(int(kept.size))
        self.kept_local.append(kept)

    def build(self, x: Tensor) -> Batch:
        dev = x.device
        node_ids = np.concatenate(self.node_ids) if self.node_ids else np.zeros(0, dtype=np.int64)
        edges = np.concatenate(self.edges, axis=1) if self.edges else np.zeros((2, 0), dtype=np.int64)
        sizes = np.asarray(self.sizes, dtype=np.int64)
        ptr = np.zeros(len(self.sizes) + 1, dtype=np.int64)
        np.cumsum(sizes, out=ptr[1:])
        xv = x.index_select(0, torch.from_numpy(node_ids).to(dev))
        if self.mask_rows:
            flat = torch.from_numpy(np.concatenate(self.mask_rows)).to(dev)
            xv.view(-1).index_fill_(0, flat, 0.0)
        batch_vec = torch.from_numpy(np.repeat(np.arange(len(self.sizes), dtype=np.int64), sizes)).to(dev)
        out = Batch.from_tensors(xv, torch.from_numpy(np.ascontiguousarray(edges)).to(dev), batch_vec,
                                 torch.from_numpy(ptr).to(dev))
        out._ptr_host = ptr.tolist()
        return out


def _plan_view(n: int, edges_local: np.ndarray, num_feats: int, gen: torch.Generator):
    """One augmented view of one graph (reference augmentations.py:63-74).  Returns (kept nodes ascending,
    relabelled edges [2, e'], masked feature columns or None)."""
    # node drop (augmentations.py:45-60)
    if n >= NODE_DROP_MIN_NUM_NODES:
        keep = n - max(1, int(n * NODE_DROP_RATE))
        kept = np.sort(torch.randperm(n, generator=gen)[:keep].numpy())
        new_id = np.full(n, -1, dtype=np.int64)
        new_id[kept] = np.arange(kept.size, dtype=np.int64)
        src, dst = edges_local[0], edges_local[1]
        alive = (new_id[src] >= 0) & (new_id[dst] >= 0)
        edges = np.stack([new_id[src[alive]], new_id[dst[alive]]])
    else:
        kept = np.arange(n, dtype=np.int64)
        edges = edges_local
    # edge drop with probability 0.2 (augmentations.py:30-42,68-69); survivors follow the permutation's order
    if torch.rand(1, generator=gen).item() < EDGE_DROP_PROB:
        e = edges.shape[1]
        if e >= EDGE_DROP_MIN_NUM_EDGES:
            keep_e = e - max(1, int(e * EDGE_DROP_RATE))
            cols = torch.randperm(e, generator=gen)[:keep_e].numpy()
            edges = edges[:, cols]
    # attribute mask with probability 0.2 (augmentations.py:17-27,71-72)
    feat_cols = None
    if torch.rand(1, generator=gen).item() < ATTR_MASK_PROB:
        if num_feats >= ATTR_MASK_MIN_NUM_FEATURES:
            k = max(1, int(num_feats * ATTR_MASK_RATE))
            feat_cols = torch.randperm(num_feats, generator=gen)[:k].numpy().astype(np.int64)
    return kept, edges, feat_cols


class GraphAugmentor:
    """reference src/pretrain/augmentations.py:88-111 — same signature and return values."""

    @staticmethod
    def create_two_views(batch, generator: torch.Generator) -> Tuple[Batch, Batch, List[Tensor], List[Tensor]]:
        x = batch.x
        dev = x.device
        num_feats = 1 if x.dim() == 1 else x.size(-1)
        ptr = getattr(batch, '_ptr_host', None)
        if ptr is None:
            ptr = batch.ptr.tolist()
        ei = getattr(batch, '_edge_index_host', None)       # host mirror kept by gnnb200.loader: no device read-back
        if ei is None:
            ei = batch.edge_index.cpu().numpy()             # one transfer; every per-graph decision is host arithmetic
        # edges of graph g are the columns whose source lies in [ptr[g], ptr[g+1]) (PyG batches keep them grouped)
        graph_of_edge = np.searchsorted(np.asarray(ptr[1:]), ei[0], side='right') if ei.shape[1] else np.zeros(0, dtype=np.int64)
        order_ok = ei.shape[1] == 0 or bool(np.all(np.diff(graph_of_edge) >= 0))
        if not order_ok:
            raise ValueError('edge_index columns must be grouped by graph (Batch.from_data_list order)')
        cuts = np.searchsorted(graph_of_edge, np.arange(len(ptr)), side='left')
        plans = (_ViewPlan(), _ViewPlan())
        masks_a, masks_b = [], []
        for g in range(len(ptr) - 1):
            start, n = ptr[g], ptr[g + 1] - ptr[g]
            local = ei[:, cuts[g]:cuts[g + 1]] - start
            views = []
            for plan in plans:
                kept, edges, feat_cols = _plan_view(n, local, num_feats, generator)
                plan.add(start, kept, edges, feat_cols, num_feats)
                views.append(kept)
            # augmentations.py:77-85: nodes present in both views (membership through a per-graph bitmap: the kept
            # sets are subsets of range(n), so this is isin() without its sort/concatenate machinery)
            in_a, in_b = np.zeros(n, dtype=bool), np.zeros(n, dtype=bool)
            in_a[views[0]] = True
            in_b[views[1]] = True
            masks_a.append(in_b[views[0]])
            masks_b.append(in_a[views[1]])
        v1, v2 = plans[0].build(x), plans[1].build(x)
        # all masks travel in one upload and are handed back as per-graph views of it
        def to_device(masks):
            if not masks:
                return []
            flat = torch.from_numpy(np.concatenate(masks)).to(dev)
            return list(torch.split(flat, [m.size for m in masks]))
        dev_a, dev_b = to_device(masks_a), to_device(masks_b)
        # rows of each view that survive in both (what NodeContrastiveTask gathers, reference tasks.py:153-164), computed
        # here from the host masks: the device masks would cost one nonzero() = one device sync per graph and view
        for view, host_masks, dev_masks in ((v1, masks_a, dev_a), (v2, masks_b, dev_b)):
            starts = view._ptr_host
            rows = [np.flatnonzero(m) + starts[g] for g, m in enumerate(host_masks)]
            view._common_rows_host = (dev_masks, np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64))
        return v1, v2, dev_a, dev_b
