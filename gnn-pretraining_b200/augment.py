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

from .data import Batch, host_mirror

ATTR_MASK_MIN_NUM_FEATURES = 3
ATTR_MASK_PROB = 0.2
ATTR_MASK_RATE = 0.2
EDGE_DROP_MIN_NUM_EDGES = 3
EDGE_DROP_PROB = 0.2
EDGE_DROP_RATE = 0.2
NODE_DROP_MIN_NUM_NODES = 3
NODE_DROP_RATE = 0.2


class _ViewDraws:
    """What the random stream decided for one view of the whole batch (the draws are per graph and sequential, because they
    must consume the reference's CPU generator in its order; everything derived from them is assembled batch-wide)."""

    def __init__(self, total_nodes: int):
        self.alive = np.zeros(total_nodes, dtype=bool)      # node kept by the node drop
        self.edge_picks = {}                                # graph -> positions (permuted) inside its surviving edge block
        self.feat_cols = {}                                 # graph -> masked feature columns


def _draw_view(view: _ViewDraws, g: int, start: int, n: int, edge_lo: int, edge_hi: int, src: np.ndarray, dst: np.ndarray,
               num_feats: int, gen: torch.Generator) -> None:
    """The draws of one augmented view of one graph, in the reference's order (augmentations.py:63-74; App. A.8):
    randperm(n) -> rand(1) -> [randperm(E')] -> rand(1) -> [randperm(F)]."""
    # node drop (augmentations.py:45-60)
    if n >= NODE_DROP_MIN_NUM_NODES:
        keep = n - max(1, int(n * NODE_DROP_RATE))
        view.alive[torch.randperm(n, generator=gen)[:keep].numpy() + start] = True
    else:
        view.alive[start:start + n] = True
    # edge drop with probability 0.2 (augmentations.py:30-42,68-69); survivors follow the permutation's order
    if torch.rand(1, generator=gen).item() < EDGE_DROP_PROB:
        e = edge_hi - edge_lo
        if e and n >= NODE_DROP_MIN_NUM_NODES:              # edges that survived this graph's node drop
            e = int(np.count_nonzero(view.alive[src[edge_lo:edge_hi]] & view.alive[dst[edge_lo:edge_hi]]))
        if e >= EDGE_DROP_MIN_NUM_EDGES:
            keep_e = e - max(1, int(e * EDGE_DROP_RATE))
            view.edge_picks[g] = torch.randperm(e, generator=gen)[:keep_e].numpy()
    # attribute mask with probability 0.2 (augmentations.py:17-27,71-72)
    if torch.rand(1, generator=gen).item() < ATTR_MASK_PROB:
        if num_feats >= ATTR_MASK_MIN_NUM_FEATURES:
            k = max(1, int(num_feats * ATTR_MASK_RATE))
            view.feat_cols[g] = torch.randperm(num_feats, generator=gen)[:k].numpy().astype(np.int64)


def _assemble_view(view: _ViewDraws, x: Tensor, src: np.ndarray, dst: np.ndarray, graph_of_node: np.ndarray,
                   graph_of_edge: np.ndarray, num_graphs: int, num_feats: int):
    """Batch-wide assembly of one view: kept nodes (ascending = grouped by graph), relabelled edges in original order
    (permuted selections spliced in for the graphs that drew an edge drop), masked attribute cells.  Returns the view
    Batch and its kept global node ids."""
    dev = x.device
    node_ids = np.flatnonzero(view.alive)
    sizes = np.bincount(graph_of_node[node_ids], minlength=num_graphs).astype(np.int64)
    ptr = np.zeros(num_graphs + 1, dtype=np.int64)
    np.cumsum(sizes, out=ptr[1:])
    new_id = np.cumsum(view.alive, dtype=np.int64) - 1                   # id in the view batch (offsets included)
    survive = view.alive[src] & view.alive[dst]
    edges = np.stack([new_id[src[survive]], new_id[dst[survive]]])       # original column order
    if view.edge_picks:
        per_graph = np.bincount(graph_of_edge[survive], minlength=num_graphs)
        block = np.zeros(num_graphs + 1, dtype=np.int64)
        np.cumsum(per_graph, out=block[1:])
        pieces, at = [], 0
        for g in sorted(view.edge_picks):
            if block[g] > at:
                pieces.append(np.arange(at, block[g], dtype=np.int64))
            pieces.append(view.edge_picks[g] + block[g])
            at = block[g + 1]
        if edges.shape[1] > at:
            pieces.append(np.arange(at, edges.shape[1], dtype=np.int64))
        edges = edges[:, np.concatenate(pieces)] if pieces else edges[:, :0]
    xv = x.index_select(0, torch.from_numpy(node_ids).to(dev))
    cells = [(np.arange(ptr[g], ptr[g + 1], dtype=np.int64)[:, None] * num_feats + cols[None, :]).reshape(-1)
             for g, cols in view.feat_cols.items() if sizes[g]]
    if cells:
        xv.view(-1).index_fill_(0, torch.from_numpy(np.concatenate(cells)).to(dev), 0.0)
    batch_vec = torch.from_numpy(np.repeat(np.arange(num_graphs, dtype=np.int64), sizes)).to(dev)
    out = Batch.from_tensors(xv, torch.from_numpy(np.ascontiguousarray(edges)).to(dev), batch_vec, torch.from_numpy(ptr).to(dev))
    out._ptr_host = ptr.tolist()
    return out, node_ids, sizes


class GraphAugmentor:
    """reference src/pretrain/augmentations.py:88-111 — same signature and return values."""

    @staticmethod
    def create_two_views(batch, generator: torch.Generator) -> Tuple[Batch, Batch, List[Tensor], List[Tensor]]:
        x = batch.x
        dev = x.device
        num_feats = 1 if x.dim() == 1 else x.size(-1)
        ptr = host_mirror(batch, '_ptr_host')
        if ptr is None:
            ptr = batch.ptr.tolist()
        ei = host_mirror(batch, '_edge_index_host')       # host mirror kept by gnnb200.loader: no device read-back
        if ei is None:
            ei = batch.edge_index.cpu().numpy()             # one transfer; every per-graph decision is host arithmetic
        # edges of graph g are the columns whose source lies in [ptr[g], ptr[g+1]) (PyG batches keep them grouped)
        graph_of_edge = np.searchsorted(np.asarray(ptr[1:]), ei[0], side='right') if ei.shape[1] else np.zeros(0, dtype=np.int64)
        order_ok = ei.shape[1] == 0 or bool(np.all(np.diff(graph_of_edge) >= 0))
        if not order_ok:
            raise ValueError('edge_index columns must be grouped by graph (Batch.from_data_list order)')
        num_graphs = len(ptr) - 1
        cuts = np.searchsorted(graph_of_edge, np.arange(len(ptr)), side='left').tolist()
        total = ptr[-1]
        src, dst = (ei[0], ei[1]) if ei.shape[1] else (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))
        views = (_ViewDraws(total), _ViewDraws(total))
        for g in range(num_graphs):                          # the only per-graph work: the draws themselves
            start, n = ptr[g], ptr[g + 1] - ptr[g]
            for view in views:
                _draw_view(view, g, start, n, cuts[g], cuts[g + 1], src, dst, num_feats, generator)
        graph_of_node = np.repeat(np.arange(num_graphs, dtype=np.int64), np.diff(np.asarray(ptr, dtype=np.int64)))
        (v1, ids_a, sizes_a), (v2, ids_b, sizes_b) = (
            _assemble_view(view, x, src, dst, graph_of_node, graph_of_edge, num_graphs, num_feats) for view in views)
        # augmentations.py:77-85: nodes present in both views, per graph (one flat membership lookup per view)
        flat_a, flat_b = views[1].alive[ids_a], views[0].alive[ids_b]
        dev_a = list(torch.split(torch.from_numpy(flat_a).to(dev), sizes_a.tolist())) if num_graphs else []
        dev_b = list(torch.split(torch.from_numpy(flat_b).to(dev), sizes_b.tolist())) if num_graphs else []
        # rows of each view that survive in both (what NodeContrastiveTask gathers, reference tasks.py:153-164), from the
        # host masks: the device masks would cost one nonzero() = one device sync per graph and view
        v1._common_rows_host = (dev_a, np.flatnonzero(flat_a))
        v2._common_rows_host = (dev_b, np.flatnonzero(flat_b))
        return v1, v2, dev_a, dev_b
