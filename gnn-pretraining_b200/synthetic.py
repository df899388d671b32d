"""Synthetic workloads with the shapes of BASELINE.json's five configs.

Everything is drawn from an explicit CPU ``torch.Generator`` so the same seed yields the
same graphs for the CUDA path, the oracle and the golden fixtures.  Graphs are returned as
plain tensors / lists of per-graph dicts; each side (product, oracle) collates them with its
own Batch class.
"""
from typing import Dict, List, Tuple

import torch

# shapes of the reference's domains (src/data/data_setup.py:31-41) -------------------------------
TU_SHAPES = {          # mean nodes, std nodes, feature dim
    'MUTAG': (17.9, 4.6, 7),
    'PROTEINS': (39.1, 30.0, 4),
    'NCI1': (29.9, 13.6, 37),
    'ENZYMES': (32.6, 15.0, 21),
    'PTC_MR': (14.3, 8.0, 18),
}
C1_NODES, C1_PAIRS, C1_FEATS = 2708, 5278, 1433
CITESEER_NODES, CITESEER_PAIRS, CITESEER_FEATS = 3327, 4552, 3703
C5_NODES, C5_EDGES, C5_FEATS = 2_449_029, 61_859_140, 100


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def random_undirected_edges(n: int, pairs: int, g: torch.Generator, shuffle: bool = True) -> torch.Tensor:
    """`pairs` distinct undirected pairs without self loops, symmetrised -> [2, 2*pairs] int64 in a
    random column order (the reference feeds arbitrarily ordered COO)."""
    pairs = min(pairs, n * (n - 1) // 2)
    if pairs <= 0:
        return torch.empty(2, 0, dtype=torch.long)
    if n * (n - 1) // 2 <= 4 * pairs or n <= 2048:
        iu = torch.triu_indices(n, n, offset=1)
        pick = torch.randperm(iu.size(1), generator=g)[:pairs]
        u, v = iu[0, pick], iu[1, pick]
    else:
        keys = torch.empty(0, dtype=torch.long)
        while keys.numel() < pairs:
            need = int((pairs - keys.numel()) * 1.1) + 16
            a = torch.randint(0, n, (need,), generator=g)
            b = torch.randint(0, n, (need,), generator=g)
            ok = a != b
            lo, hi = torch.minimum(a, b)[ok], torch.maximum(a, b)[ok]
            keys = torch.unique(torch.cat([keys, lo * n + hi]))
        keys = keys[torch.randperm(keys.numel(), generator=g)[:pairs]]
        u, v = keys // n, keys % n
    ei = torch.stack([torch.cat([u, v]), torch.cat([v, u])], dim=0)
    if shuffle:
        ei = ei[:, torch.randperm(ei.size(1), generator=g)]
    return ei.contiguous()


def planetoid_like(n: int, pairs: int, feats: int, seed: int = 42, density: float = 0.0127
                   ) -> Dict[str, torch.Tensor]:
    """C1 / CiteSeer-shaped single graph: row-normalised Bernoulli bag-of-words features
    (mimics NormalizeFeatures, src/data/data_setup.py:154)."""
    g = _gen(seed)
    ei = random_undirected_edges(n, pairs, g)
    x = (torch.rand(n, feats, generator=g) < density).to(torch.float32)
    x = x / x.sum(dim=1, keepdim=True).clamp(min=1.0)
    return {'x': x, 'edge_index': ei}


def cora_like(seed: int = 42) -> Dict[str, torch.Tensor]:
    return planetoid_like(C1_NODES, C1_PAIRS, C1_FEATS, seed)


def citeseer_like(seed: int = 42) -> Dict[str, torch.Tensor]:
    return planetoid_like(CITESEER_NODES, CITESEER_PAIRS, CITESEER_FEATS, seed, density=0.0086)


def tu_like_graphs(domain: str, num_graphs: int, seed: int = 42, num_classes: int = 6
                   ) -> List[Dict[str, torch.Tensor]]:
    """C2 / C4 small graphs: n ~ clamp(round(N(mu, sd)), 2, 126), round(1.9 n) undirected edges,
    features N(0,1) clipped to [-3, 3] (src/data/data_setup.py:17-18,99), 12 graph properties."""
    mu, sd, feats = TU_SHAPES[domain]
    g = _gen(seed)
    out = []
    sizes = torch.clamp(torch.round(torch.randn(num_graphs, generator=g) * sd + mu), 2, 126).long()
    for n in sizes.tolist():
        ei = random_undirected_edges(n, int(round(1.9 * n)), g)
        x = torch.randn(n, feats, generator=g).clamp_(-3.0, 3.0)
        y = torch.randint(0, num_classes, (1,), generator=g)
        props = torch.randn(12, generator=g)
        out.append({'x': x, 'edge_index': ei, 'y': y, 'graph_properties': props})
    return out


def products_like(num_nodes: int = C5_NODES, num_edges: int = C5_EDGES, feats: int = C5_FEATS,
                  seed: int = 42, locality: float = 0.0, blocks: int = 64, device='cpu', skew: float = 0.0
                  ) -> Dict[str, torch.Tensor]:
    """C5: one large directed multigraph in COO, random column order.  `locality` is the
    fraction of edges whose source is drawn from the destination's block of N/blocks nodes
    (0.0 = uniform random = worst-case gather locality; 0.9 mimics community structure).
    Duplicates/self loops are kept: GINConv treats them as ordinary edges (App. A.1).
    `skew` > 1 draws both endpoints from a power law instead (node rank r with probability ~ r^(1/skew - 1), ranks
    scattered over the id space): skew = 1.8 gives the largest hub ~2.8e-4 of all edges, ogbn-products' ratio (17 k of
    62 M) — SURVEY §8d input 5(ii), "degree-skewed".  skew = 0 (default) consumes exactly the stream it always did."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if skew > 1.0:
        spread = torch.randperm(num_nodes, generator=g, device=device)

        def draw():
            rank = (torch.rand(num_edges, generator=g, device=device, dtype=torch.float64) ** skew * num_nodes).long()
            return spread[rank.clamp_(max=num_nodes - 1)]
        dst, src = draw(), draw()
    else:
        dst = torch.randint(0, num_nodes, (num_edges,), generator=g, device=device)
        src = torch.randint(0, num_nodes, (num_edges,), generator=g, device=device)
    if locality > 0.0:
        bs = (num_nodes + blocks - 1) // blocks
        local = torch.rand(num_edges, generator=g, device=device) < locality
        off = torch.randint(0, bs, (num_edges,), generator=g, device=device)
        near = torch.clamp((dst // bs) * bs + off, max=num_nodes - 1)
        src = torch.where(local, near, src)
    x = torch.randn(num_nodes, feats, generator=g, device=device)
    return {'x': x, 'edge_index': torch.stack([src, dst], dim=0)}


def collate_tensors(graphs: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """Plain-tensor collate (App. A.6 layout): x cat, edge_index offset+cat, batch, ptr."""
    sizes = [gr['x'].size(0) for gr in graphs]
    ptr = torch.zeros(len(graphs) + 1, dtype=torch.long)
    ptr[1:] = torch.tensor(sizes).cumsum(0)
    out = {
        'x': torch.cat([gr['x'] for gr in graphs], dim=0),
        'edge_index': torch.cat([gr['edge_index'] + off for gr, off in zip(graphs, ptr[:-1].tolist())], dim=1),
        'batch': torch.repeat_interleave(torch.arange(len(graphs)), torch.tensor(sizes)),
        'ptr': ptr,
    }
    if 'y' in graphs[0]:
        out['y'] = torch.cat([gr['y'] for gr in graphs], dim=0)
    if 'graph_properties' in graphs[0]:
        out['graph_properties'] = torch.cat([gr['graph_properties'] for gr in graphs], dim=0)
    return out


def edge_count(graphs: List[Dict[str, torch.Tensor]]) -> Tuple[int, int]:
    return sum(gr['x'].size(0) for gr in graphs), sum(gr['edge_index'].size(1) for gr in graphs)
