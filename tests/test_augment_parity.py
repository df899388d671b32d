"""Host-planned / batch-applied two-view augmentation (gnnb200/augment.py) against the oracle's restatement
of src/pretrain/augmentations.py (itself pinned to the unmodified reference in test_oracle_reference.py):
same CPU generator -> bit-identical kept nodes, relabelled edge lists (including the permuted order after an
edge drop), masked attribute columns and common-node masks, and the same number of draws consumed."""
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import augment, data as pdata, synthetic
from helpers import oracle_batch
from oracle import modules as orc


def _product_batch(graphs):
    return pdata.Batch.from_data_list([pdata.Data(**{k: v.clone() for k, v in g.items()}) for g in graphs])


@pytest.mark.parametrize('domain,count,seed', [('ENZYMES', 32, 1), ('MUTAG', 16, 2), ('PROTEINS', 8, 3), ('NCI1', 8, 4)])
def test_two_views_bit_identical(domain, count, seed):
    graphs = synthetic.tu_like_graphs(domain, count, seed=seed)
    for gen_seed in (0, 7, 123):
        ga, gb = torch.Generator().manual_seed(gen_seed), torch.Generator().manual_seed(gen_seed)
        o1, o2, om1, om2 = orc.GraphAugmentor.create_two_views(oracle_batch(graphs), ga)
        p1, p2, pm1, pm2 = augment.GraphAugmentor.create_two_views(_product_batch(graphs), gb)
        for o, p in ((o1, p1), (o2, p2)):
            assert torch.equal(o.x, p.x)
            assert torch.equal(o.edge_index, p.edge_index)
            assert torch.equal(o.batch, p.batch)
            assert torch.equal(o.ptr, p.ptr)
            assert o.num_graphs == p.num_graphs
        for a, b in zip(om1 + om2, pm1 + pm2):
            assert torch.equal(a, b)
        # both sides consumed the same number of draws: the next draw agrees
        assert torch.equal(torch.rand(3, generator=ga), torch.rand(3, generator=gb))


def test_tiny_and_edgeless_graphs():
    g = torch.Generator().manual_seed(0)
    graphs = [
        {'x': torch.randn(2, 5, generator=g), 'edge_index': torch.tensor([[0, 1], [1, 0]])},     # < 3 nodes: no node drop
        {'x': torch.randn(6, 5, generator=g), 'edge_index': torch.empty(2, 0, dtype=torch.long)},  # no edges
        {'x': torch.randn(4, 5, generator=g), 'edge_index': torch.tensor([[0, 1, 2], [1, 2, 3]])},
        {'x': torch.randn(1, 5, generator=g), 'edge_index': torch.empty(2, 0, dtype=torch.long)},
    ]
    for gen_seed in range(20):
        ga, gb = torch.Generator().manual_seed(gen_seed), torch.Generator().manual_seed(gen_seed)
        o1, o2, om1, om2 = orc.GraphAugmentor.create_two_views(oracle_batch(graphs), ga)
        p1, p2, pm1, pm2 = augment.GraphAugmentor.create_two_views(_product_batch(graphs), gb)
        assert torch.equal(o1.x, p1.x) and torch.equal(o2.x, p2.x)
        assert torch.equal(o1.edge_index, p1.edge_index) and torch.equal(o2.edge_index, p2.edge_index)
        assert all(torch.equal(a, b) for a, b in zip(om1 + om2, pm1 + pm2))


def test_single_large_graph_cora_shape():
    d = synthetic.cora_like(5)
    ga, gb = torch.Generator().manual_seed(9), torch.Generator().manual_seed(9)
    o1, o2, om1, om2 = orc.GraphAugmentor.create_two_views(oracle_batch([d]), ga)
    p1, p2, pm1, pm2 = augment.GraphAugmentor.create_two_views(_product_batch([d]), gb)
    assert torch.equal(o1.x, p1.x) and torch.equal(o1.edge_index, p1.edge_index)
    assert torch.equal(o2.x, p2.x) and torch.equal(o2.edge_index, p2.edge_index)
    assert torch.equal(om1[0], pm1[0]) and torch.equal(om2[0], pm2[0])


def test_common_rows_planned_on_the_host_equal_the_mask_path():
    """NodeContrastiveTask._common_rows: the host-planned row indices attached to the views equal what the per-graph
    nonzero() over the device masks gives (reference tasks.py:153-164), without one device sync per graph."""
    from gnnb200.tasks import NodeContrastiveTask
    graphs = synthetic.tu_like_graphs('ENZYMES', 12, seed=6)
    for seed in (0, 3):
        v1, v2, m1, m2 = augment.GraphAugmentor.create_two_views(_product_batch(graphs), torch.Generator().manual_seed(seed))
        for view, masks in ((v1, m1), (v2, m2)):
            planned = NodeContrastiveTask._common_rows(view, masks)
            assert view._common_rows_host[0] is masks
            via_masks = NodeContrastiveTask._common_rows(view, list(masks))        # a different list object: mask path
            assert planned.dtype == via_masks.dtype == torch.long and torch.equal(planned, via_masks)
        # both views keep the same nodes in common, graph by graph
        assert NodeContrastiveTask._common_rows(v1, m1).numel() == NodeContrastiveTask._common_rows(v2, m2).numel()
