"""gnnb200.loader (device-resident datasets, index-gather batching, the balanced multi-domain sampler) on CPU: batches
are identical tensor for tensor to `Batch.from_data_list` of the same picks (product Batch and the PyG shim's), and the
sampler reproduces the reference's own `BalancedMultiDomainSampler` stream (container only for that part)."""
import numpy as np
import pytest
import torch

from helpers import oracle_batch
from oracle.reference_loader import load_reference, reference_available

import gnnb200  # noqa: F401
from gnnb200 import loader, synthetic
from gnnb200.data import Batch, Data


def _graphs(domain, n, seed):
    return [Data(**g) for g in synthetic.tu_like_graphs(domain, n, seed=seed)]


def _same(a, b, names=('x', 'edge_index', 'batch', 'ptr', 'y', 'graph_properties')):
    for k in names:
        va, vb = getattr(a, k), getattr(b, k)
        assert va.dtype == vb.dtype and torch.equal(va, vb), k
    assert a.num_graphs == b.num_graphs


@pytest.mark.parametrize('picks', [[0], [3, 3, 1], [11, 0, 5, 7, 2, 2, 9, 10], list(range(12))])
def test_batch_of_equals_from_data_list(picks):
    graphs = _graphs('ENZYMES', 12, seed=1)
    graphs[5] = Data(x=graphs[5].x[:2], edge_index=torch.empty(2, 0, dtype=torch.long), y=graphs[5].y,
                     graph_properties=graphs[5].graph_properties)                     # a graph without edges
    res = loader.ResidentDomain(graphs)
    got = res.batch_of(picks)
    want = Batch.from_data_list([graphs[i] for i in picks])
    _same(got, want)
    # host mirrors the planners use instead of reading the structure back from the device
    assert got._ptr_host == want.ptr.tolist()
    assert np.array_equal(got._edge_index_host, want.edge_index.numpy())
    # to_data_list() round trip (the reference's augmentation path, augmentations.py:91)
    for g, orig in zip(got.to_data_list(), (graphs[i] for i in picks)):
        for k in ('x', 'edge_index', 'y', 'graph_properties'):
            assert torch.equal(getattr(g, k), getattr(orig, k)), k
    # ... and against the oracle's PyG shim
    _same(got, oracle_batch([{k: getattr(graphs[i], k) for k in graphs[i].keys()} for i in picks]))


def test_segments_helper():
    assert loader._segments(np.array([5, 0, 9]), np.array([2, 0, 3])).tolist() == [5, 6, 9, 10, 11]
    assert loader._segments(np.array([4]), np.array([0])).size == 0


def test_sampler_draws_and_sequential_batches():
    graphs = {d: _graphs(d, 20, seed=i) for i, d in enumerate(['MUTAG', 'ENZYMES'])}
    sets = {d: loader.GraphDataset(g, list(range(0, 20, 2))) for d, g in graphs.items()}
    s = loader.BalancedMultiDomainSampler(sets, torch.Generator().manual_seed(3), batch_size=8)
    assert len(s) == 10 // 4 and s.samples_per_domain == 4
    g2 = torch.Generator().manual_seed(3)
    for step in s:
        assert list(step) == ['MUTAG', 'ENZYMES']
        for d in step:
            picks = torch.randint(0, 10, (4,), generator=g2).tolist()
            _same(step[d], Batch.from_data_list([sets[d][i] for i in picks]))
    val = loader.sequential_batches(sets['MUTAG'], batch_size=4)
    assert [b.num_graphs for b in val] == [4, 4, 2]
    _same(val[2], Batch.from_data_list([sets['MUTAG'][8], sets['MUTAG'][9]]))


@pytest.mark.skipif(not reference_available(), reason='/root/reference not present')
def test_sampler_equals_reference_sampler():
    load_reference()
    import src.data.pretrain_data_loaders as ref_loaders
    from torch_geometric.data import Data as ShimData
    domains = ['MUTAG', 'PROTEINS', 'NCI1', 'ENZYMES']
    raw = {d: synthetic.tu_like_graphs(d, 30, seed=10 + i) for i, d in enumerate(domains)}
    splits = {d: torch.randperm(30, generator=torch.Generator().manual_seed(i))[:21] for i, d in enumerate(domains)}
    ref_sets = {d: ref_loaders.GraphDataset([ShimData(**g) for g in raw[d]], splits[d].numpy()) for d in domains}
    my_sets = {d: loader.GraphDataset([Data(**g) for g in raw[d]], splits[d].numpy()) for d in domains}
    a = ref_loaders.BalancedMultiDomainSampler(ref_sets, torch.Generator().manual_seed(42))
    b = loader.BalancedMultiDomainSampler(my_sets, torch.Generator().manual_seed(42))
    assert len(a) == len(b) and a.samples_per_domain == b.samples_per_domain == 8
    steps = 0
    for want, got in zip(a, b):
        assert list(want) == list(got)
        for d in domains:
            _same(got[d], want[d])
        steps += 1
    assert steps == len(a) == 2


def test_resident_batches_feed_the_host_planners_without_readback():
    """A loader batch (host mirrors attached) gives the augmentation planner the same views as the collated batch."""
    from gnnb200 import augment
    graphs = _graphs('ENZYMES', 16, seed=4)
    picks = [3, 3, 0, 15, 8, 9, 2, 11]
    mirrored = loader.ResidentDomain(graphs).batch_of(picks)
    plain = Batch.from_data_list([graphs[i] for i in picks])
    for mirror in ('_edge_index_host', '_ptr_host'):          # force the read-back path for the comparison
        plain.__dict__.pop(mirror, None)
    assert getattr(plain, '_edge_index_host', None) is None
    for seed in (0, 5):
        a = augment.GraphAugmentor.create_two_views(mirrored, torch.Generator().manual_seed(seed))
        b = augment.GraphAugmentor.create_two_views(plain, torch.Generator().manual_seed(seed))
        for va, vb in zip(a[:2], b[:2]):
            assert torch.equal(va.x, vb.x) and torch.equal(va.edge_index, vb.edge_index) and torch.equal(va.ptr, vb.ptr)
        assert all(torch.equal(x, y) for x, y in zip(a[2] + a[3], b[2] + b[3]))


@pytest.mark.skipif(not reference_available(), reason='/root/reference not present')
def test_finetune_loaders_equal_reference_loaders():
    """NodeBatches / LinkBatches against the reference's item-wise datasets + collate functions driven by a real
    torch DataLoader (src/data/finetune_data_loaders.py:14-107)."""
    load_reference()
    import src.data.finetune_data_loaders as ref
    from torch_geometric.data import Data as ShimData
    d = synthetic.planetoid_like(200, 400, 16, seed=2)
    y = torch.randint(0, 7, (200,), generator=torch.Generator().manual_seed(0))
    split = torch.randperm(200, generator=torch.Generator().manual_seed(1))[:70]
    gen = torch.Generator().manual_seed(9)
    for bs in (-1, 32):
        ref_data = ShimData(x=d['x'], edge_index=d['edge_index'], y=y)
        ds = ref.NodeDataset(ref_data, split.numpy())
        want = list(torch.utils.data.DataLoader(ds, batch_size=len(ds) if bs == -1 else bs, generator=gen,
                                                collate_fn=ref._collate_node_batch))
        got = list(loader.NodeBatches(Data(x=d['x'], edge_index=d['edge_index'], y=y), split.numpy(), bs))
        assert len(want) == len(got) == len(loader.NodeBatches(Data(x=d['x'], edge_index=d['edge_index'], y=y), split.numpy(), bs))
        for (wd, wi, wl), (gd, gi, gl) in zip(want, got):
            assert torch.equal(wd.x, gd.x) and torch.equal(wd.edge_index, gd.edge_index)
            assert wi.dtype == gi.dtype and torch.equal(wi, gi) and wl.dtype == gl.dtype and torch.equal(wl, gl)
    e = d['edge_index']
    splits = {'train_pos': e[:, :300], 'val_pos': e[:, 300:350], 'val_neg': e[:, 350:400].flip(0),
              'test_pos': e[:, 400:470], 'test_neg': e[:, 470:540].flip(0)}
    for name in ('train', 'val', 'test'):
        ref_data = ShimData(x=d['x'], edge_index=d['edge_index'])
        ds = ref.LinkPredictionDataset(ref_data, splits, name)
        want = list(torch.utils.data.DataLoader(ds, batch_size=64, generator=gen, collate_fn=ref._collate_link_batch))
        mine = loader.LinkBatches(Data(x=d['x'], edge_index=d['edge_index']), splits, name, 64)
        got = list(mine)
        assert len(want) == len(got) == len(mine) and torch.equal(mine.dataset.train_edges, ds.train_edges)
        for (wd, we, wl), (gd, ge, gl) in zip(want, got):
            assert we.dtype == ge.dtype and torch.equal(we, ge) and wl.dtype == gl.dtype and torch.equal(wl, gl)


@pytest.mark.skipif(not reference_available(), reason='/root/reference not present')
def test_on_disk_loader_factories_equal_the_reference(tmp_path, monkeypatch):
    """The reference's processed-data layout (data.pt / splits.pt / graph_properties.pt, data_setup.py:66-72) written by
    the REFERENCE's save_processed_data, then read by the reference's loader factories and by gnnb200.loader's factories
    of the same names: identical batches in identical order (pretrain_data_loaders.py:56-83, finetune_data_loaders.py:68-119)."""
    load_reference()
    import src.data.data_setup as ref_setup
    import src.data.finetune_data_loaders as ref_ft
    import src.data.pretrain_data_loaders as ref_pt
    from torch_geometric.data import Data as ShimData
    for mod in (ref_setup, ref_ft, ref_pt):
        monkeypatch.setattr(mod, 'PROCESSED_DIR', tmp_path)
    # two TU-shaped graph domains with properties, one node-classification graph, one link-prediction graph
    for i, dom in enumerate(['ENZYMES', 'PTC_MR']):
        graphs = [ShimData(x=g['x'], edge_index=g['edge_index'], y=g['y']) for g in synthetic.tu_like_graphs(dom, 40, seed=5 + i)]
        perm = torch.randperm(40, generator=torch.Generator().manual_seed(i)).numpy()
        splits = {'train': perm[:24], 'val': perm[24:32], 'test': perm[32:]}
        props = torch.randn(40, 12, generator=torch.Generator().manual_seed(20 + i))
        ref_setup.save_processed_data(dom, graphs, splits, props)
    d = synthetic.planetoid_like(150, 300, 12, seed=3)
    y = torch.randint(0, 7, (150,), generator=torch.Generator().manual_seed(0))
    perm = torch.randperm(150, generator=torch.Generator().manual_seed(1)).numpy()
    ref_setup.save_processed_data('Cora_NC', [ShimData(x=d['x'], edge_index=d['edge_index'], y=y)],
                                  {'train': perm[:60], 'val': perm[60:100], 'test': perm[100:]})
    e = d['edge_index']
    ref_setup.save_processed_data('Cora_LP', [ShimData(x=d['x'], edge_index=e[:, :300])],
                                  {'train_pos': e[:, :300], 'val_pos': e[:, 300:340], 'val_neg': e[:, 340:380].flip(0),
                                   'test_pos': e[:, 380:440], 'test_neg': e[:, 440:500].flip(0)})

    # ---- pre-training: balanced sampler over the train splits (graph_properties attached), validation loader
    want = ref_pt.create_train_data_loader(['ENZYMES', 'PTC_MR'], torch.Generator().manual_seed(7))
    got = loader.create_train_data_loader(['ENZYMES', 'PTC_MR'], torch.Generator().manual_seed(7), root=tmp_path)
    assert len(want) == len(got)
    for w, g in zip(want, got):
        assert list(w) == list(g)
        for dom in w:
            _same(g[dom], w[dom])
    want_val = list(ref_pt.create_val_data_loader('PTC_MR', torch.Generator().manual_seed(1)))
    got_val = loader.create_val_data_loader('PTC_MR', torch.Generator().manual_seed(1), root=tmp_path)
    assert len(want_val) == len(got_val)
    for w, g in zip(want_val, got_val):
        _same(g, w)
    # ---- fine-tuning: the dispatcher and the three loader kinds
    for split, bs in (('train', 16), ('test', 5)):
        w_all = list(ref_ft.create_finetune_data_loader('ENZYMES', split, bs, torch.Generator().manual_seed(2)))
        g_all = loader.create_finetune_data_loader('ENZYMES', split, bs, torch.Generator().manual_seed(2), root=tmp_path)
        assert len(w_all) == len(g_all)
        for w, g in zip(w_all, g_all):
            _same(g, w, names=('x', 'edge_index', 'batch', 'ptr', 'y'))
    for bs in (-1, 32):
        w_all = list(ref_ft.create_finetune_data_loader('Cora_NC', 'val', bs, torch.Generator().manual_seed(2)))
        g_all = list(loader.create_finetune_data_loader('Cora_NC', 'val', bs, torch.Generator().manual_seed(2), root=tmp_path))
        assert len(w_all) == len(g_all)
        for (wd, wi, wl), (gd, gi, gl) in zip(w_all, g_all):
            assert torch.equal(wd.x, gd.x) and torch.equal(wi, gi) and torch.equal(wl, gl) and wl.dtype == gl.dtype
    for split in ('train', 'val', 'test'):
        ref_loader = ref_ft.create_finetune_data_loader('Cora_LP', split, 64, torch.Generator().manual_seed(2))
        mine = loader.create_finetune_data_loader('Cora_LP', split, 64, torch.Generator().manual_seed(2), root=tmp_path)
        assert torch.equal(mine.dataset.train_edges, ref_loader.dataset.train_edges)
        for (wd, we, wl), (gd, ge, gl) in zip(list(ref_loader), list(mine)):
            assert torch.equal(wd.edge_index, gd.edge_index) and torch.equal(we, ge) and torch.equal(wl, gl)


def test_processed_data_round_trip_with_own_classes(tmp_path):
    """save_processed_data / create_* with the package's own Data class (no PyG, no reference)."""
    graphs = _graphs('ENZYMES', 10, seed=3)
    props = torch.stack([g.graph_properties.view(-1) for g in graphs])
    bare = [Data(x=g.x, edge_index=g.edge_index, y=g.y) for g in graphs]
    loader.save_processed_data('ENZYMES', bare, {'train': np.arange(6), 'val': np.arange(6, 10)}, props, root=tmp_path)
    s = loader.create_train_data_loader(['ENZYMES'], torch.Generator().manual_seed(0), root=tmp_path)
    step = s.draw()['ENZYMES']
    assert step.num_graphs == loader.BATCH_SIZE and step.graph_properties.numel() == loader.BATCH_SIZE * 12   # PyG cat layout
    val = loader.create_val_data_loader('ENZYMES', torch.Generator(), root=tmp_path)
    assert [b.num_graphs for b in val] == [4] and torch.equal(val[0].graph_properties.view(4, 12), props[6:10])
    ft = loader.create_finetune_data_loader('ENZYMES', 'val', 3, torch.Generator(), root=tmp_path)
    assert [b.num_graphs for b in ft] == [3, 1]


def test_from_data_list_keeps_host_mirrors_until_the_structure_changes():
    """Batch.from_data_list on host tensors keeps `_ptr_host` / `_edge_index_host` (what the augmentation, negative-sampling
    and masking planners read) across .to(device); assigning new structure drops them."""
    graphs = _graphs('MUTAG', 5, seed=2)
    b = Batch.from_data_list(graphs)
    assert b._ptr_host == b.ptr.tolist() and np.array_equal(b._edge_index_host, b.edge_index.numpy())
    moved = b.to('cpu')
    assert moved._ptr_host == b.ptr.tolist()
    c = b.clone()
    assert np.array_equal(c._edge_index_host, b.edge_index.numpy())
    b.x = b.x * 2                                            # features may change freely
    assert hasattr(b, '_edge_index_host')
    b.edge_index = b.edge_index[:, :3]
    assert '_edge_index_host' not in b.__dict__ and '_ptr_host' not in b.__dict__


def test_host_mirrors_are_dropped_after_an_in_place_edit_of_the_structure():
    """The planners (augmentation, node masking, negative sampling) read the structure from host mirrors; an in-place edit of
    edge_index / ptr / batch must invalidate them (the kernels see the new tensor: gnnb200.graph.graph_of checks `_version`)."""
    from gnnb200 import synthetic
    from gnnb200.data import Batch, Data, host_mirror
    graphs = [Data(**g) for g in synthetic.tu_like_graphs('ENZYMES', 4, seed=1)]
    b = Batch.from_data_list(graphs)
    assert host_mirror(b, '_ptr_host') == b.ptr.tolist() and host_mirror(b, '_edge_index_host') is not None
    moved = b.clone().to('cpu')
    assert host_mirror(moved, '_ptr_host') == b.ptr.tolist()           # copies keep valid mirrors
    b.edge_index[:, 0] = b.edge_index[:, 1]                             # in-place: the mirror no longer describes the tensor
    assert host_mirror(b, '_edge_index_host') is None and host_mirror(b, '_ptr_host') is None
    assert host_mirror(moved, '_edge_index_host') is not None           # the clone has its own tensors
    b2 = Batch.from_data_list(graphs)
    b2.edge_index = b2.edge_index.clone()                               # re-assignment drops them as before
    assert host_mirror(b2, '_edge_index_host') is None
