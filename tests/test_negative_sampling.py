"""gnnb200.utils.batched_negative_sampling (the deterministic branch vectorised over all graphs of the batch, the
`random.sample` branch per graph in order) against the oracle's restatement of PyG's per-graph algorithm (SURVEY.md
App. A.5): bit-identical negatives in the same order and the same position of Python's `random` stream afterwards."""
import random

import numpy as np
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import synthetic, utils
from gnnb200.data import Batch, Data
from oracle import install_pyg_shim

install_pyg_shim()
import torch_geometric.utils as pyg_utils  # noqa: E402  (the oracle shim)


def _check(graphs, quota=None, seed=0):
    b = Batch.from_data_list([Data(x=g['x'], edge_index=g['edge_index']) for g in graphs])
    und = pyg_utils.to_undirected(b.edge_index)
    q = b.edge_index.size(1) if quota is None else quota
    random.seed(seed)
    want = pyg_utils.batched_negative_sampling(und, b.batch, q)
    after_want = random.random()
    random.seed(seed)
    got = utils.batched_negative_sampling(und, b.batch, q)
    after_got = random.random()
    assert got.dtype == want.dtype == torch.long and torch.equal(got, want)
    assert after_got == after_want                                   # same number of random.sample draws consumed
    return got


@pytest.mark.parametrize('domain', ['ENZYMES', 'PROTEINS', 'MUTAG', 'NCI1'])
@pytest.mark.parametrize('seed', [0, 1])
def test_tu_shaped_batches_take_the_deterministic_branch(domain, seed):
    graphs = synthetic.tu_like_graphs(domain, 32, seed=seed)
    neg = _check(graphs)
    n = sum(g['x'].size(0) * (g['x'].size(0) - 1) for g in graphs)
    e = pyg_utils.to_undirected(Batch.from_data_list([Data(x=g['x'], edge_index=g['edge_index']) for g in graphs]).edge_index).size(1)
    assert 0 < neg.size(1) <= n - e          # at most ALL non-edges of every graph (the usual outcome, App. A.5 consequence)


def test_mixed_batch_interleaves_random_and_deterministic_graphs_in_order():
    graphs = (synthetic.tu_like_graphs('ENZYMES', 5, seed=3) + [synthetic.planetoid_like(500, 900, 21, seed=1)] +
              synthetic.tu_like_graphs('ENZYMES', 4, seed=4) + [synthetic.planetoid_like(400, 700, 21, seed=2)])
    for seed in (0, 9):
        _check(graphs, seed=seed)


def test_single_large_graph_uses_python_random():
    d = synthetic.cora_like(42)
    neg = _check([d], seed=4)
    assert neg.size(1) == d['edge_index'].size(1)                    # exactly the quota (App. A.5)


@pytest.mark.parametrize('quota', [None, 3, 1, 1000])
def test_edge_cases(quota):
    graphs = [{'x': torch.randn(2, 3), 'edge_index': torch.tensor([[0, 1], [1, 0]])},          # complete: no negative exists
              {'x': torch.randn(1, 3), 'edge_index': torch.empty(2, 0, dtype=torch.long)},     # single node
              {'x': torch.randn(4, 3), 'edge_index': torch.tensor([[0, 1, 2, 2], [1, 0, 2, 3]])},   # self loop
              {'x': torch.randn(3, 3), 'edge_index': torch.empty(2, 0, dtype=torch.long)},     # edgeless, in the middle
              {'x': torch.randn(5, 3), 'edge_index': torch.tensor([[0, 4], [4, 0]])},
              {'x': torch.randn(6, 3), 'edge_index': torch.empty(2, 0, dtype=torch.long)}]     # trailing edgeless: skipped upstream
    _check(graphs, quota=quota)


def test_no_edges_at_all():
    b = Batch.from_data_list([Data(x=torch.randn(3, 2), edge_index=torch.empty(2, 0, dtype=torch.long))])
    assert utils.batched_negative_sampling(b.edge_index, b.batch, 5).shape == (2, 0)


def test_ungrouped_edges_are_refused():
    b = Batch.from_data_list([Data(x=torch.randn(3, 2), edge_index=torch.tensor([[0], [1]])),
                              Data(x=torch.randn(3, 2), edge_index=torch.tensor([[0], [2]]))])
    with pytest.raises(ValueError):
        utils.batched_negative_sampling(b.edge_index.flip(1), b.batch, 2)


@pytest.mark.parametrize('domain', ['ENZYMES', 'MUTAG'])
def test_link_prediction_negatives_from_the_host_mirror(domain):
    """LinkPredictionTask._negatives on a gnnb200.loader batch (structure mirrored on the host: symmetrise + coalesce +
    sampling never touch the device) equals the oracle's to_undirected + batched_negative_sampling on the same batch."""
    from gnnb200 import loader
    from gnnb200.tasks import LinkPredictionTask
    graphs = [Data(**g) for g in synthetic.tu_like_graphs(domain, 20, seed=3)]
    graphs[7] = Data(x=graphs[7].x, edge_index=torch.empty(2, 0, dtype=torch.long), y=graphs[7].y,
                     graph_properties=graphs[7].graph_properties)
    picks = [4, 7, 7, 0, 19, 12, 4, 3]
    b = loader.ResidentDomain(graphs).batch_of(picks)
    random.seed(2)
    got = LinkPredictionTask._negatives(b, b.edge_index)
    random.seed(2)
    want = pyg_utils.batched_negative_sampling(pyg_utils.to_undirected(b.edge_index), b.batch, b.edge_index.size(1))
    assert got.dtype == torch.long and torch.equal(got, want)
    assert np.array_equal(utils.to_undirected_host(b._edge_index_host, int(b.edge_index.max()) + 1),
                          pyg_utils.to_undirected(b.edge_index).numpy())


def test_py_sample_range_is_random_sample_bit_for_bit():
    """gnnb200.utils.py_sample_range (Mersenne Twister + CPython's two `sample` branches in C) against the interpreter's own
    random.sample(range(n), k): same values, same stream position afterwards (also mid-block and across the pool / set
    branch boundary n <= 21 + 4 ** ceil(log(3k, 4)))."""
    chooser = random.Random(123)
    sizes = [1, 2, 5, 6, 21, 22, 30, 100, 1000, 9900, 16405, 16406, 20000, 70000, 7_300_000, 2 ** 31 + 5, 2 ** 32 + 1]
    for trial in range(120):
        n = chooser.choice(sizes)
        k = min(n, chooser.choice([0, 1, 2, 5, 6, 7, 20, 50, 4000, 11600, n if n < 50000 else 100]))
        random.seed(trial)
        burn = [random.random() for _ in range(trial * 7 % 700)]
        want, after_want, gauss_want = random.sample(range(n), k), random.random(), random.gauss(0, 1)
        random.seed(trial)
        assert burn == [random.random() for _ in range(trial * 7 % 700)]
        got = utils.py_sample_range(n, k)
        assert got.dtype == np.int64 and got.tolist() == want, (n, k)
        assert (random.random(), random.gauss(0, 1)) == (after_want, gauss_want), (n, k)
