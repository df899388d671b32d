"""The device branch of gnnb200.utils.batched_negative_sampling (bitmap complement, csrc/negsample.cu) against the oracle's
restatement of PyG's per-graph sampler: the same negatives in the same order, bit for bit, and the same
position of Python's `random` stream afterwards — also when some graph of the batch falls back to the host sampler."""
import random

import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import ops, synthetic, utils
from gnnb200.data import Batch, Data
from oracle import install_pyg_shim

install_pyg_shim()
import torch_geometric.utils as pyg_utils  # noqa: E402  (the oracle shim)

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _check(graphs, quota=None, seed=0, expect_device=True):
    b = Batch.from_data_list([Data(x=g['x'], edge_index=g['edge_index']) for g in graphs])
    und = pyg_utils.to_undirected(b.edge_index)
    q = b.edge_index.size(1) if quota is None else quota
    random.seed(seed)
    want = pyg_utils.batched_negative_sampling(und, b.batch, q)
    after_want = random.random()
    random.seed(seed)
    ops.reset_counters()
    got = utils.batched_negative_sampling(und.to(DEV), b.batch.to(DEV), q)
    after_got = random.random()
    calls = ops.call_counts()
    assert calls.get('gnnb200_negsample_count_i64', 0) == 1
    if expect_device is not None:
        assert (calls.get('gnnb200_negsample_write_i64', 0) == 1) == (expect_device and want.size(1) > 0)
    assert got.is_cuda and got.dtype == torch.long and torch.equal(got.cpu(), want)
    assert after_got == after_want
    return got


@pytest.mark.parametrize('domain', ['ENZYMES', 'PROTEINS', 'MUTAG', 'NCI1'])
@pytest.mark.parametrize('count,seed', [(32, 0), (128, 1), (1, 2)])
def test_tu_shaped_batches_on_the_device(domain, count, seed):
    # whether the whole batch stays on the device depends on the data (a large graph, or a single graph whose quota is its
    # own edge count, draws from random.sample upstream and sends the batch to the host sampler): both outcomes are checked
    # against the oracle; the device-only outcome is pinned by the tests below
    _check(synthetic.tu_like_graphs(domain, count, seed=seed), expect_device=None)


@pytest.mark.parametrize('quota', [None, 3, 1, 1000])
def test_edge_cases_on_the_device(quota):
    graphs = [{'x': torch.randn(2, 3), 'edge_index': torch.tensor([[0, 1], [1, 0]])},          # complete: no negative exists
              {'x': torch.randn(1, 3), 'edge_index': torch.empty(2, 0, dtype=torch.long)},     # single node
              {'x': torch.randn(4, 3), 'edge_index': torch.tensor([[0, 1, 2, 2], [1, 0, 2, 3]])},   # self loop
              {'x': torch.randn(3, 3), 'edge_index': torch.empty(2, 0, dtype=torch.long)},     # edgeless, in the middle
              {'x': torch.randn(5, 3), 'edge_index': torch.tensor([[0, 4], [4, 0]])},
              {'x': torch.randn(6, 3), 'edge_index': torch.empty(2, 0, dtype=torch.long)}]     # trailing edgeless: skipped upstream
    # a small quota makes upstream draw from random.sample for the larger graphs: the kernel reports it and the host sampler
    # runs; quota 1000 keeps every graph in the deterministic branch
    _check(graphs, quota=quota, expect_device=True if quota == 1000 else None)


def test_a_graph_that_needs_python_random_sends_the_batch_to_the_host_sampler():
    graphs = (synthetic.tu_like_graphs('ENZYMES', 5, seed=3) + [synthetic.planetoid_like(500, 900, 21, seed=1)] +
              synthetic.tu_like_graphs('ENZYMES', 4, seed=4))
    _check(graphs, seed=5, expect_device=False)


def test_dense_and_large_graphs_fill_whole_bitmap_words():
    """n (n - 1) not a multiple of 32, nearly complete graphs, and the largest size the shared-memory bitmap takes."""
    g = torch.Generator().manual_seed(0)
    graphs = []
    for n, keep in ((33, 0.9), (64, 0.5), (724, 0.01), (7, 1.0), (100, 0.97)):
        full = torch.ones(n, n).triu(1).nonzero().t()
        ei = full[:, torch.rand(full.size(1), generator=g) < keep]
        graphs.append({'x': torch.randn(n, 2), 'edge_index': torch.cat([ei, ei.flip(0)], dim=1)})
    _check(graphs, quota=10 ** 7)


def test_ungrouped_edges_are_refused_on_the_device():
    b = Batch.from_data_list([Data(x=torch.randn(3, 2), edge_index=torch.tensor([[0], [1]])),
                              Data(x=torch.randn(3, 2), edge_index=torch.tensor([[0], [2]]))])
    with pytest.raises(ValueError):
        utils.batched_negative_sampling(b.edge_index.flip(1).to(DEV), b.batch.to(DEV), 2)
