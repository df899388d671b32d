"""gnnb200.finetune (hard-negative miner, process_batch, train_step) pinned on CPU against the UNMODIFIED reference
`src/finetune/finetune.py` imported over the torch_geometric shim.  The miner's similarity kernel is a CUDA kernel; here
it is replaced by the reference's own CPU expression, so what is pinned is the selection logic (forbidden cells, counts,
top-k over the flat matrix, index decoding, the random-fill branch).  Container only."""
import pytest
import torch
import torch.nn.functional as F

from helpers import oracle_batch
from oracle.reference_loader import load_reference, reference_available

import gnnb200  # noqa: F401
from gnnb200 import finetune as prod
from gnnb200 import synthetic

pytestmark = pytest.mark.skipif(not reference_available(), reason='/root/reference not present')


@pytest.fixture(scope='module')
def ref():
    ns = load_reference()
    import src.finetune.finetune as ref_finetune
    ns.finetune = ref_finetune
    return ns


def _cpu_similarity(h):
    z = F.normalize(h, dim=1)
    return torch.mm(z, z.t())


def _canonical(edges):
    """Set of directed (u, v) pairs."""
    return set(map(tuple, edges.t().tolist()))


def _scores(h, edges):
    s = _cpu_similarity(h)
    return s[edges[0], edges[1]]


@pytest.mark.parametrize('n,pairs,num_neg,seed', [
    (60, 100, 40, 0),          # the usual case: k = num_negatives < 0.3 * admissible cells, no random fill
    (200, 500, 256, 1),        # Cora_LP batch size
    (5, 3, 12, 2),             # tiny graph: 0.3 * 14 admissible cells -> 8 hard (the minimum) + 4 random
    (4, 6, 5, 3),              # complete graph: no admissible cell -> empty result
    (6, 0, 3, 4),              # no existing edges at all
    (3, 1, 10, 5),             # more negatives requested than admissible cells exist
])
def test_miner_equals_reference(ref, n, pairs, num_neg, seed):
    g = torch.Generator().manual_seed(seed)
    h = torch.randn(n, 32, generator=g)
    existing = synthetic.random_undirected_edges(n, pairs, g)[:, ::2].contiguous() if pairs else torch.empty(2, 0, dtype=torch.long)
    pos = existing[:, :max(1, existing.size(1) // 2)]
    torch.manual_seed(77)
    want = ref.finetune.LinkPredictionHardNegativeMiner().mine_hard_negatives_for_edges(h, pos, num_neg, existing)
    torch.manual_seed(77)
    got = prod.LinkPredictionHardNegativeMiner(similarity=_cpu_similarity).mine_hard_negatives_for_edges(h, pos, num_neg, existing)
    assert got.shape == want.shape and got.dtype == want.dtype == torch.long
    if want.size(1) == 0:
        return
    # admissible: never a self pair, never an existing edge in either direction
    banned = _canonical(existing) | _canonical(existing.flip(0))
    assert all(u != v and (u, v) not in banned for u, v in _canonical(got))
    assert len(_canonical(got)) == got.size(1)                                   # no duplicates
    # the same scores in the same (descending) order for the hard part; identical cells wherever the score is not tied
    sw, sg = _scores(h, want), _scores(h, got)
    num_potential = n * n - len(banned | {(i, i) for i in range(n)})
    num_hard = min(max(prod.MIN_HARD_NEGATIVES, int(num_potential * prod.HARD_NEGATIVE_RATIO)), num_potential, num_neg)
    assert torch.equal(sw[:num_hard], sg[:num_hard])
    kth = float(sw[:num_hard].min()) if num_hard else float('inf')
    untied = lambda e, s: {c for c, v in zip(map(tuple, e[:, :num_hard].t().tolist()), s[:num_hard].tolist()) if v > kth}  # noqa: E731
    assert untied(want, sw) == untied(got, sg)
    if want.size(1) > num_hard:                                                  # random fill: same generator stream
        if _canonical(want[:, :num_hard]) == _canonical(got[:, :num_hard]):
            assert torch.equal(want[:, num_hard:], got[:, num_hard:])
        assert not (_canonical(got[:, num_hard:]) & (_canonical(got[:, :num_hard]) | _canonical(got[:, :num_hard].flip(0))))


def _state(model):
    return {k: v.clone() for k, v in model.state_dict().items()}


@pytest.mark.parametrize('domain,strategy', [('ENZYMES', 'full_finetune'), ('PTC_MR', 'linear_probe')])
def test_process_batch_graph_classification(ref, domain, strategy):
    torch.manual_seed(0)
    model = ref.finetune_model.FinetuneGNN(torch.device('cpu'), domain, strategy)
    graphs = synthetic.tu_like_graphs(domain, 12, seed=8, num_classes=prod.NUM_CLASSES[domain])
    out = []
    for fn in (ref.finetune.process_batch, prod.process_batch):
        model.train()
        torch.manual_seed(5)
        out.append(fn(model, oracle_batch(graphs), torch.device('cpu'), 'graph_classification', domain, None, None))
    for a, b in zip(*out):
        assert torch.equal(a, b)


def test_process_batch_node_classification_and_train_step(ref):
    from torch_geometric.data import Data
    d = synthetic.planetoid_like(300, 600, 1433, seed=3)
    data = Data(x=d['x'], edge_index=d['edge_index'], y=torch.randint(0, 7, (300,), generator=torch.Generator().manual_seed(1)))
    idx = torch.arange(0, 300, 3)
    batch = (data, idx, data.y[idx])
    mp_edges = d['edge_index'][:, :800]
    models = []
    for fn_step in ('ref', 'prod'):
        torch.manual_seed(0)
        model = ref.finetune_model.FinetuneGNN(torch.device('cpu'), 'Cora_NC', 'full_finetune')
        opt = torch.optim.AdamW(model.param_groups)
        model.train()
        torch.manual_seed(9)
        if fn_step == 'ref':
            loss, *_ = ref.finetune.process_batch(model, batch, torch.device('cpu'), 'node_classification', 'Cora_NC', None, mp_edges)
            opt.zero_grad()
            loss.backward()
            opt.step()
        else:
            loss, *_ = prod.train_step(model, opt, batch, torch.device('cpu'), 'node_classification', 'Cora_NC', None, mp_edges)
        models.append((loss.detach(), _state(model)))
    assert torch.equal(models[0][0], models[1][0])
    assert all(torch.equal(models[0][1][k], models[1][1][k]) for k in models[0][1])


@pytest.mark.parametrize('training', [True, False])
def test_process_batch_link_prediction(ref, training):
    from torch_geometric.data import Data
    d = synthetic.planetoid_like(120, 300, 1433, seed=4)
    data = Data(x=d['x'], edge_index=d['edge_index'])
    train_edges = d['edge_index'][:, ::2].contiguous()
    pos = train_edges[:, :64]
    if training:
        batch = (data, pos, None)
    else:
        neg = torch.randint(0, 120, (2, 64), generator=torch.Generator().manual_seed(2))
        batch = (data, torch.cat([pos, neg], dim=1), torch.cat([torch.ones(64), torch.zeros(64)]))
    out = []
    for miner, fn in ((ref.finetune.LinkPredictionHardNegativeMiner(), ref.finetune.process_batch),
                      (prod.LinkPredictionHardNegativeMiner(similarity=_cpu_similarity), prod.process_batch)):
        torch.manual_seed(0)
        model = ref.finetune_model.FinetuneGNN(torch.device('cpu'), 'Cora_LP', 'full_finetune')
        model.train(training)
        torch.manual_seed(6)
        out.append(fn(model, batch, torch.device('cpu'), 'link_prediction', 'Cora_LP', miner, train_edges))
    (la, ta, pa, qa), (lb, tb, pb, qb) = out
    assert torch.equal(ta, tb) and ta.numel() == 128
    if training:
        # mined negatives may swap (i, j) with (j, i) among exactly tied scores: the decoder features are symmetric in
        # (u, v), so the multiset of probabilities is the same and the mean differs at most by summation order
        assert torch.equal(qa[:64], qb[:64])
        assert torch.allclose(torch.sort(qa[64:, 1]).values, torch.sort(qb[64:, 1]).values, rtol=0, atol=1e-6)
        assert abs(float(la.detach()) - float(lb.detach())) < 1e-6
    else:
        assert torch.equal(la, lb) and torch.equal(pa, pb) and torch.equal(qa, qb)


def test_constants_match_reference(ref):
    f = ref.finetune
    assert (prod.BATCH_SIZES, prod.EPOCHS, prod.HARD_NEGATIVE_RATIO, prod.MIN_HARD_NEGATIVES, prod.PATIENCE_FRACTION) == \
        (f.BATCH_SIZES, f.EPOCHS, f.HARD_NEGATIVE_RATIO, f.MIN_HARD_NEGATIVES, f.PATIENCE_FRACTION)
    assert prod.NUM_CLASSES == f.NUM_CLASSES
