"""gnnb200.finetune on the device: the miner's similarity GEMM (normalize_rows + tcgen05 3xTF32) feeds the same selection
logic that tests/test_finetune_step.py pins on CPU; process_batch / train_step drive the product's FinetuneGNN.  Written
after the round-1 GPU budget was spent, hence opt-in until it has run once on a B200."""
import os

import pytest
import torch
import torch.nn.functional as F

from helpers import product_batch

import gnnb200  # noqa: F401
from gnnb200 import finetune, models, synthetic
from gnnb200.data import Data

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda')


@pytest.mark.parametrize('n,pairs,num_neg', [(60, 100, 40), (2708, 5278, 256), (5, 3, 12)])
def test_miner_selects_the_top_scores(n, pairs, num_neg):
    g = torch.Generator().manual_seed(n)
    h = torch.randn(n, 256, generator=g)
    existing = synthetic.random_undirected_edges(n, pairs, g)[:, ::2].contiguous()
    got = finetune.LinkPredictionHardNegativeMiner().mine_hard_negatives_for_edges(h.to(DEV), existing[:, :8].to(DEV), num_neg,
                                                                                   existing.to(DEV)).cpu()
    z = F.normalize(h.double(), dim=1)
    sim = z @ z.t()
    sim[existing[0], existing[1]] = float('-inf')
    sim[existing[1], existing[0]] = float('-inf')
    sim.fill_diagonal_(float('-inf'))
    potential = int((sim > float('-inf')).sum())
    num_hard = min(max(finetune.MIN_HARD_NEGATIVES, int(potential * finetune.HARD_NEGATIVE_RATIO)), potential, num_neg)
    assert got.size(1) == min(num_neg, potential)
    picked = sim[got[0], got[1]]
    assert bool((picked > float('-inf')).all())                                   # admissible cells only
    assert len(set(map(tuple, got.t().tolist()))) == got.size(1)
    want = torch.topk(sim.view(-1), num_hard).values
    assert torch.allclose(picked[:num_hard], want, rtol=0, atol=2e-6)             # fp32-class similarity => same ranking values


def test_process_batch_and_train_step_on_device():
    torch.manual_seed(0)
    model = models.FinetuneGNN(DEV, 'ENZYMES', 'full_finetune')
    model.train()
    opt = torch.optim.AdamW(model.param_groups)
    graphs = synthetic.tu_like_graphs('ENZYMES', 16, seed=3)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    loss, targets, pred, prob = finetune.train_step(model, opt, product_batch(graphs, DEV), DEV, 'graph_classification', 'ENZYMES')
    assert loss.is_cuda and torch.isfinite(loss) and prob.shape == (16, 6) and pred.shape == targets.shape == (16,)
    assert any(not torch.equal(v, before[k]) for k, v in model.state_dict().items() if v.is_floating_point())
    # link prediction, training mode: mined negatives + positives through the MLP link decoder
    d = synthetic.planetoid_like(300, 700, 1433, seed=4)
    lp = models.FinetuneGNN(DEV, 'Cora_LP', 'full_finetune')
    lp.train()
    train_edges = d['edge_index'][:, ::2].contiguous()
    data = Data(x=d['x'], edge_index=d['edge_index'])
    loss, targets, pred, prob = finetune.process_batch(lp, (data, train_edges[:, :64], None), DEV, 'link_prediction', 'Cora_LP',
                                                       finetune.LinkPredictionHardNegativeMiner(), train_edges.to(DEV))
    assert targets.tolist() == [1] * 64 + [0] * 64 and prob.shape == (128, 2) and torch.isfinite(loss)
