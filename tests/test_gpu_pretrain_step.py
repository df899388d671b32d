"""gnnb200.pretrain.train_step / evaluate with the product's CUDA models, tasks and gradient surgery.  The step logic
itself is pinned bit for bit on CPU (tests/test_pretrain_step.py); this checks that the same function drives the device
classes: metric keys equal the oracle's CPU run of the same step, values are finite, parameters move, schedulers
advance.  Written after the round-1 GPU budget was spent, hence opt-in until it has run once on a B200."""
import math
import os
import random

import pytest
import torch

from helpers import oracle_batch, product_batch, seeded_state_dict
from oracle import modules as orc

import gnnb200  # noqa: F401
from gnnb200 import models, pretrain, synthetic, tasks as ptasks
from gnnb200.gradient_surgery import GradientSurgery

pytestmark = pytest.mark.gpu

DOMAINS = ['MUTAG', 'ENZYMES']
S5 = ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop', 'domain_adv']


def _run(impl_models, impl_tasks, surgery_cls, make_batch, device, steps=2):
    torch.manual_seed(0)
    model = impl_models.PretrainableGNN(device, DOMAINS, S5)
    model.load_state_dict(seeded_state_dict(model, 3))
    model.train()
    grl, temp = impl_tasks.GRLScheduler(2, 5), impl_tasks.TemperatureScheduler(10)
    grl.current_step = 5
    tasks = impl_tasks.instantiate_tasks(model, S5, grl, temp)
    opt = pretrain.TaskSpecificOptimizer(model, S5)
    bal = pretrain.AdaptiveLossBalancer()
    gen = torch.Generator().manual_seed(11)
    random.seed(5)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    out = []
    for s in range(steps):
        batches = {d: make_batch(synthetic.tu_like_graphs(d, 4, seed=100 + s + i)) for i, d in enumerate(DOMAINS)}
        out.append(pretrain.train_step(model, tasks, opt, batches, gen, grl, temp, bal, surgery_cls(device), DOMAINS, epoch=1))
    return model, before, out, (grl.current_step, temp.current_step, bal.step_count)


def test_train_step_drives_the_device_classes():
    dev = torch.device('cuda')
    cpu = torch.device('cpu')
    _, _, want, counters_cpu = _run(orc, orc, orc.GradientSurgery, oracle_batch, cpu)
    model, before, got, counters = _run(models, ptasks, GradientSurgery, lambda g: product_batch(g, dev), dev)
    assert counters == counters_cpu == (7, 2, 2)
    for a, b in zip(got, want):
        assert list(a) == list(b)                                   # same metric keys in the same order
        assert all(math.isfinite(float(v)) for v in a.values())
        assert a['train/progress/epoch'] == 1 and a['train/domain_adv/lambda'] == b['train/domain_adv/lambda']
        assert a['train/loss_balancer/weight/link_pred'] == b['train/loss_balancer/weight/link_pred'] == 0.2
    moved = [k for k, v in model.state_dict().items() if v.is_floating_point() and not torch.equal(v, before[k])]
    assert any(k.startswith('gnn_backbone') for k in moved) and any(k.startswith('heads') for k in moved)


def test_evaluate_drives_the_device_classes():
    dev = torch.device('cuda')
    torch.manual_seed(0)
    model = models.PretrainableGNN(dev, DOMAINS, S5)
    grl, temp = ptasks.GRLScheduler(2, 5), ptasks.TemperatureScheduler(10)
    tasks = ptasks.instantiate_tasks(model, S5, grl, temp)
    loaders = {d: [product_batch(synthetic.tu_like_graphs(d, 3, seed=60 + 10 * i + j), dev) for j in range(2)]
               for i, d in enumerate(DOMAINS)}
    random.seed(1)
    total, metrics = pretrain.evaluate(model, tasks, loaders, torch.Generator().manual_seed(3), grl, pretrain.AdaptiveLossBalancer())
    assert total.is_cuda and math.isfinite(float(total)) and metrics['val/loss/total'] == float(total)
    assert f'val/loss/{DOMAINS[0]}/node_feat_mask' in metrics and 'val/domain_adv/loss' in metrics
    ck = pretrain.checkpoint_dict(1, model, metrics)
    assert list(ck) == ['epoch', 'model_state_dict', 'val_metrics']
