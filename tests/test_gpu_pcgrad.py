"""Flat-buffer gradient surgery (gnnb200/gradient_surgery.py + csrc/pcgrad.cu) against the oracle's restatement
of the reference's PCGrad variant (pinned to the reference in test_oracle_reference.py): same shuffle (random.seed),
same conflict / projection counts, same final gradients including the "only parameters of the first shuffled task
are overwritten" quirk."""
import random

import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200.gradient_surgery import GradientSurgery
from oracle import modules as orc

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda')


def _make(device):
    torch.manual_seed(1)
    m = torch.nn.ModuleDict({'shared': torch.nn.Linear(64, 48), 'deep': torch.nn.Linear(48, 48), 'a': torch.nn.Linear(48, 8),
                             'b': torch.nn.Linear(48, 5), 'c': torch.nn.Linear(48, 16), 'unused': torch.nn.Linear(3, 3)})
    return m.to(device)


def _losses(m, x):
    h = torch.tanh(m['deep'](torch.relu(m['shared'](x))))
    return {'t1': m['a'](h).pow(2).mean(), 't2': -m['b'](h).sum() * 0.01,
            't3': (m['c'](h) - 1).abs().mean() - m['a'](h).mean(), 't4': m['b'](h).pow(2).mean() + h.mean()}


@pytest.mark.parametrize('shuffle_seed', [0, 1, 2, 3, 4, 5])
def test_matches_oracle(shuffle_seed):
    x = torch.randn(200, 64, generator=torch.Generator().manual_seed(7))
    mo, mp = _make(torch.device('cpu')), _make(DEV)
    random.seed(shuffle_seed)
    metrics_o = orc.GradientSurgery(torch.device('cpu')).apply_gradient_surgery(mo, _losses(mo, x), ['t1', 't2', 't3', 't4'])
    random.seed(shuffle_seed)
    metrics_p = GradientSurgery(DEV).apply_gradient_surgery(mp, _losses(mp, x.to(DEV)), ['t1', 't2', 't3', 't4'])
    assert metrics_p['gradient_surgery/total_projections'] == metrics_o['gradient_surgery/total_projections']
    assert metrics_p['gradient_surgery/total_conflicts'] == metrics_o['gradient_surgery/total_conflicts']
    assert metrics_o['gradient_surgery/total_conflicts'] > 0           # the case exercises real projections
    po, pp = dict(mo.named_parameters()), dict(mp.named_parameters())
    for k in po:
        assert (po[k].grad is None) == (pp[k].grad is None), k
        if po[k].grad is not None:
            scale = po[k].grad.abs().max().clamp(min=1e-12)
            assert float((pp[k].grad.cpu() - po[k].grad).abs().max() / scale) < 1e-5, k


def test_single_task_is_a_no_op():
    m = _make(DEV)
    assert GradientSurgery(DEV).apply_gradient_surgery(m, {'only': m['a'](torch.randn(4, 48, device=DEV)).sum()}, ['only']) == {}
