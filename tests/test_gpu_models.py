"""End-to-end parity of the drop-in modules: CUDA path vs the golden vectors produced by the
unmodified reference (tests/golden/) and vs the oracle restatement on fresh seeded inputs.

Tolerances (every individual kernel is checked at 1e-5 / bit-exact / 5e-3 for tf32 in its own test file):
  * precision 'f32' (FFMA GEMMs, 1e-5 class per op): activations and losses within 2e-4 of the output scale
    after 16 BatchNorms; gradients within 2e-3 except where a pre-ReLU value lies within rounding of 0 and
    the two implementations take different branches of the kink (measured: one such flip in layer 3 of the
    Cora-shaped run moves the upstream gradients by ~1e-3) -> relative Frobenius error <= 5e-3 and at most
    2 % of the entries off by more than 2e-3 of the tensor's scale.
  * precision 'tf32_fwd3' (the default: forward GEMMs 3xTF32 on pre-split weights, backward GEMMs plain tcgen05
    kind::tf32 — the north star's 2e-2 class): activations and losses in the fp32 class (2e-4), gradients within 2e-2
    relative Frobenius error (scripts/precision_study.py: 3e-3 .. 6e-3 expected from tf32 dX / dW).
  * plain 'tf32' in the FORWARD pass is not an end-to-end configuration: its 1e-3 activation noise flips enough ReLU
    masks through 5 x (BN, ReLU) to put the gradients at 3-7 % (round 1; reproduced on CPU by the same script), which is
    why the default compensates the forward GEMMs.  The plain-tf32 kernels are held to their per-op bound in
    tests/test_gpu_gemm.py and enter the end-to-end tests through the backward pass of 'tf32_fwd3'.
  * test_error_against_fp64_oracle states the same bars against an fp64 run of the oracle, next to what the fp32 oracle
    itself achieves against fp64 (the inherent ReLU-kink / summation-order noise floor)."""
import os
import random

import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import models as prod
from gnnb200 import synthetic
from gnnb200 import tasks as ptasks
from helpers import GOLDEN, oracle_batch, product_batch, seeded_state_dict
from oracle import modules as orc

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda')
# fp32 FFMA GEMMs: 1e-5 class per op, compounded end to end; tcgen05 tf32 GEMMs: the north star's 2e-2 class.
TOLS = {'f32': (2e-4, 2e-3), 'tf32x3': (2e-4, 2e-3), 'tf32_fwd3': (2e-4, 2e-2)}
FRO = {'f32': 5e-3, 'tf32x3': 5e-3, 'tf32_fwd3': 2e-2}
FRO_TOL = FRO['f32']
ACT_TOL, GRAD_TOL = TOLS['f32']


@pytest.fixture(params=['f32', 'tf32x3', 'tf32_fwd3'], autouse=True)
def precision(request):
    from gnnb200 import nn as gnn
    global ACT_TOL, GRAD_TOL, FRO_TOL
    old = gnn.default_precision()
    gnn.set_default_precision(request.param)
    ACT_TOL, GRAD_TOL = TOLS[request.param]
    FRO_TOL = FRO[request.param]
    yield request.param
    gnn.set_default_precision(old)
    ACT_TOL, GRAD_TOL = TOLS['f32']
    FRO_TOL = FRO['f32']
TASKS = ['node_feat_mask', 'link_pred', 'node_contrast', 'graph_contrast', 'graph_prop', 'domain_adv']


def _rel(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    if float(want.abs().max()) < 1e-6:       # mathematically zero (e.g. a bias feeding BatchNorm): noise only
        return float(got.abs().max()) / 1e-3
    return float((got - want).abs().max() / want.abs().max())


def _grad_ok(got, want, key=''):
    """Gradient parity that tolerates isolated ReLU-kink flips (see the module docstring)."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    scale = float(want.abs().max())
    if scale < 1e-6:                       # mathematically zero (a bias feeding BatchNorm): noise only
        assert float(got.abs().max()) < 1e-3, key
        return
    diff = (got - want).abs()
    fro = float(diff.norm() / want.norm())
    bad = float((diff > GRAD_TOL * scale).double().mean())
    assert fro < FRO_TOL, (key, fro)
    if FRO_TOL < 1e-2:
        assert bad <= 0.02, (key, bad)


def _golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


class _NoDropout:
    """Train-mode parity with dropout p=0 on both sides (CPU and CUDA RNG streams differ)."""

    def __init__(self, *mods):
        self.mods = mods

    def __enter__(self):
        self.old = [m.DROPOUT_RATE for m in self.mods]
        for m in self.mods:
            m.DROPOUT_RATE = 0.0

    def __exit__(self, *a):
        for m, v in zip(self.mods, self.old):
            m.DROPOUT_RATE = v


def _zero_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0


def test_finetune_enzymes_against_reference_golden():
    g = _golden('finetune_enzymes.pt')
    m = prod.FinetuneGNN(DEV, 'ENZYMES', 'full_finetune')
    m.load_state_dict(seeded_state_dict(m, g['weight_seed']))
    batch = product_batch(g['graphs'], DEV)
    m.eval()
    with torch.no_grad():
        assert _rel(m(batch), g['logits_eval']) < ACT_TOL
    m.train()
    _zero_dropout(m)
    with _NoDropout(prod):
        logits = m(batch)
        loss = torch.nn.functional.cross_entropy(logits, batch.y)
        loss.backward()
    assert _rel(logits, g['logits_train']) < ACT_TOL
    assert _rel(loss, g['loss_train']) < ACT_TOL
    params = dict(m.named_parameters())
    for k, v in g['grads'].items():
        if k.endswith('eps'):
            # d(eps) = sum(g * x) is ONE cancelling sum over every element: relative error is the per-element error
            # amplified by the cancellation (FFMA 1e-3, 3xTF32 1e-2 measured); meaningless under tf32 noise in g
            if FRO_TOL < 1e-2:
                assert _rel(params[k].grad, v) < 3e-2, k
            continue
        _grad_ok(params[k].grad, v, k)


def test_finetune_cora_small_against_reference_golden():
    g = _golden('finetune_cora_small.pt')
    m = prod.FinetuneGNN(DEV, 'Cora_NC', 'full_finetune')
    m.load_state_dict(seeded_state_dict(m, g['weight_seed']))
    m.eval()
    batch = product_batch([g['graph']], DEV)
    with torch.no_grad():
        assert _rel(m(batch, message_passing_edges=batch.edge_index), g['logits_eval']) < ACT_TOL


@pytest.mark.parametrize('task', TASKS)
def test_pretrain_tasks_against_reference_golden(task):
    g = _golden('pretrain_s5_small.pt')
    m = prod.PretrainableGNN(DEV, g['domains'], TASKS)
    m.load_state_dict(seeded_state_dict(m, g['weight_seed']))
    m.eval()
    temp, grl = ptasks.TemperatureScheduler(100), ptasks.GRLScheduler(10, 10)
    grl.current_step = 80
    obj = ptasks.instantiate_tasks(m, [task], grl, temp)[task]
    gen = torch.Generator().manual_seed(11)
    random.seed(11)
    loss, per = obj.compute_loss({d: product_batch(g['graphs'][d], DEV) for d in g['domains']}, gen)
    loss.backward()
    assert _rel(loss, g['losses'][task]) < ACT_TOL
    for d in g['domains']:
        assert _rel(per[d], g['per_domain'][task][d]) < ACT_TOL
    params = dict(m.named_parameters())
    for k, v in g['grads'][task].items():
        if k.endswith('eps'):
            if FRO_TOL < 1e-2:
                assert _rel(params[k].grad, v) < 3e-2, k
            continue
        _grad_ok(params[k].grad, v, k)


@pytest.mark.parametrize('layers', [3, 5])
def test_backbone_cora_shape_against_oracle(layers):
    """BASELINE config 1 (N=2,708, E=10,556, F=1,433, H=256) at L=3 (BASELINE) and L=5 (reference)."""
    d = synthetic.cora_like(seed=42)
    a = orc.FinetuneGNN(torch.device('cpu'), 'Cora_NC', 'full_finetune', num_layers=layers)
    b = prod.FinetuneGNN(DEV, 'Cora_NC', 'full_finetune', num_layers=layers)
    sd = seeded_state_dict(a, 5)
    a.load_state_dict(sd)
    b.load_state_dict(sd)
    a.train()
    b.train()
    _zero_dropout(a)
    _zero_dropout(b)
    with _NoDropout(prod, orc):
        ba = oracle_batch([d])
        bb = product_batch([d], DEV)
        ha = a.gnn_backbone(a.input_encoder(ba.x), ba.edge_index)
        hb = b.gnn_backbone(b.input_encoder(bb.x), bb.edge_index)
        # a fixed random read-out keeps the gradient well conditioned (sum(h) through BatchNorm is almost
        # constant, which would leave only ReLU-kink noise to compare)
        w = torch.randn(ha.shape, generator=torch.Generator().manual_seed(9))
        (ha * w).sum().backward()
        (hb * w.to(DEV)).sum().backward()
    assert _rel(hb, ha) < ACT_TOL
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    for k in ('gnn_backbone.layers.0.gin_conv.nn.0.weight', 'input_encoder.linear.weight',
              f'gnn_backbone.layers.{layers - 1}.gin_conv.nn.3.weight', f'gnn_backbone.layers.{layers - 1}.batch_norm.weight'):
        _grad_ok(pb[k].grad, pa[k].grad, k)


def test_bn_running_stats_follow_the_reference():
    """Every domain batch goes through BN separately and updates the running stats (App. C.8)."""
    graphs = synthetic.tu_like_graphs('ENZYMES', 16, seed=2)
    a = orc.FinetuneGNN(torch.device('cpu'), 'ENZYMES', 'full_finetune')
    b = prod.FinetuneGNN(DEV, 'ENZYMES', 'full_finetune')
    sd = seeded_state_dict(a, 6)
    a.load_state_dict(sd)
    b.load_state_dict(sd)
    a.train()
    b.train()
    _zero_dropout(a)
    _zero_dropout(b)
    with _NoDropout(prod, orc), torch.no_grad():
        a(oracle_batch(graphs))
        b(product_batch(graphs, DEV))
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if 'running' in k:
            assert _rel(sb[k], sa[k]) < ACT_TOL, k
        if 'num_batches_tracked' in k:
            assert int(sb[k]) == int(sa[k]) == 1


def test_train_mode_dropout_runs_and_is_seeded():
    graphs = synthetic.tu_like_graphs('ENZYMES', 8, seed=3)
    m = prod.FinetuneGNN(DEV, 'ENZYMES', 'full_finetune')
    m.train()
    batch = product_batch(graphs, DEV)
    torch.manual_seed(1)
    y1 = m(batch)
    torch.manual_seed(1)
    y2 = m(batch)
    assert torch.equal(y1, y2) and torch.isfinite(y1).all()


def test_error_against_fp64_oracle(precision):
    """The bar stated against ground truth: the product's error against an fp64 run of the oracle, next to the fp32 oracle's
    own error against the same fp64 run (summation order and ReLU-kink flips: the floor any fp32 implementation sits on).
    fp32 class: product within 2x the fp32 oracle's error or 5e-3 Frobenius, whichever is larger (a single kink flip moves a
    gradient by ~1e-3); tensor-core class ('tf32_fwd3'): 2e-2 Frobenius on every gradient, 2e-4 on the logits."""
    import copy
    graphs = synthetic.tu_like_graphs('ENZYMES', 16, seed=0)
    a32 = orc.FinetuneGNN(torch.device('cpu'), 'ENZYMES', 'full_finetune')
    sd = seeded_state_dict(a32, 3)
    a32.load_state_dict(sd)
    a64 = orc.FinetuneGNN(torch.device('cpu'), 'ENZYMES', 'full_finetune')
    a64.load_state_dict(sd)
    a64 = a64.double()
    b = prod.FinetuneGNN(DEV, 'ENZYMES', 'full_finetune')
    b.load_state_dict(sd)
    for train in (False, True):
        outs = {}
        for name, m, mk, dt in (('f64', a64, oracle_batch, torch.float64), ('f32', a32, oracle_batch, torch.float32),
                                ('gpu', b, lambda gr: product_batch(gr, DEV), torch.float32)):
            m.train(train)
            _zero_dropout(m)
            for p_ in m.parameters():
                p_.grad = None
            with _NoDropout(prod, orc):
                bt = mk(graphs)
                x = bt.x.to(dt).clone().requires_grad_(True)
                bt.x = x
                y = m(bt)
                w = torch.randn(y.shape, generator=torch.Generator().manual_seed(9)).to(y.device, dt)
                (y * w).sum().backward()
            grads = {k: p_.grad.detach().double().cpu() for k, p_ in m.named_parameters() if p_.grad is not None}
            grads['x'] = x.grad.detach().double().cpu()
            outs[name] = (y.detach().double().cpu(), grads)
        fro = lambda u, v: float((u - v).norm() / v.norm().clamp(min=1e-300))       # noqa: E731
        ref_y, ref_g = outs['f64']
        floor_y = fro(outs['f32'][0], ref_y)
        assert fro(outs['gpu'][0], ref_y) < max(2 * floor_y, ACT_TOL), (train, fro(outs['gpu'][0], ref_y), floor_y)
        for k, v in ref_g.items():
            if float(v.abs().max()) < 1e-9 or k.endswith('eps'):
                continue                                        # analytically zero (bias feeding BatchNorm) / one cancelling dot
            floor = fro(outs['f32'][1][k], v)
            got = fro(outs['gpu'][1][k], v)
            assert got < max(2 * floor, FRO_TOL), (train, k, got, floor)
