"""The one-node-per-layer autograd function (gnnb200/fused.py) against the one-node-per-kernel path (ops.py): same
kernels in the same order -> identical forward bits (including the dropout mask drawn from the same seed) and
gradients equal up to the re-association of `dh = ds + A^T dz + (1+eps) dz`."""
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import models as prod
from gnnb200 import nn as gnn
from gnnb200 import synthetic
from helpers import product_batch, seeded_state_dict

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda')


def _run(fused, train, prec, graphs, seed=5):
    old = gnn.default_precision()
    gnn.set_default_precision(prec)
    prod.GINLayer.fused = fused
    try:
        m = prod.FinetuneGNN(DEV, 'ENZYMES', 'full_finetune')
        m.load_state_dict(seeded_state_dict(m, 4))
        m.train(train)
        batch = product_batch(graphs, DEV)
        x = batch.x.clone().requires_grad_(True)
        batch.x = x
        torch.manual_seed(seed)
        logits = m(batch)
        (logits * torch.arange(1, logits.numel() + 1, device=DEV).view_as(logits)).sum().backward()
        grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        stats = {k: v.clone() for k, v in m.state_dict().items() if 'running' in k or 'tracked' in k}
        return logits.detach(), x.grad.clone(), grads, stats
    finally:
        prod.GINLayer.fused = True
        gnn.set_default_precision(old)


@pytest.mark.parametrize('train', [True, False])
@pytest.mark.parametrize('prec', ['f32', 'tf32', 'tf32_fwd3'])
def test_fused_layer_equals_op_level_path(train, prec):
    graphs = synthetic.tu_like_graphs('ENZYMES', 24, seed=3)
    a = _run(True, train, prec, graphs)
    b = _run(False, train, prec, graphs)
    assert torch.equal(a[0], b[0])                                    # forward: same kernels, same order, same mask
    for k in a[3]:
        assert torch.equal(a[3][k], b[3][k]), k                       # running statistics / num_batches_tracked
    assert a[2].keys() == b[2].keys()
    scale = a[1].abs().max()
    assert float((a[1] - b[1]).abs().max() / scale) < 2e-3            # re-association (+ rare ReLU-kink flips upstream)
    for k in a[2]:
        s = b[2][k].abs().max()
        if float(s) < 1e-6:
            assert float(a[2][k].abs().max()) < 1e-4, k                # BN-fed biases: exact zeros vs rounding noise
        else:
            assert float((a[2][k] - b[2][k]).norm() / b[2][k].norm()) < 5e-3, k
