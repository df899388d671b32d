"""global_{mean,max,add}_pool fwd/bwd against the oracle's scatter (incl. torch's amax tie rule)."""
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import nn as gnn
from oracle import install_pyg_shim

install_pyg_shim()
from torch_geometric import nn as onn  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = 'cuda'

SIZES = [[5], [3, 0, 4], [1, 1, 1, 1], [32, 17, 126, 2, 64], [40] * 128, [3000, 10], [9000]]


def _batch(sizes):
    return torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))


@pytest.mark.parametrize('sizes', SIZES)
@pytest.mark.parametrize('f', [256, 37, 512])
@pytest.mark.parametrize('kind', ['mean', 'max', 'add'])
def test_pool_forward(sizes, f, kind):
    batch = _batch(sizes)
    x = torch.randn(batch.numel(), f, generator=torch.Generator().manual_seed(len(sizes) + f))
    want = getattr(onn, f'global_{kind}_pool')(x, batch, len(sizes))
    got = getattr(gnn, f'global_{kind}_pool')(x.to(DEV), batch.to(DEV), len(sizes)).cpu()
    if max(sizes) <= 2048 or kind == 'max':
        assert torch.equal(got, want)                    # same sequential order as the CPU scatter
    else:
        # chunked (2048-row partials summed in chunk order): deterministic, 1e-5 of the output scale
        assert float((got - want).abs().max() / want.abs().max()) < 1e-5


def test_pool_size_inferred_from_batch():
    batch = _batch([4, 6, 2])
    x = torch.randn(12, 64)
    assert torch.equal(gnn.global_mean_pool(x.to(DEV), batch.to(DEV)).cpu(), onn.global_mean_pool(x, batch))


@pytest.mark.parametrize('kind', ['mean', 'max', 'add'])
@pytest.mark.parametrize('sizes', [[3, 0, 4], [32, 17, 126, 2, 64], [40] * 64])
def test_pool_backward(kind, sizes):
    batch = _batch(sizes)
    f = 256
    x = torch.randn(batch.numel(), f, generator=torch.Generator().manual_seed(11))
    # ReLU-like input: exact ties at 0.0 are common after ReLU+dropout (SURVEY §7 hard parts)
    x = torch.relu(x)
    x[::3] = 0.0
    go = torch.randn(len(sizes), f, generator=torch.Generator().manual_seed(12))
    xo = x.clone().requires_grad_(True)
    getattr(onn, f'global_{kind}_pool')(xo, batch, len(sizes)).backward(go)
    xg = x.to(DEV).requires_grad_(True)
    getattr(gnn, f'global_{kind}_pool')(xg, batch.to(DEV), len(sizes)).backward(go.to(DEV))
    assert torch.equal(xg.grad.cpu(), xo.grad)


def test_max_pool_tie_rule_known_answer():
    """App. A.3: gradient split over ties, +1 tie when the max is exactly 0.0."""
    x = torch.tensor([[2.0, 0.0], [2.0, -1.0], [1.0, 0.0]], device=DEV, requires_grad=True)
    batch = torch.zeros(3, dtype=torch.long, device=DEV)
    gnn.global_max_pool(x, batch, 1).backward(torch.tensor([[6.0, 6.0]], device=DEV))
    assert torch.equal(x.grad.cpu(), torch.tensor([[3.0, 2.0], [3.0, 0.0], [0.0, 2.0]]))
