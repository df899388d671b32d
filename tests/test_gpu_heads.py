"""Pre-training head kernels against the oracle: row gather/scatter (NFM), link-prediction decoder
features fwd/bwd, NT-Xent fwd/bwd, deterministic dot."""
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import ops, tasks
from oracle import modules as orc

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _f32_default():
    """These tests check the 1e-5 class: module-level Linear layers use the fp32 FFMA GEMM here."""
    from gnnb200 import nn as gnn
    old = gnn.default_precision()
    gnn.set_default_precision('f32')
    yield
    gnn.set_default_precision(old)


def _rel(got, want):
    return float((got.double().cpu() - want.double()).abs().max() / want.double().abs().max().clamp(min=1e-30))


def test_rows_gather_scatter_forward_backward():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(500, 256, generator=g)
    idx = torch.randint(0, 500, (120,), generator=g)
    tok = torch.randn(256, generator=g)
    go = torch.randn(120, 256, generator=g)
    xr = x.clone().requires_grad_(True)
    xr[idx].backward(go)
    xg = x.to(DEV).requires_grad_(True)
    out = ops.rows_gather(xg, idx.to(DEV))
    out.backward(go.to(DEV))
    assert torch.equal(out.detach().cpu(), x[idx])
    torch.testing.assert_close(xg.grad.cpu(), xr.grad, rtol=1e-6, atol=1e-6)
    # mask-token write (pretrain_model.py:84-86): unique indices as the reference draws them
    uniq = torch.randperm(500, generator=g)[:60]
    tr = tok.clone().requires_grad_(True)
    base = x.clone().requires_grad_(True)
    masked = base.clone()
    masked[uniq] = tr.expand(60, -1)
    gm = torch.randn(500, 256, generator=g)
    masked.backward(gm)
    tg = tok.to(DEV).requires_grad_(True)
    bg = x.to(DEV).requires_grad_(True)
    mg = ops.rows_scatter(bg, tg, uniq.to(DEV))
    mg.backward(gm.to(DEV))
    assert torch.equal(mg.detach().cpu(), masked.detach())
    assert torch.equal(bg.grad.cpu(), base.grad)
    torch.testing.assert_close(tg.grad.cpu(), tr.grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('n,e,h', [(50, 400, 256), (300, 5000, 256), (40, 100, 64)])
def test_lp_features_forward_backward(n, e, h):
    g = torch.Generator().manual_seed(e)
    x = torch.randn(n, h, generator=g)
    x[::5] = x[1::5][: x[::5].size(0)]           # some exactly equal rows -> |.| kink at 0
    edges = torch.randint(0, n, (2, e), generator=g)
    go = torch.randn(e, 3 * h, generator=g)
    xr = x.clone().requires_grad_(True)
    hu, hv = xr[edges[0]], xr[edges[1]]
    fr = torch.cat([hu + hv, hu * hv, torch.abs(hu - hv)], dim=1)
    fr.backward(go)
    xg = x.to(DEV).requires_grad_(True)
    fg = ops.lp_features(xg, edges.to(DEV))
    fg.backward(go.to(DEV))
    assert torch.equal(fg.detach().cpu(), fr.detach())
    assert _rel(xg.grad, xr.grad) < 1e-5


def test_link_predictor_module_matches_oracle():
    from gnnb200.models import MLPLinkPredictor
    a = orc.MLPLinkPredictor().eval()
    b = MLPLinkPredictor().eval()
    b.load_state_dict(a.state_dict())
    b = b.to(DEV)
    g = torch.Generator().manual_seed(1)
    h = torch.randn(200, 256, generator=g)
    edges = torch.randint(0, 200, (2, 1500), generator=g)
    assert _rel(b(h.to(DEV), edges.to(DEV)), a(h, edges).detach()) < 1e-5


@pytest.mark.parametrize('m,d', [(1, 128), (2, 128), (37, 128), (64, 128), (700, 128), (100, 64), (50, 16)])
@pytest.mark.parametrize('temp', [0.5, 0.2])
def test_ntxent_forward_backward(m, d, temp):
    g = torch.Generator().manual_seed(m + d)
    z1 = torch.randn(m, d, generator=g)
    z2 = z1 + 0.3 * torch.randn(m, d, generator=g)
    a1, a2 = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
    lo, so = orc.nt_xent(a1, a2, temp)
    lo.backward()
    b1, b2 = z1.to(DEV).requires_grad_(True), z2.to(DEV).requires_grad_(True)
    lg, sg = tasks.nt_xent(b1, b2, temp)
    lg.backward()
    assert int(sg) == int(so) == 2 * m
    assert abs(float(lg) - float(lo)) < 1e-5 * max(abs(float(lo)), 1.0)
    scale = a1.grad.abs().max().clamp(min=1e-3)     # m == 1: the gradient is exactly zero
    assert float((b1.grad.cpu() - a1.grad).abs().max() / scale) < 2e-5
    assert float((b2.grad.cpu() - a2.grad).abs().max() / scale) < 2e-5


def test_dot_is_deterministic_and_accurate():
    g = torch.Generator().manual_seed(5)
    a = torch.randn(1_000_003, generator=g)
    b = torch.randn(1_000_003, generator=g)
    d1 = ops.dot(a.to(DEV), b.to(DEV))
    d2 = ops.dot(a.to(DEV), b.to(DEV))
    assert torch.equal(d1, d2)
    want = float((a.double() * b.double()).sum())
    assert abs(float(d1) - want) < 1e-4 * max(1.0, abs(want)) + 0.05


@pytest.mark.parametrize('m,d', [(64, 128), (300, 128), (1500, 128), (700, 64)])
@pytest.mark.parametrize('prec,tol', [('tf32x3_strict', 2e-5), ('tf32_strict', 5e-3)])
def test_ntxent_tensor_core_path(m, d, prec, tol):
    """Similarity matrix and its two backward contractions on tcgen05 (tasks.nt_xent takes this path for 2M >= 2048)."""
    g = torch.Generator().manual_seed(m * 3 + d)
    z1 = torch.randn(m, d, generator=g)
    z2 = z1 + 0.3 * torch.randn(m, d, generator=g)
    a = torch.cat([z1, z2]).requires_grad_(True)
    lo, _ = orc.nt_xent(a[:m], a[m:], 0.3)
    lo.backward()
    b = torch.cat([z1, z2]).to(DEV).requires_grad_(True)
    lg = ops.ntxent_tensor_core(b, 0.3, ops.PRECISIONS[prec])
    lg.backward()
    assert abs(float(lg) - float(lo)) < tol * abs(float(lo))
    scale = a.grad.abs().max()
    assert float((b.grad.cpu() - a.grad).abs().max() / scale) < max(tol * 5, 1e-4)


def test_nt_xent_dispatches_by_size():
    from gnnb200 import nn as gnn
    old = gnn.default_precision()
    try:
        gnn.set_default_precision('tf32x3')
        g = torch.Generator().manual_seed(0)
        z1, z2 = torch.randn(1100, 128, generator=g), torch.randn(1100, 128, generator=g)
        lo, so = orc.nt_xent(z1, z2, 0.5)
        before = dict(ops.call_counts())
        lg, sg = tasks.nt_xent(z1.to(DEV), z2.to(DEV), 0.5)
        after = ops.call_counts()
        assert after.get('gnnb200_ntxent_sim_fwd_f32', 0) == before.get('gnnb200_ntxent_sim_fwd_f32', 0) + 1   # tensor path
        assert int(sg) == int(so) and abs(float(lg) - float(lo)) < 2e-5 * abs(float(lo))
    finally:
        gnn.set_default_precision(old)
