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


# ---- head tails and loss sums (csrc/heads.cu) against the eager torch calls the reference makes -------------------------
@pytest.mark.parametrize('shape', [(64, 12), (128, 256), (3000, 100), (1, 1), (2100, 300)])
def test_mse_sum_matches_torch(shape):
    g = torch.Generator().manual_seed(shape[0])
    a, b = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    ar = a.clone().requires_grad_(True)
    want = torch.nn.functional.mse_loss(ar, b, reduction='sum')
    (want * 0.37).backward()
    ad = a.to(DEV).requires_grad_(True)
    got = ops.mse_sum(ad, b.to(DEV))
    (got * 0.37).backward()
    assert got.dim() == 0 and _rel(got, want.detach()) < 1e-5             # 2100 x 300 > 2^19 elements: grid + finish launch
    assert _rel(ad.grad, ar.grad) < 1e-6
    assert torch.equal(got, ops.mse_sum(ad.detach(), b.to(DEV)))          # deterministic


@pytest.mark.parametrize('n', [1, 777, 127000, 600000])
def test_sigmoid_bce_sum_matches_torch(n):
    g = torch.Generator().manual_seed(n)
    z = torch.randn(n, generator=g) * 4
    z[: min(n, 3)] = torch.tensor([120.0, -120.0, 30.0])[: min(n, 3)]      # saturated: the -100 clamp and the 1e-12 floor
    t = (torch.rand(n, generator=g) < 0.5).float()
    zr = z.clone().requires_grad_(True)
    pr = torch.sigmoid(zr)
    want = torch.nn.functional.binary_cross_entropy(pr, t, reduction='sum')
    (want / n).backward()
    zd = z.to(DEV).requires_grad_(True)
    probs, got = ops.sigmoid_bce_sum(zd, t.to(DEV))
    (got / n).backward()
    assert _rel(probs, pr.detach()) < 1e-6 and _rel(got, want.detach()) < 1e-5
    assert _rel(zd.grad, zr.grad) < 1e-5


@pytest.mark.parametrize('rows,cols', [(128, 4), (2708, 7), (1, 6), (300000, 2)])
def test_cross_entropy_sum_matches_torch(rows, cols):
    g = torch.Generator().manual_seed(rows + cols)
    z = torch.randn(rows, cols, generator=g) * 3
    t = torch.randint(0, cols, (rows,), generator=g)
    zr = z.clone().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(zr, t, reduction='sum')
    (want / rows).backward()
    zd = z.to(DEV).requires_grad_(True)
    got, lse = ops.cross_entropy_sum(zd, t.to(DEV))
    (got / rows).backward()
    assert _rel(got, want.detach()) < 1e-5 and _rel(lse, torch.logsumexp(z, dim=1)) < 1e-6
    assert _rel(zd.grad, zr.grad) < 1e-5


@pytest.mark.parametrize('p', [0.0, 0.2, 0.5])
@pytest.mark.parametrize('rows,fin,fout', [(500, 256, 128), (37, 768, 256), (4000, 256, 512)])
def test_linear_act_matches_linear_relu_dropout(p, rows, fin, fout):
    """One hidden layer of MLPHead (Linear -> ReLU -> Dropout, reference src/models/heads.py:41-45): p = 0 against torch
    exactly in the fp32 class; p > 0: kept entries are relu(xW^T+b)/(1-p), the kept fraction is 1-p, the same seed gives the
    same mask, and the backward uses exactly the forward's mask."""
    g = torch.Generator().manual_seed(rows + fin)
    x, w, b = torch.randn(rows, fin, generator=g), torch.randn(fout, fin, generator=g) * 0.1, torch.randn(fout, generator=g)
    go = torch.randn(rows, fout, generator=g)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    base = torch.relu(torch.nn.functional.linear(xr, wr, br))
    xd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    y = ops.linear_act(xd, wd, bd, ops.PRECISIONS['f32'], p, 1234)
    if p == 0.0:
        base.backward(go)
        y.backward(go.to(DEV))
        assert _rel(y, base.detach()) < 1e-5
        assert _rel(xd.grad, xr.grad) < 2e-5 and _rel(wd.grad, wr.grad) < 2e-5 and _rel(bd.grad, br.grad) < 2e-5
        return
    yc, bc = y.detach().cpu(), base.detach()
    kept = yc != 0
    pos = bc > 0
    assert not (kept & ~pos).any()
    assert abs(float((kept & pos).sum()) / float(pos.sum()) - (1 - p)) < 0.02
    assert _rel(yc[kept], bc[kept] / (1 - p)) < 1e-5
    assert torch.equal(y.detach(), ops.linear_act(xd.detach(), wd.detach(), bd.detach(), ops.PRECISIONS['f32'], p, 1234))
    assert not torch.equal(y.detach(), ops.linear_act(xd.detach(), wd.detach(), bd.detach(), ops.PRECISIONS['f32'], p, 99))
    mask = kept.float() / (1 - p)                                          # replay the product's mask through torch
    (base * mask).backward(go)
    y.backward(go.to(DEV))
    assert _rel(xd.grad, xr.grad) < 2e-5 and _rel(wd.grad, wr.grad) < 2e-5 and _rel(bd.grad, br.grad) < 2e-5


def test_gradient_reversal_scales_by_minus_lambda():
    x = torch.randn(40, 256, device=DEV, requires_grad=True)
    y = ops.gradient_reversal(x, 0.35)
    assert torch.equal(y, x)
    go = torch.randn(40, 256, device=DEV)
    y.backward(go)
    assert _rel(x.grad, (-0.35 * go).cpu()) < 1e-6


def test_mlp_head_modules_match_the_oracle_heads():
    """MLPHead / DomainClassifierHead / MLPLinkPredictor.loss with the fused tails against the oracle's eager modules."""
    from gnnb200 import models as prod
    torch.manual_seed(0)
    a = orc.MLPHead([256, 512, 12])
    b = prod.MLPHead([256, 512, 12])
    b.load_state_dict(a.state_dict())
    b = b.to(DEV)
    a.eval()
    b.eval()
    x = torch.randn(77, 256)
    xr, xd = x.clone().requires_grad_(True), x.to(DEV).requires_grad_(True)
    ya, yb = a(xr), b(xd)
    ya.square().sum().backward()
    yb.square().sum().backward()
    assert _rel(yb, ya.detach()) < 1e-5 and _rel(xd.grad, xr.grad) < 2e-5
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert _rel(pb.grad, pa.grad) < 2e-5, k
    da, db = orc.DomainClassifierHead(), prod.DomainClassifierHead()
    db.load_state_dict(da.state_dict())
    db = db.to(DEV)
    da.eval()
    db.eval()
    xr, xd = x.clone().requires_grad_(True), x.to(DEV).requires_grad_(True)
    t = torch.randint(0, 4, (77,))
    la = torch.nn.functional.cross_entropy(da(xr, 0.6), t, reduction='sum')
    lb, _ = ops.cross_entropy_sum(db(xd, 0.6), t.to(DEV))
    la.backward()
    lb.backward()
    assert _rel(lb, la.detach()) < 1e-5 and _rel(xd.grad, xr.grad) < 2e-5      # reversed and scaled by lambda on both sides
    pa_, pb_ = orc.MLPLinkPredictor(), prod.MLPLinkPredictor()
    pb_.load_state_dict(pa_.state_dict())
    pb_ = pb_.to(DEV)
    pa_.eval()
    pb_.eval()
    h = torch.randn(300, 256)
    e = torch.randint(0, 300, (2, 900))
    lab = (torch.rand(900) < 0.5).float()
    hr, hd = h.clone().requires_grad_(True), h.to(DEV).requires_grad_(True)
    probs_a = pa_(hr, e)
    la = torch.nn.functional.binary_cross_entropy(probs_a, lab, reduction='sum')
    probs_b, lb = pb_.loss(hd, e.to(DEV), lab.to(DEV))
    la.backward()
    lb.backward()
    assert _rel(probs_b, probs_a.detach()) < 1e-5 and _rel(lb, la.detach()) < 1e-5 and _rel(hd.grad, hr.grad) < 2e-5
    assert _rel(pb_(hd.detach(), e.to(DEV)), probs_a.detach()) < 1e-5            # the module's own forward still returns probs
