"""Dense transforms: fp32 FFMA path within 1e-5 (relative to the output scale) of torch fp32 on the
CPU; all operand layouts, ragged sizes, split-K, bias/ReLU epilogues, and nn.Linear autograd."""
import os
import subprocess
import sys

import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import ops
from gnnb200.nn import Linear

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
TOL_F32 = 1e-5      # north_star: 1e-5 relative for fp32 paths
TOL_TF32 = 2e-2     # north_star: 2e-2 for reduced-precision GEMM paths


def _rel(got, want):
    return float((got.double().cpu() - want.double()).abs().max() / want.double().abs().max().clamp(min=1e-30))


@pytest.mark.parametrize('m,n,k', [(1, 1, 1), (5, 3, 2), (64, 64, 16), (130, 70, 33), (2708, 256, 1433), (4100, 512, 256),
                                   (300, 256, 512), (17, 1, 256), (512, 256, 20000)])
@pytest.mark.parametrize('ta,tb', [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_f32_layouts(m, n, k, ta, tb):
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    a = torch.randn((k, m) if ta else (m, k), generator=g)
    b = torch.randn((n, k) if tb else (k, n), generator=g)
    want = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
    got = ops.gemm(a.to(DEV), ta, b.to(DEV), tb, None, False, 0)
    assert got.shape == (m, n)
    assert _rel(got, want) < TOL_F32


def test_gemm_bias_relu_epilogue_and_strided_input():
    g = torch.Generator().manual_seed(0)
    big = torch.randn(300, 600, generator=g)
    a = big[:, 100:356]                       # leading dimension 600, unit inner stride
    w = torch.randn(512, 256, generator=g)
    bias = torch.randn(512, generator=g)
    want = torch.relu(a.double() @ w.double().t() + bias.double())
    got = ops.gemm(big.to(DEV)[:, 100:356], False, w.to(DEV), True, bias.to(DEV), True, 0)
    assert _rel(got, want) < TOL_F32


@pytest.mark.parametrize('rows,fin,fout', [(4096, 256, 512), (333, 21, 256), (2708, 1433, 256), (128, 768, 1)])
def test_linear_autograd(rows, fin, fout):
    g = torch.Generator().manual_seed(rows)
    ref = torch.nn.Linear(fin, fout)
    lin = Linear(fin, fout)
    lin.load_state_dict(ref.state_dict())
    lin = lin.to(DEV)
    x = torch.randn(rows, fin, generator=g)
    go = torch.randn(rows, fout, generator=g)
    xr = x.clone().requires_grad_(True)
    ref(xr).backward(go)
    xg = x.to(DEV).requires_grad_(True)
    y = lin(xg)
    y.backward(go.to(DEV))
    assert _rel(y, ref(x).detach()) < TOL_F32
    assert _rel(xg.grad, xr.grad) < TOL_F32
    assert _rel(lin.weight.grad, ref.weight.grad) < 2e-5
    assert _rel(lin.bias.grad, ref.bias.grad) < 2e-5


def test_colstats_match_batchnorm_statistics():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5000, 512, generator=g) * 3 + 50.0      # large mean: catches cancellation
    s, m2 = ops.colstats(x.to(DEV))
    assert _rel(s / 5000, x.double().mean(0)) < 1e-6
    assert _rel(m2 / 5000, x.double().var(0, unbiased=False)) < 1e-5
    assert torch.equal(ops.colsum(x.to(DEV)), s)          # same kernel, same bits


# ---- tcgen05 / TMEM / TMA path (kind::tf32): the 2e-2 tolerance class ---------------------------------
@pytest.mark.parametrize('m,n,k', [(128, 256, 32), (128, 256, 256), (256, 512, 256), (4100, 512, 256), (4100, 256, 512),
                                   (333, 256, 100), (5000, 128, 64), (1000, 64, 96), (777, 100, 256), (129, 8, 40),
                                   (20000, 256, 256), (1, 256, 256)])
@pytest.mark.parametrize('ta,tb', [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_tf32_tensor_core_path(m, n, k, ta, tb):
    g = torch.Generator().manual_seed(m + 3 * n + 7 * k)
    if (ta and m % 4) or (not ta and k % 4) or (tb and k % 4) or (not tb and n % 4):
        pytest.skip('TMA needs 16-byte row pitches')
    a = torch.randn((k, m) if ta else (m, k), generator=g)
    b = torch.randn((n, k) if tb else (k, n), generator=g)
    bias = torch.randn(n, generator=g)
    want = (a.t() if ta else a).double() @ (b.t() if tb else b).double() + bias.double()
    got = ops.gemm(a.to(DEV), ta, b.to(DEV), tb, bias.to(DEV), False, ops.PRECISIONS['tf32_strict'])
    err = _rel(got, want)
    assert err < TOL_TF32, err
    assert err < 5e-3, err          # tf32 (10-bit mantissa) with fp32 accumulate is far inside the class


def test_gemm_tf32_split_k_weight_gradient_shape():
    """dW = dY^T X with K = number of nodes: both operands MN-major, split-K, deterministic."""
    g = torch.Generator().manual_seed(1)
    dy = torch.randn(60000, 512, generator=g)
    x = torch.randn(60000, 256, generator=g)
    want = dy.double().t() @ x.double()
    got1 = ops.gemm(dy.to(DEV), True, x.to(DEV), False, None, False, ops.PRECISIONS['tf32_strict'])
    got2 = ops.gemm(dy.to(DEV), True, x.to(DEV), False, None, False, ops.PRECISIONS['tf32_strict'])
    assert torch.equal(got1, got2)
    assert _rel(got1, want) < 5e-3


def test_gemm_tf32_relu_epilogue_and_exact_small_integers():
    """Small integers are exact in tf32: any layout / swizzle / descriptor mistake shows up as a wrong
    integer, not as rounding noise."""
    g = torch.Generator().manual_seed(2)
    a = torch.randint(-4, 5, (640, 256), generator=g).float()
    w = torch.randint(-4, 5, (512, 256), generator=g).float()
    bias = torch.randint(-4, 5, (512,), generator=g).float()
    want = torch.relu(a @ w.t() + bias)
    got = ops.gemm(a.to(DEV), False, w.to(DEV), True, bias.to(DEV), True, ops.PRECISIONS['tf32_strict'])
    assert torch.equal(got.cpu(), want)


def test_linear_autograd_tf32():
    g = torch.Generator().manual_seed(4)
    ref = torch.nn.Linear(256, 512)
    lin = Linear(256, 512)
    lin.precision = 'tf32'
    lin.load_state_dict(ref.state_dict())
    lin = lin.to(DEV)
    x = torch.randn(9000, 256, generator=g)
    go = torch.randn(9000, 512, generator=g)
    xr = x.clone().requires_grad_(True)
    ref(xr).backward(go)
    xg = x.to(DEV).requires_grad_(True)
    y = lin(xg)
    y.backward(go.to(DEV))
    assert _rel(y, ref(x).detach()) < 5e-3
    assert _rel(xg.grad, xr.grad) < 5e-3
    assert _rel(lin.weight.grad, ref.weight.grad) < 5e-3
    assert _rel(lin.bias.grad, ref.bias.grad) < 1e-5


def test_tf32_falls_back_to_ffma_only_for_illegal_layouts():
    a = torch.randn(50, 1433, device=DEV)      # row pitch 1433 floats: not 16-byte aligned
    w = torch.randn(256, 1433, device=DEV)
    with pytest.raises(Exception):
        ops.gemm(a, False, w, True, None, False, ops.PRECISIONS['tf32_strict'])
    got = ops.gemm(a, False, w, True, None, False, ops.PRECISIONS['tf32'])
    assert _rel(got, a.double().cpu() @ w.double().cpu().t()) < TOL_F32


@pytest.mark.parametrize('prec', ['f32', 'tf32'])
def test_linear_residual_epilogue(prec):
    """GINLayer's `gin_conv(h) + h` is folded into the GEMM epilogue; the residual gets the identity gradient."""
    g = torch.Generator().manual_seed(6)
    ref = torch.nn.Linear(512, 256)
    lin = Linear(512, 256)
    lin.precision = prec
    lin.load_state_dict(ref.state_dict())
    lin = lin.to(DEV)
    x = torch.randn(3000, 512, generator=g)
    h = torch.randn(3000, 256, generator=g)
    go = torch.randn(3000, 256, generator=g)
    xr, hr = x.clone().requires_grad_(True), h.clone().requires_grad_(True)
    (ref(xr) + hr).backward(go)
    xg, hg = x.to(DEV).requires_grad_(True), h.to(DEV).requires_grad_(True)
    y = lin(xg, residual=hg)
    y.backward(go.to(DEV))
    tol = TOL_F32 if prec == 'f32' else 5e-3
    assert _rel(y, (ref(x) + h).detach()) < tol
    assert _rel(xg.grad, xr.grad) < tol
    assert torch.equal(hg.grad.cpu(), go)


# ---- 3xTF32: error-compensated split on the tensor pipe, fp32-class accuracy ----------------------------------
@pytest.mark.parametrize('m,n,k', [(128, 128, 32), (4100, 512, 256), (4100, 256, 512), (333, 256, 100), (1000, 64, 96),
                                   (777, 100, 256), (20000, 256, 256), (1, 256, 256)])
@pytest.mark.parametrize('ta,tb', [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_tf32x3_is_fp32_class(m, n, k, ta, tb):
    g = torch.Generator().manual_seed(m + 5 * n + 11 * k)
    if (ta and m % 4) or (not ta and k % 4) or (tb and k % 4) or (not tb and n % 4):
        pytest.skip('TMA needs 16-byte row pitches')
    a = torch.randn((k, m) if ta else (m, k), generator=g)
    b = torch.randn((n, k) if tb else (k, n), generator=g)
    bias = torch.randn(n, generator=g)
    want = (a.t() if ta else a).double() @ (b.t() if tb else b).double() + bias.double()
    got = ops.gemm(a.to(DEV), ta, b.to(DEV), tb, bias.to(DEV), False, ops.PRECISIONS['tf32x3_strict'])
    assert _rel(got, want) < TOL_F32


def test_gemm_tf32x3_split_k_and_residual():
    g = torch.Generator().manual_seed(8)
    dy = torch.randn(50000, 512, generator=g)
    x = torch.randn(50000, 256, generator=g)
    got = ops.gemm(dy.to(DEV), True, x.to(DEV), False, None, False, ops.PRECISIONS['tf32x3_strict'])
    assert _rel(got, dy.double().t() @ x.double()) < 3e-5          # K = 50,000 fp32 accumulations
    w = torch.randn(256, 512, generator=g)
    res = torch.randn(50000, 256, generator=g)
    got = ops._gemm_raw(dy.to(DEV), False, w.to(DEV), True, None, True, ops.PRECISIONS['tf32x3_strict'], res.to(DEV))
    assert _rel(got, torch.relu(dy.double() @ w.double().t() + res.double())) < TOL_F32


@pytest.mark.parametrize('prec', ['tf32_strict', 'tf32x3_strict', 'f32'])
@pytest.mark.parametrize('m,n,k', [(4100, 512, 256), (70001, 256, 512), (333, 256, 100), (31, 64, 64)])
def test_gemm_epilogue_column_statistics(prec, m, n, k):
    """BatchNorm batch statistics of the GEMM output come out of the epilogue (tensor path) / a second pass (FFMA):
    they must equal the statistics of the C that was actually written."""
    g = torch.Generator().manual_seed(m + n)
    a = torch.randn(m, k, generator=g) + 0.5
    w = torch.randn(n, k, generator=g)
    bias = torch.randn(n, generator=g) * 3
    res = torch.randn(m, n, generator=g)
    c, s, m2 = ops._gemm_raw(a.to(DEV), False, w.to(DEV), True, bias.to(DEV), False, ops.PRECISIONS[prec], res.to(DEV), True)
    cd = c.double().cpu()
    assert _rel(s, cd.sum(0)) < 1e-5
    assert _rel(m2, ((cd - cd.mean(0)) ** 2).sum(0)) < 1e-4
    tol = TOL_F32 if prec != 'tf32_strict' else 5e-3
    assert _rel(c, a.double() @ w.double().t() + bias.double() + res.double()) < tol


_TMA_STORE_CASES = [(128, 256, 32, False, True, True, False), (4100, 512, 256, False, True, True, True),
                    (333, 256, 100, False, False, True, False), (777, 100, 256, False, True, False, True),
                    (129, 8, 40, True, False, True, True), (20000, 256, 256, False, True, True, False),
                    (5000, 128, 64, True, True, False, False), (1, 256, 256, False, True, True, True),
                    (70001, 512, 256, False, True, True, False)]


def _tma_store_outputs():
    outs = []
    for m, n, k, ta, tb, with_bias, relu in _TMA_STORE_CASES:
        g = torch.Generator().manual_seed(m + 3 * n + 7 * k)
        a = torch.randn((k, m) if ta else (m, k), generator=g)
        b = torch.randn((n, k) if tb else (k, n), generator=g)
        bias = torch.randn(n, generator=g) if with_bias else None
        outs.append(ops.gemm(a.to(DEV), ta, b.to(DEV), tb, None if bias is None else bias.to(DEV), relu,
                             ops.PRECISIONS['tf32_strict']).cpu())
    return outs


@pytest.mark.skipif(not os.environ.get('GNNB200_RUN_UNVERIFIED'),
                    reason='TMA-store epilogue (GNNB200_GEMM_TMA_STORE=1): written after the round-1 GPU budget was spent')
def test_gemm_tma_store_epilogue_is_bit_identical(tmp_path):
    """Same accumulators, same bias add and ReLU, only the way the tile leaves the SM differs: the opt-in TMA-store
    epilogue must reproduce the st.global epilogue bit for bit (ragged M / N edges are clipped by the tensor map).
    The switch is read once per process, so the TMA leg runs in a child process."""
    want = _tma_store_outputs()
    out = tmp_path / 'tma.pt'
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ('import sys, importlib.util, torch; sys.path.insert(0, %r); '
            's = importlib.util.spec_from_file_location("gemm_cases", %r); t = importlib.util.module_from_spec(s); '
            's.loader.exec_module(t); torch.save(t._tma_store_outputs(), %r)') % (root, os.path.abspath(__file__), str(out))
    subprocess.run([sys.executable, '-c', code], check=True, timeout=600, env=dict(os.environ, GNNB200_GEMM_TMA_STORE='1'))
    got = torch.load(out)
    for case, w, g_ in zip(_TMA_STORE_CASES, want, got):
        assert torch.equal(w, g_), case
