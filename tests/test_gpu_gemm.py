"""Dense transforms: fp32 FFMA path within 1e-5 (relative to the output scale) of torch fp32 on the
CPU; all operand layouts, ragged sizes, split-K, bias/ReLU epilogues, and nn.Linear autograd."""
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
    """Row pitches that are not a multiple of 16 bytes (feature widths 1433, 3703, 21, 7, 37): the strict mode refuses them,
    the module-level modes re-pitch the operand once (ops._tma_rows, cached on the tensor) and stay on the tensor cores;
    only outputs the kernel cannot store 128 bits at a time (N % 4 != 0, N = 1) take the FFMA kernel."""
    a = torch.randn(50, 1433, device=DEV)      # row pitch 1433 floats: not 16-byte aligned
    w = torch.randn(256, 1433, device=DEV)
    with pytest.raises(Exception):
        ops.gemm(a, False, w, True, None, False, ops.PRECISIONS['tf32_strict'])
    want = a.double().cpu() @ w.double().cpu().t()
    ops.reset_counters()
    got = ops.gemm(a, False, w, True, None, False, ops.PRECISIONS['tf32'])
    err = _rel(got, want)
    assert 1e-6 < err < 5e-3, err               # tf32 rounding is visible: the tensor-core kernel ran
    assert ops._tma_rows(a) is ops._tma_rows(a) and ops._tma_rows(a).stride(0) == 1436          # cached, re-pitched
    assert _rel(ops.gemm(a, False, w, True, None, False, ops.PRECISIONS['tf32x3']), want) < 3e-5     # K = 1433 accumulations
    a.add_(1.0)                                 # an in-place edit invalidates the cached copy
    assert _rel(ops.gemm(a, False, w, True, None, False, ops.PRECISIONS['tf32x3']), a.double().cpu() @ w.double().cpu().t()) < 3e-5
    w1 = torch.randn(1, 256, device=DEV)        # N = 1 (the link decoder's last layer): FFMA, fp32 class
    x = torch.randn(300, 256, device=DEV)
    assert _rel(ops.gemm(x, False, w1, True, None, False, ops.PRECISIONS['tf32']), x.double().cpu() @ w1.double().cpu().t()) < TOL_F32


@pytest.mark.parametrize('rows,fin,fout', [(2708, 1433, 256), (3327, 3703, 256), (4340, 21, 256), (600, 7, 256), (900, 37, 256)])
def test_encoder_widths_run_on_the_tensor_cores(rows, fin, fout):
    """InputEncoder's Linear for every dataset's feature width (src/models/gnn.py:13, src/data/data_setup.py:31-41) under the
    default precision: forward in the fp32 class, dW through the transposed product x^T g (the kernel stores N % 4 == 0)."""
    g = torch.Generator().manual_seed(rows + fin)
    ref = torch.nn.Linear(fin, fout)
    lin = Linear(fin, fout)
    lin.precision = 'tf32_fwd3'
    lin.load_state_dict(ref.state_dict())
    lin = lin.to(DEV)
    x = torch.randn(rows, fin, generator=g)
    go = torch.randn(rows, fout, generator=g)
    xr = x.clone().requires_grad_(True)
    ref(xr).backward(go)
    xd = x.to(DEV)
    ops.reset_counters()
    y = lin(xd)
    y.backward(go.to(DEV))
    assert 'gnnb200_linear_x3w_f32' in ops.call_counts()
    assert _rel(y, ref(x).detach()) < (TOL_F32 if fin <= 1024 else 3e-5)     # fp32 accumulation over up to 3703 terms
    assert lin.weight.grad.shape == ref.weight.grad.shape and lin.weight.grad.is_contiguous()
    err = _rel(lin.weight.grad, ref.weight.grad)
    assert 1e-7 < err < 5e-3, err               # tf32 noise: the weight gradient came from the tensor-core kernel
    assert _rel(lin.bias.grad, ref.bias.grad) < 2e-5
    xg = x.to(DEV).requires_grad_(True)          # input gradient (not needed by the encoders, but part of the op)
    lin(xg).backward(go.to(DEV))
    assert _rel(xg.grad, xr.grad) < 5e-3


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


# ---- 3xTF32 with pre-split weights: the forward GEMM of every Linear under the default precision 'tf32_fwd3' -------------
def _trunc_tf32(t):
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


def test_tf32_mma_truncates_operands():
    """kind::tf32 reads the upper 19 bits of each fp32 word (sign, exponent, 10 mantissa bits) and ignores the rest: a GEMM
    on operands with the 13 low mantissa bits cleared gives the same BITS as on the raw operands.  gnnb200_linear_x3w_f32's
    raw_hi mode (X3 = 3 in csrc/gemm_tcgen05.cu) relies on this to use the raw tiles as the hi operands."""
    g = torch.Generator().manual_seed(21)
    for m, n, k, ta, tb in ((640, 256, 256, False, True), (300, 512, 96, False, False), (512, 128, 320, True, False)):
        a = torch.randn((k, m) if ta else (m, k), generator=g).to(DEV)
        b = torch.randn((n, k) if tb else (k, n), generator=g).to(DEV)
        prec = ops.PRECISIONS['tf32_strict']
        raw = ops.gemm(a, ta, b, tb, None, False, prec)
        assert torch.equal(raw, ops.gemm(_trunc_tf32(a), ta, _trunc_tf32(b), tb, None, False, prec))
        assert not torch.equal(a, _trunc_tf32(a))


def test_split_weight_is_exact_and_cached():
    w = torch.randn(512, 256, device=DEV)
    ops.X3W_RAW_HI, old = False, ops.X3W_RAW_HI
    try:
        hi, lo = ops.split_weight(w)
        assert torch.equal(hi, _trunc_tf32(w)) and torch.equal(hi + lo, w)
        assert float(lo.abs().max()) <= float(w.abs().max()) * 2.0 ** -10
        assert ops.split_weight(w)[1] is lo                       # cached on the tensor
        w.mul_(1.5)                                               # an optimizer step bumps the version counter
        hi2, lo2 = ops.split_weight(w)
        assert lo2 is not lo and torch.equal(hi2 + lo2, w)
    finally:
        ops.X3W_RAW_HI = old


@pytest.mark.parametrize('raw_hi', [False, True])
@pytest.mark.parametrize('m,n,k', [(128, 256, 32), (4100, 512, 256), (4100, 256, 512), (333, 256, 100), (1000, 64, 96),
                                   (777, 100, 256), (20000, 256, 256), (1, 256, 256), (70001, 512, 256), (513, 128, 768),
                                   (300, 40, 64), (50, 256, 1433), (128, 1, 256)])
def test_linear_x3w_is_fp32_class(monkeypatch, raw_hi, m, n, k):
    """Forward of a Linear under 'tf32_fwd3': 3xTF32 with the weights split once (gnnb200_linear_x3w_f32, both split modes),
    bias / residual / ReLU epilogues, ragged tiles; layouts TMA cannot take (K % 4 != 0, N = 1) fall to the FFMA kernel."""
    monkeypatch.setattr(ops, 'X3W_RAW_HI', raw_hi)
    g = torch.Generator().manual_seed(m + 5 * n + 11 * k)
    x = torch.randn(m, k, generator=g)
    w = torch.randn(n, k, generator=g)
    bias = torch.randn(n, generator=g)
    res = torch.randn(m, n, generator=g)
    prec = ops.PRECISIONS['tf32_fwd3']
    want = x.double() @ w.double().t() + bias.double()
    xd, wd, bd = x.to(DEV), w.to(DEV), bias.to(DEV)
    tol = TOL_F32 if k <= 1024 else 3e-5                                  # fp32 accumulation over K terms
    assert _rel(ops._linear_fwd_raw(xd, wd, bd, False, prec), want) < tol
    assert _rel(ops._linear_fwd_raw(xd, wd, bd, True, prec, res.to(DEV)), torch.relu(want + res.double())) < tol
    y, s, m2 = ops._linear_fwd_raw(xd, wd, bd, False, prec, None, True)
    yd = y.double().cpu()
    assert _rel(y, want) < tol and _rel(s, yd.sum(0)) < 1e-5 and _rel(m2, ((yd - yd.mean(0)) ** 2).sum(0)) < 1e-4


def test_linear_autograd_tf32_fwd3():
    """The default precision: forward in the fp32 class, backward GEMMs plain tf32 (5e-3 per op)."""
    g = torch.Generator().manual_seed(4)
    ref = torch.nn.Linear(256, 512)
    lin = Linear(256, 512)
    lin.precision = 'tf32_fwd3'
    lin.load_state_dict(ref.state_dict())
    lin = lin.to(DEV)
    x = torch.randn(9000, 256, generator=g)
    go = torch.randn(9000, 512, generator=g)
    xr = x.clone().requires_grad_(True)
    ref(xr).backward(go)
    xg = x.to(DEV).requires_grad_(True)
    y = lin(xg)
    y.backward(go.to(DEV))
    assert _rel(y, ref(x).detach()) < TOL_F32
    assert _rel(xg.grad, xr.grad) < 5e-3
    assert _rel(lin.weight.grad, ref.weight.grad) < 5e-3
    with torch.no_grad():
        lin.weight.add_(0.25)                                     # the cached split must follow the weights
        ref.weight.add_(0.25)
    assert _rel(lin(x.to(DEV)), ref(x).detach()) < TOL_F32


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
                    (132, 8, 40, True, False, True, True), (20000, 256, 256, False, True, True, False),
                    (5000, 128, 64, True, True, False, False), (1, 256, 256, False, True, True, True),
                    (70001, 512, 256, False, True, True, False)]


@pytest.mark.parametrize('m,n,k,ta,tb,with_bias,relu', _TMA_STORE_CASES)
def test_gemm_tma_store_epilogue_is_bit_identical(m, n, k, ta, tb, with_bias, relu):
    """Same accumulators, same bias add and ReLU, only the way the tile leaves the SM differs: the TMA-store epilogue (what
    a plain tf32 GEMM without residual takes) must reproduce the st.global epilogue bit for bit, ragged M / N edges
    included (clipped by the tensor map).  A residual of zeros forces the st.global epilogue: x + 0.0 == x exactly."""
    g = torch.Generator().manual_seed(m + 3 * n + 7 * k)
    a = torch.randn((k, m) if ta else (m, k), generator=g).to(DEV)
    b = torch.randn((n, k) if tb else (k, n), generator=g).to(DEV)
    bias = torch.randn(n, generator=g).to(DEV) if with_bias else None
    prec = ops.PRECISIONS['tf32_strict']
    via_tma = ops._gemm_raw(a, ta, b, tb, bias, relu, prec)
    via_st_global = ops._gemm_raw(a, ta, b, tb, bias, relu, prec, torch.zeros(m, n, device=DEV))
    assert torch.equal(via_tma, via_st_global)
