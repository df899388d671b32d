"""Aggregation forward/backward at BASELINE config 5's FULL size on the device, through the checker that
tests/test_full_scale_properties.py validates on CPU.  (Named to sort last: the driver runs the GPU suite with -x.)"""
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import synthetic
from test_full_scale_properties import check_aggregation_properties


@pytest.mark.gpu
def test_aggregation_full_c5_scale():
    from gnnb200 import _lib as L, ops
    from gnnb200.graph import Graph
    dev = torch.device('cuda')
    cache = {}

    def graph(ei):
        if cache.get('ei') is not ei:
            cache['ei'], cache['g'] = ei, Graph(ei, synthetic.C5_NODES)
        return cache['g']

    def forward(x, eps, ei):
        g = graph(ei)
        return ops._aggregate_raw(x, g.rowptr, g.col, L.AGG_SUM, x, eps, None)

    def backward(gy, eps, ei):
        g = graph(ei)
        return ops._aggregate_raw(gy, g.rowptr_t, g.col_t, L.AGG_SUM, gy, eps, None)

    check_aggregation_properties(forward, backward, dev, synthetic.C5_NODES, synthetic.C5_EDGES, 256)


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['tf32_strict', 'tf32x3_strict'])
def test_gemm_full_c5_rows_exact_on_small_integers(precision):
    """The three GEMM layouts of a GIN layer at M (or K) = 2,449,029 rows.  Small integers are exact in tf32 and every
    partial sum stays far below 2^24, so the tensor-core result must EQUAL cuBLAS fp32 — any tile-index, TMA-coordinate
    or split-K mistake that only shows beyond 2^31 bytes turns up as a wrong integer."""
    from gnnb200 import ops
    dev = torch.device('cuda')
    torch.backends.cuda.matmul.allow_tf32 = False
    prec = ops.PRECISIONS[precision]
    m = synthetic.C5_NODES
    g = torch.Generator(device=dev).manual_seed(5)
    z = torch.randint(-2, 3, (m, 256), device=dev, generator=g).float()
    w1 = torch.randint(-2, 3, (512, 256), device=dev, generator=g).float()
    b1 = torch.randint(-2, 3, (512,), device=dev, generator=g).float()
    a1 = ops.gemm(z, False, w1, True, b1, True, prec)                      # forward  Y = relu(Z W1^T + b)   (K-major / K-major)
    assert torch.equal(a1, torch.relu(z @ w1.t() + b1))
    assert torch.equal(a1[-300:], torch.relu(z[-300:] @ w1.t() + b1))     # the ragged last tile
    da1 = torch.randint(-2, 3, (m, 512), device=dev, generator=g).float()
    del a1
    dz = ops.gemm(da1, False, w1, False, None, False, prec)                # dX = dY W                        (K-major / MN-major)
    assert torch.equal(dz, da1 @ w1)
    del dz
    dw = ops.gemm(da1, True, z, False, None, False, prec)                  # dW = dY^T Z, K = rows, split-K   (MN-major / MN-major)
    want = (da1.double().t() @ z.double()).float()
    assert float(want.abs().max()) < 2 ** 23
    assert torch.equal(dw, want)


@pytest.mark.gpu
@pytest.mark.parametrize('cols', [256, 512])
def test_batchnorm_full_c5_rows_against_torch(cols):
    """Fused BatchNorm+ReLU forward/backward over 2,449,029 rows against torch's own CUDA batch_norm + relu (fp32)."""
    from gnnb200.nn import BatchNormAct
    dev = torch.device('cuda')
    rows = synthetic.C5_NODES
    g = torch.Generator(device=dev).manual_seed(cols)
    x = torch.randn(rows, cols, device=dev, generator=g) * 2.0 + 3.0
    go = torch.randn(rows, cols, device=dev, generator=g)
    ref = torch.nn.BatchNorm1d(cols).to(dev)
    with torch.no_grad():
        ref.weight.copy_(1 + 0.2 * torch.randn(cols, device=dev, generator=g))
        ref.bias.copy_(0.3 * torch.randn(cols, device=dev, generator=g))
    mine = BatchNormAct(cols, relu=True).to(dev)
    mine.load_state_dict(ref.state_dict())
    xr = x.clone().requires_grad_(True)
    zr = ref(xr)
    torch.relu(zr).backward(go)
    xg = x.clone().requires_grad_(True)
    yg = mine(xg)
    yg.backward(go)
    rel = lambda a, b: float((a.detach() - b.detach()).abs().max() / b.detach().abs().max())   # noqa: E731
    assert rel(yg, torch.relu(zr)) < 1e-5
    away = zr.detach().abs() > 1e-4                                        # away from the ReLU kink (see tests/test_gpu_bn.py)
    gdiff = (xg.grad - xr.grad).abs() / xr.grad.abs().max()
    assert float(gdiff[away].max()) < 5e-5
    assert rel(mine.weight.grad, ref.weight.grad) < 5e-3 and rel(mine.bias.grad, ref.bias.grad) < 5e-3
    assert rel(mine.running_mean, ref.running_mean) < 1e-5 and rel(mine.running_var, ref.running_var) < 1e-5


@pytest.mark.gpu
def test_aggregation_full_c5_every_row_against_the_c_oracle():
    """Not a sample: the WHOLE [2,449,029 x 256] output of the forward aggregation at BASELINE config 5's size, bit for bit
    against the plain-C edge-order loop (oracle/c/oracle_c.c, one host thread, ~1 min), and the transposed pass likewise."""
    from gnnb200 import _lib as L, ops
    from gnnb200.graph import Graph
    from oracle import c_oracle
    dev = torch.device('cuda')
    d = synthetic.products_like(synthetic.C5_NODES, synthetic.C5_EDGES, 4, seed=11)
    ei = d['edge_index']
    x = torch.randn(synthetic.C5_NODES, 256, generator=torch.Generator().manual_seed(12))
    eps = torch.tensor([0.25])
    g = Graph(ei.to(dev), synthetic.C5_NODES)
    xd, ed = x.to(dev), eps.to(dev)
    got = ops._aggregate_raw(xd, g.rowptr, g.col, L.AGG_SUM, xd, ed, None).cpu()
    assert torch.equal(got, c_oracle.gin_aggregate(x, ei, 0.25))
    del got
    got_t = ops._aggregate_raw(xd, g.rowptr_t, g.col_t, L.AGG_SUM, xd, ed, None).cpu()
    assert torch.equal(got_t, c_oracle.gin_aggregate(x, ei, 0.25, transposed=True))
