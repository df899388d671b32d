"""The public operator surface ``torch.ops.gnnb200.*`` (torch.library definitions with CUDA kernels, fake
kernels and autograd formulas) gives the same bits as the package's internal fast path."""
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import ops
from gnnb200.graph import Graph

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def test_registered_ops_exist_with_schemas():
    names = ['csr_build', 'segment_ptr', 'coalesce', 'aggregate', 'dot', 'gin_aggregate', 'segment_pool', 'segment_pool_bwd',
             'rows_gather', 'rows_gather_bwd', 'rows_scatter', 'gemm', 'colsum', 'colstats', 'linear', 'bn_batch_stats',
             'bn_act', 'bn_act_bwd', 'bn_act_bwd_reduce', 'bn_act_bwd_apply', 'lp_features', 'lp_features_bwd',
             'ntxent_fwd', 'ntxent_bwd']
    for n in names:
        assert hasattr(torch.ops.gnnb200, n), n


def test_dispatcher_path_matches_fast_path_with_autograd():
    g = torch.Generator().manual_seed(0)
    n, e, f = 300, 2500, 256
    ei = torch.randint(0, n, (2, e), generator=g).to(DEV)
    gr = Graph(ei, n)
    x = torch.randn(n, f, generator=g).to(DEV)
    eps = torch.tensor([0.1], device=DEV)
    go = torch.randn(n, f, generator=g).to(DEV)
    outs = []
    for call in (ops.gin_aggregate, torch.ops.gnnb200.gin_aggregate):
        xa, ea = x.clone().requires_grad_(True), eps.clone().requires_grad_(True)
        z = call(xa, ea, gr.rowptr, gr.col, gr.rowptr_t, gr.col_t)
        z.backward(go)
        outs.append((z.detach(), xa.grad, ea.grad))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    # linear + fused BN through the dispatcher
    w = torch.randn(512, 256, generator=g).to(DEV)
    bias = torch.randn(512, generator=g).to(DEV)
    res = []
    for call in (ops.linear, torch.ops.gnnb200.linear):
        xa, wa = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        y = call(xa, wa, bias, 0, None)
        y.sum().backward()
        res.append((y.detach(), xa.grad, wa.grad))
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_fake_kernels_give_shapes_without_running():
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(100, 256, device=DEV)
        rp = torch.empty(101, dtype=torch.int32, device=DEV)
        col = torch.empty(900, dtype=torch.int32, device=DEV)
        out = torch.ops.gnnb200.aggregate(x, rp, col, 0, None, None, None)
        assert out.shape == (100, 256)
        y = torch.ops.gnnb200.gemm(x, False, torch.empty(512, 256, device=DEV), True, None, False, 0)
        assert y.shape == (100, 512)
