"""Fused BatchNorm1d(+ReLU)(+dropout) kernels against torch's batch_norm / relu / dropout on the CPU
(the ops the reference calls at src/models/gnn.py:19-22,31-32,41-43): forward, running statistics,
full backward (batch-statistics Jacobian), eval mode, and the dropout mask's statistics/determinism."""
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200.nn import BatchNormAct

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rel(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp(min=1e-30))


@pytest.mark.parametrize('rows,cols', [(2, 256), (37, 512), (4096, 256), (70001, 512), (300, 64)])
@pytest.mark.parametrize('relu', [True, False])
def test_train_mode_matches_torch(rows, cols, relu):
    g = torch.Generator().manual_seed(rows + cols)
    x = torch.randn(rows, cols, generator=g) * 2.0 + 3.0
    go = torch.randn(rows, cols, generator=g)
    ref = torch.nn.BatchNorm1d(cols)
    with torch.no_grad():
        ref.weight.copy_(1 + 0.2 * torch.randn(cols, generator=g))
        ref.bias.copy_(0.3 * torch.randn(cols, generator=g))
    mine = BatchNormAct(cols, relu=relu)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV)
    xr = x.clone().requires_grad_(True)
    zr = ref(xr)
    yr = torch.relu(zr) if relu else zr
    yr.backward(go)
    xg = x.to(DEV).requires_grad_(True)
    yg = mine(xg)
    yg.backward(go.to(DEV))
    assert _rel(yg, yr) < 1e-5
    # elements whose pre-ReLU value is within rounding of 0 may take the other branch of the kink: a
    # measure-zero effect of comparing two fp32 implementations, excluded here and counted below
    away = (zr.detach().abs() > 1e-4) if relu else torch.ones_like(zr, dtype=torch.bool)
    gdiff = (xg.grad.cpu() - xr.grad).abs() / xr.grad.abs().max()
    assert float(gdiff[away].max()) < 2e-5
    assert float((gdiff > 2e-5).sum()) <= 1e-5 * gdiff.numel() + 2
    # column sums over `rows` fp32 terms: summation-order noise grows with sqrt(rows); with ReLU a kink flip
    # (see above) moves one column's sum by a whole |g| term
    ptol = 5e-3 if relu else 5e-4
    assert _rel(mine.weight.grad, ref.weight.grad) < ptol
    assert _rel(mine.bias.grad, ref.bias.grad) < ptol
    assert _rel(mine.running_mean, ref.running_mean) < 1e-5
    assert _rel(mine.running_var, ref.running_var) < 1e-5
    assert int(mine.num_batches_tracked) == int(ref.num_batches_tracked) == 1


def test_eval_mode_uses_running_stats():
    g = torch.Generator().manual_seed(0)
    ref = torch.nn.BatchNorm1d(256)
    with torch.no_grad():
        ref.running_mean.copy_(torch.randn(256, generator=g))
        ref.running_var.copy_(0.5 + torch.rand(256, generator=g))
    mine = BatchNormAct(256, relu=True)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV).eval()
    ref.eval()
    x = torch.randn(1000, 256, generator=g)
    go = torch.randn(1000, 256, generator=g)
    xr = x.clone().requires_grad_(True)
    torch.relu(ref(xr)).backward(go)
    xg = x.to(DEV).requires_grad_(True)
    y = mine(xg, drop_p=0.2)            # dropout is inactive in eval mode
    y.backward(go.to(DEV))
    assert _rel(y, torch.relu(ref(x))) < 1e-5
    assert _rel(xg.grad, xr.grad) < 1e-5
    assert _rel(mine.running_mean, ref.running_mean) == 0.0


def test_fused_dropout_statistics_and_backward_mask():
    torch.manual_seed(5)
    mine = BatchNormAct(256, relu=True).to(DEV)
    x = torch.randn(20000, 256, device=DEV) + 1.0
    xg = x.clone().requires_grad_(True)
    torch.manual_seed(7)
    y = mine(xg, drop_p=0.2)
    torch.manual_seed(7)
    y2 = mine(x, drop_p=0.2)
    assert torch.equal(y, y2)                                  # reproducible under torch.manual_seed
    with torch.no_grad():
        base = mine(x, drop_p=0.0)
    pos = base > 0
    kept = (y != 0) & pos
    frac = float(kept.sum()) / float(pos.sum())
    assert abs(frac - 0.8) < 0.005                              # Bernoulli(1 - p)
    assert _rel(y[kept], base[kept] / 0.8) < 1e-5 * 10          # 1/(1-p) scaling
    per_col = kept.float().sum(0) / pos.float().sum(0).clamp(min=1)
    assert float((per_col - 0.8).abs().max()) < 0.03            # no column/row structure in the mask
    # the backward regenerates the same mask: dropped or ReLU-clamped positions get no gradient
    y.sum().backward()
    torch.manual_seed(8)
    y3 = mine(x, drop_p=0.2)
    assert not torch.equal(y3, y2)                              # a new seed gives a new mask


@pytest.mark.parametrize('parts,cols', [(1, 256), (2, 512), (8, 256), (3, 100)])
def test_merge_finalize_kernel_equals_the_moments_of_the_whole(parts, cols):
    """The node-partitioned BatchNorm statistics: per-shard (n, sum, m2) merged by gnnb200_bn_merge_finalize_f32 must give
    the mean / invstd / running buffers of the unpartitioned rows (fp32 class), with shards of different sizes."""
    from gnnb200 import _lib as L, ops, partition
    g = torch.Generator().manual_seed(parts * 1000 + cols)
    sizes = [int(v) for v in torch.randint(1, 4000, (parts,), generator=g)]
    x = torch.randn(sum(sizes), cols, generator=g) * 3 + 20.0
    moments = []
    lo = 0
    for n in sizes:
        xs = x[lo:lo + n].to(DEV)
        lo += n
        if cols % 4 == 0:
            s, m2 = ops.colstats(xs)
        else:
            s, m2 = xs.sum(0), ((xs - xs.mean(0)) ** 2).sum(0)
        moments.append(torch.stack([torch.full_like(s, float(n)), s, m2]))
    moments = torch.stack(moments).contiguous()                      # [P, 3, C]
    rm, rv = torch.zeros(cols, device=DEV), torch.ones(cols, device=DEV)
    mean, invstd = torch.empty(cols, device=DEV), torch.empty(cols, device=DEV)
    L.check(ops._invoke('gnnb200_bn_merge_finalize_f32', moments.data_ptr(), parts, cols, 1e-5, 0.1, rm.data_ptr(), rv.data_ptr(),
                        mean.data_ptr(), invstd.data_ptr(), ops._stream(moments)), 'merge')
    xd = x.double()
    assert _rel(mean, xd.mean(0)) < 1e-6
    assert _rel(invstd, 1.0 / torch.sqrt(xd.var(0, unbiased=False) + 1e-5)) < 1e-5
    assert _rel(rm, 0.1 * xd.mean(0)) < 1e-6
    assert _rel(rv, 0.9 + 0.1 * xd.var(0, unbiased=True)) < 1e-5
    n_tot, s_tot, m2_tot = partition.merge_moments(moments)             # the torch expression it replaces
    assert _rel(mean, (s_tot / n_tot).double().cpu()) < 1e-6 and _rel(invstd, torch.rsqrt(m2_tot / n_tot + 1e-5).double().cpu()) < 1e-5
