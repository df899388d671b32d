"""Aggregation at BASELINE config 5's FULL size (2,449,029 nodes, 61,859,140 edges, 256 features): too large for the CPU
oracle as a whole, so the same contract is checked through size-independent properties plus an exact comparison on
sampled destination rows:
  * all-ones identity: every row of A·1 + (1+eps)·1 equals in-degree + 1 + eps, exactly;
  * sampled rows: the fp32 sum over a row's neighbours in ORIGINAL edge order, then + (1+eps)·x_i, recomputed on the CPU
    from the COO list, equals the kernel's row bit for bit (forward on the by-destination CSR, backward on the by-source CSR);
  * adjointness: <A x, y> == <x, A^T y> (the backward pass is the transpose of the forward pass);
  * determinism: a second run gives the same bits.
The checker itself is validated on CPU against the oracle's aggregation at a small size (the `not gpu` test below), so a
failure on the GPU box means the kernel, not the test."""
import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import synthetic


def check_aggregation_properties(forward, backward, device, n, e, f, samples=48, seed=0):
    """forward(x, eps, edge_index) -> A x + (1+eps) x ;  backward(g, eps, edge_index) -> A^T g + (1+eps) g."""
    d = synthetic.products_like(n, e, 4, seed=seed + 1, device=device)
    ei = d['edge_index']
    src, dst = ei[0], ei[1]
    eps = torch.tensor([0.25], device=device)
    gen = torch.Generator(device=device).manual_seed(seed)
    # ---- all-ones identity (exact: small integers + 1.25) ----
    ones = torch.ones(n, f, device=device)
    z1 = forward(ones, eps, ei)
    indeg = torch.bincount(dst, minlength=n).to(torch.float32)
    assert torch.equal(z1, (indeg + 1.25).view(-1, 1).expand(n, f))
    g1 = backward(ones, eps, ei)
    outdeg = torch.bincount(src, minlength=n).to(torch.float32)
    assert torch.equal(g1, (outdeg + 1.25).view(-1, 1).expand(n, f))
    del ones, z1, g1
    # ---- sampled rows, bit for bit against an edge-order fp32 sum on the CPU ----
    x = torch.randn(n, f, device=device, generator=gen)
    z = forward(x, eps, ei)
    zt = backward(x, eps, ei)
    rows = torch.randint(0, n, (samples,), device=device, generator=gen)
    rows = torch.cat([rows, dst[:4], src[:4]]).unique()
    one_plus_eps = (1 + eps).cpu()
    for key, other, got in ((dst, src, z), (src, dst, zt)):
        hit = torch.isin(key, rows)
        k_sel, o_sel = key[hit].cpu(), other[hit].cpu()                  # original edge order is preserved by the mask
        x_nb = x[other[hit]].cpu()
        for r in rows.tolist():
            mine = k_sel == r
            acc = torch.zeros(1, f).index_add_(0, torch.zeros(int(mine.sum()), dtype=torch.long), x_nb[mine])
            want = acc[0] + one_plus_eps * x[r].cpu()
            assert torch.equal(got[r].cpu(), want), f'row {r} ({int(mine.sum())} neighbours)'
        del hit, k_sel, o_sel, x_nb
    # ---- adjointness of forward and backward (fp32 accumulation of 6e8 products: compare in fp64, relative) ----
    y = torch.randn(n, f, device=device, generator=gen)
    lhs = float((z.double() * y.double()).sum())
    rhs = float((x.double() * backward(y, eps, ei).double()).sum())
    scale = float(z.double().norm() * y.double().norm())
    assert abs(lhs - rhs) <= 1e-6 * scale, (lhs, rhs, scale)
    # ---- determinism ----
    assert torch.equal(z, forward(x, eps, ei))


def _oracle_forward(x, eps, ei):
    return torch.zeros_like(x).index_add_(0, ei[1], x[ei[0]]) + (1 + eps) * x


def _oracle_backward(g, eps, ei):
    return torch.zeros_like(g).index_add_(0, ei[0], g[ei[1]]) + (1 + eps) * g


def test_checker_accepts_the_oracle_and_catches_a_wrong_kernel():
    dev = torch.device('cpu')
    check_aggregation_properties(_oracle_forward, _oracle_backward, dev, n=3000, e=40000, f=16)
    # a kernel that sums a row's neighbours in a different order fails the bit-for-bit row check
    def reordered(x, eps, ei):
        perm = torch.argsort(ei[0], stable=True)
        return torch.zeros_like(x).index_add_(0, ei[1][perm], x[ei[0][perm]]) + (1 + eps) * x
    with pytest.raises(AssertionError):
        check_aggregation_properties(reordered, _oracle_backward, dev, n=3000, e=40000, f=16)
    # ... and a kernel that drops the self term fails the all-ones identity
    with pytest.raises(AssertionError):
        check_aggregation_properties(lambda x, eps, ei: torch.zeros_like(x).index_add_(0, ei[1], x[ei[0]]),
                                     _oracle_backward, dev, n=500, e=4000, f=8)
