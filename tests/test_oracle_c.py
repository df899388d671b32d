"""The plain-C restatement (oracle/c/oracle_c.c) against the PyTorch shim (oracle/pyg_shim) — two independent
restatements of the same App. A semantics must agree BIT FOR BIT on the integer and edge-order arithmetic: CSR build,
coalesce / to_undirected, segment pointers, GIN aggregation forward and transposed, sum / mean / max pooling and the amax
backward tie rule, the LP decoder input, the deterministic negative-sampling branch, row gather / fill.  Edge cases:
empty edge lists, isolated nodes, empty graphs in a batch, duplicate edges and self loops, exact ties at 0.0."""
import random

import pytest
import torch

from oracle import c_oracle, install_pyg_shim

install_pyg_shim()
from torch_geometric.nn import global_max_pool, global_mean_pool  # noqa: E402
from torch_geometric.utils import batched_negative_sampling, coalesce, scatter, to_undirected  # noqa: E402


def _graph(n, e, seed):
    return torch.randint(0, n, (2, e), generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize('n,e', [(1, 0), (7, 0), (5, 40), (300, 2000), (2708, 10556)])
@pytest.mark.parametrize('by_src', [False, True])
def test_csr_build(n, e, by_src):
    ei = _graph(n, e, n + e)
    key, other = (ei[0], ei[1]) if by_src else (ei[1], ei[0])
    perm = torch.sort(key, stable=True).indices                                   # the contract of SURVEY §8a
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(key, minlength=n).cumsum(0)])
    got_ptr, got_col, got_eid = c_oracle.csr_build(ei, n, by_src)
    assert torch.equal(got_ptr.long(), rowptr) and torch.equal(got_col.long(), other[perm]) and torch.equal(got_eid.long(), perm)


def test_segment_ptr_with_empty_graphs():
    batch = torch.tensor([0, 0, 2, 2, 2, 5])
    assert c_oracle.segment_ptr(batch, 7).tolist() == [0, 2, 2, 5, 5, 5, 6, 6]
    assert c_oracle.segment_ptr(torch.empty(0, dtype=torch.long), 2).tolist() == [0, 0, 0]


@pytest.mark.parametrize('n,e', [(4, 0), (6, 30), (200, 1500)])
def test_coalesce_and_to_undirected(n, e):
    ei = _graph(n, e, 3 * n + e)
    assert torch.equal(c_oracle.coalesce(ei, n), coalesce(ei, n) if e else ei)
    assert torch.equal(c_oracle.coalesce(ei, n, symmetrise=True), to_undirected(ei, num_nodes=n) if e else ei)


@pytest.mark.parametrize('n,e,f', [(1, 0, 8), (10, 0, 8), (64, 300, 256), (300, 2000, 100), (50, 5000, 21), (2708, 10556, 64)])
def test_gin_aggregate_forward_and_transposed(n, e, f):
    ei = _graph(n, e, n + e + f)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    eps = torch.tensor([0.37])
    fwd = scatter(x.index_select(0, ei[0]), ei[1], dim=0, dim_size=n, reduce='sum') + (1 + eps) * x
    bwd = scatter(x.index_select(0, ei[1]), ei[0], dim=0, dim_size=n, reduce='sum') + (1 + eps) * x
    assert torch.equal(c_oracle.gin_aggregate(x, ei, 0.37), fwd)
    assert torch.equal(c_oracle.gin_aggregate(x, ei, 0.37, transposed=True), bwd)
    assert torch.equal(c_oracle.gin_aggregate(x, ei, 0.0, with_self=False),
                       scatter(x.index_select(0, ei[0]), ei[1], dim=0, dim_size=n, reduce='sum'))


def _batch(sizes):
    return torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))


@pytest.mark.parametrize('sizes', [[5], [3, 0, 4, 1], [30, 2, 17, 0, 0, 9], [0, 0, 3]])
def test_pools_and_amax_backward_rule(sizes):
    batch = _batch(sizes)
    n, B, f = int(batch.numel()), len(sizes), 24
    g = torch.Generator().manual_seed(n)
    x = torch.relu(torch.randn(n, f, generator=g))                # ReLU outputs: exact ties at 0.0, as on the real path
    x[torch.rand(n, f, generator=g) < 0.1] = 1.5                   # and ties at a positive maximum
    assert torch.equal(c_oracle.segment_pool(x, batch, B, 'mean'), global_mean_pool(x, batch, size=B))
    assert torch.equal(c_oracle.segment_pool(x, batch, B, 'sum'), scatter(x, batch, dim=0, dim_size=B, reduce='sum'))
    xr = x.clone().requires_grad_(True)
    out = global_max_pool(xr, batch, size=B)
    assert torch.equal(c_oracle.segment_pool(x, batch, B, 'max'), out.detach())
    gout = torch.randn(B, f, generator=g)
    out.backward(gout)
    assert torch.equal(c_oracle.segment_max_bwd(gout, x, out.detach(), batch), xr.grad)
    neg = -torch.rand(n, f, generator=g) - 0.1                     # all-negative rows: the zero init must not win the max
    assert torch.equal(c_oracle.segment_pool(neg, batch, B, 'max'), global_max_pool(neg, batch, size=B))


def test_lp_features_and_rows():
    g = torch.Generator().manual_seed(0)
    h = torch.randn(40, 32, generator=g)
    edges = torch.randint(0, 40, (2, 300), generator=g)
    hu, hv = h[edges[0]], h[edges[1]]
    assert torch.equal(c_oracle.lp_features(h, edges), torch.cat([hu + hv, hu * hv, (hu - hv).abs()], dim=-1))   # heads.py:59-65
    idx = torch.tensor([3, 3, 39, 0])
    assert torch.equal(c_oracle.rows_gather(h, idx), h[idx])
    token = torch.randn(32, generator=g)
    want = h.clone()
    want[idx] = token                                              # pretrain_model.py:84-86
    assert torch.equal(c_oracle.rows_fill(h, idx, token), want)


@pytest.mark.parametrize('sizes', [[4, 7, 2, 12], [30, 25], [2, 2, 3]])
def test_deterministic_negative_sampling_branch(sizes):
    """At the reference's call site (tasks.py:107-111) the per-graph quota is the whole batch's edge count, so small graphs
    return ALL their non-edges in ascending code order and consume no random number (App. A.5)."""
    g = torch.Generator().manual_seed(sum(sizes))
    pieces, start = [], 0
    for n in sizes:
        e = torch.randint(0, n, (2, max(1, n)), generator=g)
        pieces.append(to_undirected(e, num_nodes=n) + start)
        start += n
    ei = torch.cat(pieces, dim=1)
    batch = _batch(sizes)
    state = random.getstate()
    quota = 10000                                                   # 1.1 * quota exceeds every graph's n(n-1) population
    want = batched_negative_sampling(ei, batch, quota)
    assert random.getstate() == state                               # the deterministic branch drew nothing
    got, start, at = [], 0, 0
    for n, piece in zip(sizes, pieces):
        got.append(c_oracle.all_non_edges(piece - start, n, quota) + start)
        start += n
    assert torch.equal(torch.cat(got, dim=1), want)


def test_agrees_with_the_products_host_helpers():
    """The product's host-side mirrors (gnnb200.utils) against the C restatement."""
    import numpy as np
    import gnnb200  # noqa: F401
    from gnnb200 import utils
    ei = _graph(50, 400, 9)
    und = utils.to_undirected_host(ei.numpy(), 50)
    assert torch.equal(torch.from_numpy(und), c_oracle.coalesce(ei, 50, symmetrise=True))
    sizes = np.array([6, 9, 3])
    pieces, start = [], 0
    for n in sizes:
        pieces.append(to_undirected(torch.randint(0, int(n), (2, 5), generator=torch.Generator().manual_seed(int(n))), num_nodes=int(n)) + start)
        start += int(n)
    ei = torch.cat(pieces, dim=1)
    neg = utils.batched_negative_sampling_host(ei.numpy(), sizes, 1000)
    want, start = [], 0
    for n, piece in zip(sizes, pieces):
        want.append(c_oracle.all_non_edges(piece - start, int(n), 1000) + start)
        start += int(n)
    assert torch.equal(torch.from_numpy(neg), torch.cat(want, dim=1))
