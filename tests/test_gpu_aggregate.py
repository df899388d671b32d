"""Neighbour aggregation fwd/bwd against the oracle (CPU scatter_add_ in edge order): bit-exact
for the sum / (1+eps) self term / transposed backward; 1e-5 for the eps gradient (tree reduce)."""
import os

import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import _lib as L
from gnnb200 import ops, synthetic
from gnnb200.graph import Graph
from gnnb200.nn import GINConv
from oracle import install_pyg_shim

install_pyg_shim()
from torch_geometric.nn import GINConv as OracleGINConv  # noqa: E402
from torch_geometric.utils import scatter  # noqa: E402

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


def _graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (2, e), generator=g)


@pytest.mark.parametrize('n,e,f', [(1, 0, 256), (10, 0, 256), (64, 300, 256), (2708, 10556, 256), (500, 4000, 512),
                                   (300, 2000, 128), (300, 2000, 64), (300, 2000, 16), (300, 2000, 100),
                                   (300, 2000, 21), (200, 1500, 7), (100, 900, 1028), (50, 5000, 256)])
def test_gin_aggregate_forward_bit_exact(n, e, f):
    ei = _graph(n, e, n + e + f)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    eps = torch.tensor([0.37])
    want = scatter(x.index_select(0, ei[0]), ei[1], dim=0, dim_size=n, reduce='sum') + (1 + eps) * x
    gr = Graph(ei.to(DEV), n)
    got = ops.gin_aggregate(x.to(DEV), eps.to(DEV), gr.rowptr, gr.col, gr.rowptr.new_empty(0), gr.col.new_empty(0))
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize('n,e,f', [(10, 0, 256), (64, 300, 256), (2708, 10556, 256), (300, 2000, 100), (50, 5000, 256)])
def test_csr_and_aggregation_against_the_c_oracle(n, e, f):
    """The same inputs as above against the SECOND oracle (plain C, oracle/c/oracle_c.c): CSR / CSC arrays and the forward and
    transposed sums, bit for bit."""
    from oracle import c_oracle
    try:
        c_oracle.lib()
    except Exception as exc:                                  # noqa: BLE001 — no C compiler on this box: nothing to compare with
        pytest.skip(f'C oracle unavailable: {exc}')
    ei = _graph(n, e, n + e + f)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
    eps = torch.tensor([0.37])
    gr = Graph(ei.to(DEV), n)
    for by_src, (rowptr, col) in ((False, (gr.rowptr, gr.col)), (True, (gr.rowptr_t, gr.col_t))):
        want_ptr, want_col, _ = c_oracle.csr_build(ei, n, by_src)
        assert torch.equal(rowptr.cpu(), want_ptr) and torch.equal(col.cpu(), want_col)
        got = ops._aggregate_raw(x.to(DEV), rowptr, col, L.AGG_SUM, x.to(DEV), eps.to(DEV), None)
        assert torch.equal(got.cpu(), c_oracle.gin_aggregate(x, ei, 0.37, transposed=by_src))


@pytest.mark.parametrize('n,e,f', [(64, 300, 256), (2708, 10556, 256), (300, 2000, 100), (200, 1500, 7)])
def test_gin_aggregate_backward(n, e, f):
    ei = _graph(n, e, 3 * n + e)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(2))
    gout = torch.randn(n, f, generator=torch.Generator().manual_seed(3))
    xo = x.clone().requires_grad_(True)
    eo = torch.tensor([0.21], requires_grad=True)
    zo = scatter(xo.index_select(0, ei[0]), ei[1], dim=0, dim_size=n, reduce='sum') + (1 + eo) * xo
    zo.backward(gout)
    gr = Graph(ei.to(DEV), n)
    xg = x.to(DEV).requires_grad_(True)
    eg = torch.tensor([0.21], device=DEV, requires_grad=True)
    z = ops.gin_aggregate(xg, eg, gr.rowptr, gr.col, gr.rowptr_t, gr.col_t)
    z.backward(gout.to(DEV))
    assert torch.equal(z.detach().cpu(), zo.detach())
    assert torch.equal(xg.grad.cpu(), xo.grad)                      # a+b == b+a: still bit-exact
    torch.testing.assert_close(eg.grad.cpu(), eo.grad, rtol=1e-5, atol=1e-5)


def test_ginconv_module_matches_oracle_module():
    n, e, h = 500, 3000, 256
    ei = _graph(n, e, 9)
    x = torch.randn(n, h, generator=torch.Generator().manual_seed(4))
    mk = lambda lin: torch.nn.Sequential(lin(h, 2 * h), torch.nn.ReLU(), lin(2 * h, h))
    from gnnb200.nn import Linear
    a = OracleGINConv(mk(torch.nn.Linear), train_eps=True)
    b = GINConv(mk(Linear), train_eps=True)
    b.load_state_dict(a.state_dict())
    b = b.to(DEV)
    ya = a(x, ei)
    yb = b(x.to(DEV), ei.to(DEV))
    assert float((ya - yb.cpu()).abs().max() / ya.abs().max()) < 1e-5    # fp32 class


@pytest.mark.parametrize('f', [256, 100, 30])
def test_mean_and_gcn_modes_formula_oracle(f):
    """Extra modes requested by the north star (not used by the reference): SURVEY.md App. A.7."""
    n, e = 400, 3000
    ei = _graph(n, e, 77)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(5))
    gr = Graph(ei.to(DEV), n)
    want_mean = scatter(x.index_select(0, ei[0]), ei[1], dim=0, dim_size=n, reduce='mean')
    got_mean = ops.aggregate(x.to(DEV), gr.rowptr, gr.col, L.AGG_MEAN)
    assert torch.equal(got_mean.cpu(), want_mean)
    # GCN symmetric norm with self loops appended after the existing edges
    deg = torch.bincount(ei[1], minlength=n).float() + 1.0
    dinv = deg.pow(-0.5)
    src = torch.cat([ei[0], torch.arange(n)])
    dst = torch.cat([ei[1], torch.arange(n)])
    w = dinv[src] * dinv[dst]
    want_gcn = scatter(w.view(-1, 1) * x.index_select(0, src), dst, dim=0, dim_size=n, reduce='sum')
    got_gcn = ops.aggregate(x.to(DEV), gr.rowptr, gr.col, L.AGG_GCN, x.to(DEV), None, dinv.to(DEV))
    assert torch.equal(got_gcn.cpu(), want_gcn)


def test_large_graph_properties():
    """Size-independent checks at a scale the CPU oracle would take long on: linearity in x and the
    all-ones identity (row sums = in-degree + 1 + eps)."""
    n, e, f = 200_000, 5_000_000, 256
    d = synthetic.products_like(n, e, 8, seed=1, device=DEV)
    gr = Graph(d['edge_index'], n)
    eps = torch.tensor([0.5], device=DEV)
    empty = gr.rowptr.new_empty(0)
    ones = torch.ones(n, f, device=DEV)
    z = ops.gin_aggregate(ones, eps, gr.rowptr, gr.col, empty, empty)
    want = (gr.in_degree().float() + 1.5).view(-1, 1).expand(n, f)
    assert torch.equal(z, want)
    a = torch.randn(n, f, device=DEV)
    b = torch.randn(n, f, device=DEV)
    za = ops.gin_aggregate(a, eps, gr.rowptr, gr.col, empty, empty)
    zb = ops.gin_aggregate(b, eps, gr.rowptr, gr.col, empty, empty)
    zab = ops.gin_aggregate(a + b, eps, gr.rowptr, gr.col, empty, empty)
    assert float((zab - (za + zb)).abs().max()) < 1e-3 * float(zab.abs().max())
    # determinism: same bits on a second run
    assert torch.equal(za, ops.gin_aggregate(a, eps, gr.rowptr, gr.col, empty, empty))


@pytest.mark.parametrize('n,e,f,world', [(1000, 9000, 256, 4), (257, 3000, 128, 8), (2708, 10556, 512, 2),
                                         (300, 2000, 64, 3), (64, 0, 256, 2), (999, 8000, 1024, 5)])
def test_peer_gather_equals_single_device(n, e, f, world):
    """gnnb200_aggregate_peer_f32 with `world` virtual ranks on ONE device: every rank's rows live in their own
    allocation (different base pointers), columns carry (owner slot, row) — each rank's output must equal its rows
    of the single-device kernel bit for bit, forward and transposed."""
    from gnnb200 import partition
    ei = _graph(n, e, n + 7 * e + f).to(DEV)
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(5)).to(DEV)
    eps = torch.tensor([0.37], device=DEV)
    gr = Graph(ei, n)
    want = ops.gin_aggregate(x, eps, gr.rowptr, gr.col, gr.rowptr.new_empty(0), gr.col.new_empty(0))
    per = (n + world - 1) // world
    shards = [torch.zeros(per, f, device=DEV) for _ in range(world)]          # the "published buffers"
    for r, buf in enumerate(shards):
        lo, hi = r * per, min(n, (r + 1) * per)
        if hi > lo:
            buf[: hi - lo] = x[lo:hi]
    table = torch.tensor([b.data_ptr() for b in shards], dtype=torch.int64, device=DEV)
    for r in range(world):
        lo, hi = r * per, min(n, (r + 1) * per)
        if hi <= lo:
            continue
        own = (ei[1] >= lo) & (ei[1] < hi)
        pairs = torch.stack([partition.encode_peer_columns(ei[0][own], per), ei[1][own] - lo])
        rowptr, col, _ = ops.csr_build(pairs, hi - lo, False)
        got = ops.aggregate_peer(table, f, rowptr, col, f, x[lo:hi], eps)
        assert torch.equal(got, want[lo:hi]), (r, world)
        no_self = ops.aggregate_peer(table, f, rowptr, col, f, None, None)
        assert torch.equal(no_self, ops.aggregate(x, gr.rowptr, gr.col, L.AGG_SUM)[lo:hi])


@pytest.mark.parametrize('f', [256, 64, 512, 1024])
def test_long_rows_take_the_block_per_row_kernel(monkeypatch, f):
    """A graph with three hubs (5,000 / 1,500 / 1,025 in-neighbours, one of them also a source hub) among ordinary rows:
    with GNNB200_LONG_ROWS the hubs go through gnnb200_aggregate_long_rows_f32 — equal to the edge-order CPU sum to fp32
    rounding — and every other row stays bit-identical to the single-kernel path; forward, transposed and accumulate."""
    from gnnb200 import graph as graph_mod
    n = 4000
    g = torch.Generator().manual_seed(f)
    base = torch.randint(0, n, (2, 30000), generator=g)
    hubs = [(7, 5000), (2500, 1500), (3999, 1025)]
    extra = [torch.stack([torch.randint(0, n, (k,), generator=g), torch.full((k,), h)]) for h, k in hubs]
    extra.append(torch.stack([torch.full((3000,), 7), torch.randint(0, n, (3000,), generator=g)]))      # node 7: source hub too
    ei = torch.cat([base] + extra, dim=1)
    ei = ei[:, torch.randperm(ei.size(1), generator=g)].contiguous()
    x = torch.randn(n, f, generator=g)
    eps = torch.tensor([0.37])
    want = scatter(x.index_select(0, ei[0]), ei[1], dim=0, dim_size=n, reduce='sum') + (1 + eps) * x
    want_t = scatter(x.index_select(0, ei[1]), ei[0], dim=0, dim_size=n, reduce='sum') + (1 + eps) * x
    plain = Graph(ei.to(DEV), n)
    assert plain.long_rows is None                                         # below LONG_ROW_MIN_EDGES: never checked
    monkeypatch.setattr(graph_mod, 'LONG_ROWS', True)
    monkeypatch.setattr(graph_mod, 'LONG_ROW_MIN_EDGES', 1)
    gr = Graph(ei.to(DEV).clone(), n)
    assert sorted(gr.long_rows.tolist()) == [7, 2500, 3999] and gr.long_rows_t.tolist() == [7]
    xd, ed = x.to(DEV), eps.to(DEV)
    for rowptr, col, long_rows, ref in ((gr.rowptr, gr.col, gr.long_rows, want), (gr.rowptr_t, gr.col_t, gr.long_rows_t, want_t)):
        single = ops._aggregate_raw(xd, rowptr, col, L.AGG_SUM, xd, ed, None)
        split = ops._aggregate_raw(xd, rowptr, col, L.AGG_SUM, xd, ed, None, long_rows=long_rows)
        assert torch.equal(single.cpu(), ref)                               # the sequential kernel is exact
        hub = torch.zeros(n, dtype=torch.bool)
        hub[long_rows.cpu()] = True
        assert torch.equal(split.cpu()[~hub], ref[~hub])
        err = (split.cpu()[hub] - ref[hub]).abs().max() / ref[hub].abs().max()
        assert 0 <= float(err) < 1e-5, float(err)
        again = ops._aggregate_raw(xd, rowptr, col, L.AGG_SUM, xd, ed, None, long_rows=long_rows)
        assert torch.equal(split, again)                                    # deterministic
        start = torch.randn(n, f, generator=g).to(DEV)                       # accumulate mode (fused backward: dh = ds + ...)
        acc_single = ops._aggregate_raw(xd, rowptr, col, L.AGG_SUM, xd, ed, None, out=start.clone())
        acc_split = ops._aggregate_raw(xd, rowptr, col, L.AGG_SUM, xd, ed, None, out=start.clone(), long_rows=long_rows)
        assert torch.equal(acc_split.cpu()[~hub], acc_single.cpu()[~hub])
        assert float((acc_split.cpu()[hub] - acc_single.cpu()[hub]).abs().max() / ref[hub].abs().max()) < 1e-5
