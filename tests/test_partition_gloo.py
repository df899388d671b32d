"""Host-side logic of the node-partitioned multi-GPU path (gnnb200/partition.py) on CPU with the gloo
backend, world size 2: shard bounds, padded all-gather of row shards, the partition algebra of the
aggregation forward and its transposed backward (the CUDA gather is stood in for by the oracle's CPU
scatter — only the decomposition + collectives are under test here), Chan merge of BatchNorm moments,
and the flat gradient all-reduce.  The real kernels over NCCL are covered by tests/test_gpu_partition.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gnnb200  # noqa: F401
from gnnb200 import partition


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _oracle_agg(x_full, rowptr, col, self_x, eps):
    """CPU stand-in for gnnb200_aggregate_f32 (sequential edge-order sums over a CSR)."""
    n = rowptr.numel() - 1
    out = torch.zeros(n, x_full.size(1))
    deg = (rowptr[1:] - rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n), deg)
    out.index_add_(0, rows, x_full[col.long()])
    return out + (1 + eps) * self_x


def _csr(keys, others, n):
    perm = torch.sort(keys, stable=True).indices
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(keys, minlength=n).cumsum(0)])
    return rowptr, others[perm]


def _worker(rank, world, port, n, e, f, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        ei = torch.randint(0, n, (2, e), generator=g)
        x = torch.randn(n, f, generator=g)
        gout = torch.randn(n, f, generator=g)
        eps = torch.tensor([0.3])
        lo, hi, per = partition.shard_bounds(n, rank, world)

        # ---- single-graph answer (every rank computes it for comparison) ----
        xr = x.clone().requires_grad_(True)
        z = torch.zeros(n, f).index_add_(0, ei[1], xr[ei[0]]) + (1 + eps) * xr
        z.backward(gout)

        # ---- partitioned: forward over own dst rows, backward over own src rows ----
        shell = partition.PartitionedGraph.__new__(partition.PartitionedGraph)
        shell.num_nodes, shell.rank, shell.world, shell.group = n, rank, world, None
        shell.lo, shell.hi, shell.per, shell.n_local = lo, hi, per, hi - lo
        shell.chunks = 3
        shell.rpc = (per + shell.chunks - 1) // shell.chunks
        own_dst = (ei[1] >= lo) & (ei[1] < hi)
        rp, col = _csr(ei[1][own_dst] - lo, ei[0][own_dst], hi - lo)
        own_src = (ei[0] >= lo) & (ei[0] < hi)
        rpt, colt = _csr(ei[0][own_src] - lo, ei[1][own_src], hi - lo)
        h_full = shell.all_gather_rows(x[lo:hi])
        assert torch.equal(h_full, x)
        z_loc = _oracle_agg(h_full, rp, col, x[lo:hi], eps)
        torch.testing.assert_close(z_loc, z.detach()[lo:hi], rtol=1e-6, atol=1e-6)
        g_full = shell.all_gather_rows(gout[lo:hi])
        assert torch.equal(g_full, gout)
        gh_loc = _oracle_agg(g_full, rpt, colt, gout[lo:hi], eps)
        torch.testing.assert_close(gh_loc, xr.grad[lo:hi], rtol=1e-5, atol=1e-5)

        # ---- pipelined exchange: piece buffers + composite-key sub-CSRs reproduce the same rows ----
        pieces = shell.gather_pieces_async(x[lo:hi])
        own = (ei[1] >= lo) & (ei[1] < hi)
        o, m = ei[0][own], ei[1][own] - lo
        owner = o // per
        off = o - owner * per
        piece = off // shell.rpc
        col_in_piece = owner * shell.rpc + (off - piece * shell.rpc)
        acc = torch.zeros(hi - lo, f)
        for c, (work, buf) in enumerate(pieces):
            work.wait()
            sel = piece == c
            acc.index_add_(0, m[sel], buf[col_in_piece[sel]])
        acc = acc + (1 + eps) * x[lo:hi]
        torch.testing.assert_close(acc, z.detach()[lo:hi], rtol=1e-5, atol=1e-5)

        # ---- BatchNorm moments: Chan merge over ranks == moments of the whole activation ----
        a = torch.randn(n, 8, generator=g) * 3 + 10
        loc = a[lo:hi]
        s, m2 = loc.sum(0), ((loc - loc.mean(0)) ** 2).sum(0)
        tn, ts, tm2 = partition.merge_moments(partition.gather_moments(hi - lo, s, m2, None))
        assert torch.equal(tn, torch.full((8,), float(n)))
        torch.testing.assert_close(ts / tn, a.mean(0), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(tm2 / tn, a.var(0, unbiased=False), rtol=1e-4, atol=1e-5)

        # ---- flat gradient all-reduce ----
        lin = torch.nn.Linear(4, 3)
        with torch.no_grad():
            for p in lin.parameters():
                p.fill_(1.0)
        lin.weight.grad = torch.full_like(lin.weight, float(rank + 1))
        lin.bias.grad = torch.full_like(lin.bias, 10.0 * (rank + 1))
        partition.allreduce_gradients(lin)
        assert torch.equal(lin.weight.grad, torch.full_like(lin.weight, float(sum(range(1, world + 1)))))
        assert torch.equal(lin.bias.grad, torch.full_like(lin.bias, 10.0 * sum(range(1, world + 1))))

        # ---- replicas disagree on WHICH parameters carry a gradient (gradient surgery writes only the first shuffled
        #      task's parameters, reference gradient_surgery.py:60-68): same wire layout on every rank, union of owners ----
        net = torch.nn.ModuleDict({'a': torch.nn.Linear(3, 2), 'b': torch.nn.Linear(2, 2), 'c': torch.nn.Linear(2, 1)})
        owners = {'a': (0, 1), 'b': (0,), 'c': ()}               # a: both ranks, b: rank 0 only, c: nobody
        for name, mod in net.items():
            for p in mod.parameters():
                p.grad = torch.full_like(p, float(rank + 1)) if rank in owners[name] else None
        partition.allreduce_gradients(net)
        for p in net['a'].parameters():
            assert torch.equal(p.grad, torch.full_like(p, 3.0))
        for p in net['b'].parameters():
            assert torch.equal(p.grad, torch.full_like(p, 1.0))    # rank 1 had none: receives rank 0's
        for p in net['c'].parameters():
            assert p.grad is None                                  # nobody had one: the optimizer keeps skipping it
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def _sparse_worker(rank, world, port, n, e, f, locality, out_dir):
    """Sparse halo exchange (partition.HaloPlan): only referenced remote rows travel, delivered behind the rank's own rows;
    the renumbered CSR over that buffer gives the same rows — bit for bit — as the CSR over the all-gathered matrix."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    partition._pack_rows = lambda x, idx: x.index_select(0, idx)        # CPU stand-in for the rows_gather kernel
    try:
        from gnnb200 import synthetic
        d = synthetic.products_like(n, e, f, seed=7, locality=locality, blocks=8)
        ei, x = d['edge_index'], d['x']
        gout = torch.randn(n, f, generator=torch.Generator().manual_seed(1))
        eps = torch.tensor([0.3])
        lo, hi, per = partition.shard_bounds(n, rank, world)
        n_local = hi - lo
        for mine_row, other_row, feat in ((1, 0, x), (0, 1, gout)):        # forward (own dst), transposed (own src)
            mine, other = ei[mine_row], ei[other_row]
            own = (mine >= lo) & (mine < hi)
            o, m = other[own], mine[own] - lo
            plan = partition.HaloPlan(o, lo, hi, per, world)
            # -- the plan: needed ids are exactly the distinct remote endpoints; every peer serves what was asked of it
            remote = (o < lo) | (o >= hi)
            want = torch.unique(o[remote])
            assert plan.halo_rows == want.numel() and sum(plan.need_cnt) == want.numel() and plan.need_cnt[rank] == 0
            assert plan.serve_cnt[rank] == 0 and plan.serve_idx.numel() == sum(plan.serve_cnt)
            assert bool(((plan.serve_idx >= 0) & (plan.serve_idx < n_local)).all())
            assert bool(((plan.col >= 0) & (plan.col < n_local + plan.halo_rows)).all())
            # -- the exchange delivers own rows, then the needed remote rows in ascending global id
            buf = plan.exchange(feat[lo:hi])
            assert torch.equal(buf[:n_local], feat[lo:hi]) and torch.equal(buf[n_local:], feat[want])
            assert torch.equal(buf[plan.col], feat[o])                   # every owned edge finds its endpoint's row
            # -- same CSR order over the renumbered columns => identical sums to the dense (all-gather) exchange
            rp, col_sparse = _csr(m, plan.col, max(n_local, 1))
            _, col_dense = _csr(m, o, max(n_local, 1))
            z_sparse = _oracle_agg(buf, rp, col_sparse, feat[lo:hi], eps)
            z_dense = _oracle_agg(feat, rp, col_dense, feat[lo:hi], eps)
            assert torch.equal(z_sparse, z_dense)
            # -- 'sparse_overlap': asynchronous exchange of the halo rows alone; local-source edges summed first (over the rank's
            #    own rows), remote-source edges continue the sum over the halo buffer: the same rows to fp32 rounding
            work, halo = plan.exchange_async(feat[lo:hi])
            work.wait()
            assert torch.equal(halo, feat[want])
            far = plan.col >= n_local
            rp_l, col_l = _csr(m[~far], plan.col[~far], max(n_local, 1))
            rp_h, col_h = _csr(m[far], plan.col[far] - n_local, max(n_local, 1))
            z_split = _oracle_agg(feat[lo:hi], rp_l, col_l, feat[lo:hi], eps)
            if plan.halo_rows:
                z_split = z_split + _oracle_agg(halo, rp_h, col_h, torch.zeros_like(z_split), torch.tensor([-1.0]))
            torch.testing.assert_close(z_split, z_dense, rtol=1e-5, atol=1e-5)
            # -- the 'auto' criterion is the same number on every rank
            frac = partition.remote_fraction_needed(o, lo, hi, n)
            every = [None] * world
            dist.all_gather_object(every, frac)
            assert len(set(every)) == 1 and 0.0 <= frac <= 1.0
            if locality >= 0.9 and world > 1:
                assert frac < 0.9                                        # a graph with locality needs a strict subset
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


class _FakePeerDevice:
    """CPU emulation of the five peer entry points (include/gnnb200.h) over POSIX shared memory, so that the host
    logic of the 'peer' halo — handle exchange, pointer tables, two-buffer rotation, column encoding — runs for real
    across gloo processes.  "Device pointers" are ids into a per-process registry of mapped segments."""

    def __init__(self):
        from multiprocessing import shared_memory
        self.shm, self.seg, self.owned = shared_memory, {}, []

    @staticmethod
    def _floats(ptr, count):
        import ctypes
        import numpy as np
        return np.ctypeslib.as_array((ctypes.c_float * count).from_address(ptr))

    @staticmethod
    def _ints(ptr, count, ctype):
        import numpy as np
        return np.ctypeslib.as_array((ctype * count).from_address(ptr))

    def invoke(self, name, *a):
        import ctypes
        import numpy as np
        if name == 'gnnb200_peer_alloc':
            nbytes, ptr_ref, handle = a
            seg = self.shm.SharedMemory(create=True, size=nbytes)
            self.owned.append(seg)
            tag = seg.name.encode()
            assert len(tag) < 64
            for i in range(64):
                handle[i] = tag[i] if i < len(tag) else 0
            key = 1000 + len(self.seg)
            self.seg[key] = seg
            ptr_ref._obj.value = key
            return 0
        if name == 'gnnb200_peer_open':
            raw, ptr_ref = a
            seg = self.shm.SharedMemory(name=bytes(raw).rstrip(b'\0').decode())
            key = 1000 + len(self.seg)
            self.seg[key] = seg
            ptr_ref._obj.value = key
            return 0
        if name == 'gnnb200_peer_publish_f32':
            src, lds, rows, feat, dst, ldd, _ = a
            view = np.frombuffer(self.seg[dst].buf, dtype=np.float32)
            view[: rows * ldd].reshape(rows, ldd)[:, :feat] = self._floats(src, rows * lds).reshape(rows, lds)[:, :feat]
            return 0
        if name == 'gnnb200_aggregate_peer_f32':
            table, peers, ldx, rowptr, col, n_rows, feat, self_x, lds, eps, out, ldo, _ = a
            bases = self._ints(table, peers, ctypes.c_int64)
            rp = self._ints(rowptr, n_rows + 1, ctypes.c_int32)
            cols = self._ints(col, int(rp[-1]), ctypes.c_int32).astype(np.int64) & 0xffffffff if rp[-1] else np.zeros(0, np.int64)
            bufs = [np.frombuffer(self.seg[int(b)].buf, dtype=np.float32) for b in bases]
            rows_of = [torch.from_numpy(b[: b.size // ldx * ldx].reshape(-1, ldx)[:, :feat].copy()) for b in bufs]
            gathered = torch.stack([rows_of[int(c) >> 28][int(c) & ((1 << 28) - 1)] for c in cols]) if cols.size else torch.zeros(0, feat)
            acc = torch.zeros(n_rows, feat)
            acc.index_add_(0, torch.repeat_interleave(torch.arange(n_rows), torch.from_numpy(np.diff(rp)).long()), gathered)
            if self_x:
                mine = torch.from_numpy(self._floats(self_x, n_rows * lds).reshape(n_rows, lds)[:, :feat].copy())
                acc = acc + (1 + float(self._floats(eps, 1)[0]) if eps else 1.0) * mine
            self._floats(out, n_rows * ldo).reshape(n_rows, ldo)[:, :feat] = acc.numpy()
            return 0
        if name == 'gnnb200_peer_copy_f32':
            dst, src, count, _ = a
            self._floats(dst, count)[:] = np.frombuffer(self.seg[src].buf, dtype=np.float32)[:count]
            return 0
        if name == 'gnnb200_aggregate_f32':                      # the ordinary single-buffer gather ('peercopy' ends in it)
            x, ldx, rowptr, col, n_rows, feat, mode, self_x, lds, eps, dinv, out, ldo, _ = a
            accumulate = bool(mode & 8)                          # GNNB200_AGG_ACCUMULATE ('sparse_pull' / 'sparse_overlap' halo pass)
            assert mode & 7 == 0 and not dinv
            rp = self._ints(rowptr, n_rows + 1, ctypes.c_int32)
            cols = self._ints(col, int(rp[-1]), ctypes.c_int32).astype(np.int64) if rp[-1] else np.zeros(0, np.int64)
            rows_x = int(cols.max()) + 1 if cols.size else 0
            xs = torch.from_numpy(self._floats(x, max(rows_x, 1) * ldx).reshape(-1, ldx)[:, :feat].copy())
            acc = torch.zeros(n_rows, feat)
            if accumulate:
                acc = torch.from_numpy(self._floats(out, n_rows * ldo).reshape(n_rows, ldo)[:, :feat].copy())
            acc.index_add_(0, torch.repeat_interleave(torch.arange(n_rows), torch.from_numpy(np.diff(rp)).long()),
                           xs[torch.from_numpy(cols)] if cols.size else torch.zeros(0, feat))
            if self_x:
                mine = torch.from_numpy(self._floats(self_x, n_rows * lds).reshape(n_rows, lds)[:, :feat].copy())
                acc = acc + (1 + float(self._floats(eps, 1)[0])) * mine
            self._floats(out, n_rows * ldo).reshape(n_rows, ldo)[:, :feat] = acc.numpy()
            return 0
        if name in ('gnnb200_peer_close', 'gnnb200_peer_free'):
            return 0
        raise AssertionError(name)

    def release(self):
        for seg in self.seg.values():
            seg.close()
        for seg in self.owned:
            seg.unlink()


def _peer_worker(rank, world, port, n, e, f, out_dir, halo='peer'):
    """halo='peer' / 'peercopy' end to end on the host: PartitionedGraph(halo='peer').aggregate -> PeerRows (alloc, handle
    all-gather, open, tables) -> publish + barrier -> encoded-column gather; five passes per direction so that both
    buffers are reused.  Expected: the rank's rows of the full-graph answer."""
    from gnnb200 import ops
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    fake = _FakePeerDevice()
    ops._invoke, ops.on_device, ops._stream = fake.invoke, (lambda t: True), (lambda t: 0)

    def csr_build(pairs, n_rows, by_src):
        rp, col = _csr(pairs[1], pairs[0], n_rows)
        return rp.to(torch.int32), col.to(torch.int32), None
    ops.csr_build = csr_build
    try:
        g = torch.Generator().manual_seed(3)
        ei = torch.randint(0, n, (2, e), generator=g)
        eps = torch.tensor([0.3])
        graph = partition.PartitionedGraph(ei, n, rank, world, halo=halo)
        assert graph.halo == (halo if world > 1 else 'dense')
        lo, hi = graph.lo, graph.hi
        for it in range(5):
            x = torch.randn(n, f, generator=g)
            fwd = torch.zeros(n, f).index_add_(0, ei[1], x[ei[0]]) + (1 + eps) * x
            bwd = torch.zeros(n, f).index_add_(0, ei[0], x[ei[1]]) + (1 + eps) * x
            for transposed, want in ((False, fwd), (True, bwd)):
                got = graph.aggregate(x[lo:hi].contiguous(), eps, transposed)
                assert torch.allclose(got, want[lo:hi], rtol=0, atol=1e-5), (it, transposed)
        rows = partition.PeerRows.get(graph.per, f, rank, world, None, x.device)
        assert rows.turn == 10 and tuple(rows.tables.shape) == (2, world)
        dist.barrier()
        open(os.path.join(out_dir, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.barrier()
        fake.release()
        dist.destroy_process_group()


@pytest.mark.parametrize('halo', ['peer', 'peercopy', 'sparse_pull'])
@pytest.mark.parametrize('world,n,e', [(2, 101, 700), (3, 50, 400), (2, 7, 5)])
def test_peer_halo_host_logic(tmp_path, world, n, e, halo):
    mp.spawn(_peer_worker, args=(world, _free_port(), n, e, 8, str(tmp_path), halo), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))


def test_encode_peer_columns():
    ids = torch.tensor([0, 5, 99, 100, 101, 799])
    code = partition.encode_peer_columns(ids, 100)
    assert code.tolist() == [0, 5, 99, 1 << 28, (1 << 28) | 1, (7 << 28) | 99]
    assert int(code.max()) < 2 ** 31                                     # 8 ranks fit the int32 CSR column


@pytest.mark.parametrize('world,n,e,locality', [(2, 101, 700, 0.0), (3, 101, 400, 0.95), (2, 7, 5, 0.0), (3, 2, 6, 0.0)])
def test_sparse_halo_plan_and_exchange(tmp_path, world, n, e, locality):
    mp.spawn(_sparse_worker, args=(world, _free_port(), n, e, 8, locality, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))


@pytest.mark.parametrize('n,e', [(101, 700), (64, 300)])
def test_partition_algebra_world2(tmp_path, n, e):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, e, 16, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))


def test_shard_bounds_cover_everything():
    for n in (1, 7, 100, 2_449_029):
        for world in (1, 2, 4, 8):
            spans = [partition.shard_bounds(n, r, world)[:2] for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
