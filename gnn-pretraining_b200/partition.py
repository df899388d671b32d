"""Node-partitioned execution of the backbone on one large graph over N GPUs (BASELINE config 5;
SURVEY.md §5.8b / §8e — a capability the reference does not have: it is single-device, full-graph).

Rank r owns the contiguous destination-row range [lo_r, hi_r).  Per layer:
  forward : h_full = all_gather(h_local)                  (NCCL over NVLink; every row crosses once)
            z_local = CSR_r gather over h_full + (1+eps) h_local      (rows of this rank only);
            the exchange is cut into pieces and piece c+1 is in flight while piece c is being gathered
            dense transforms on local rows; BatchNorm statistics reduced over all ranks (Chan merge of the
            per-rank (n, sum, m2) in rank order -> every rank holds bit-identical statistics)
  backward: g_full = all_gather(g_local);  dh_local = CSC_r gather over g_full + (1+eps) g_local
            (the transposed partition: no reduce-scatter of float partial sums, the order stays fixed)
            BatchNorm dgamma/dbeta all-reduced before the dx pass; d(eps) and the weight gradients are
            per-rank partial sums added by one flat all-reduce before the optimizer step.
With world size 1 every collective is skipped and the result equals the single-device path.
"""
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib as L
from . import nn as gnn
from . import ops


# Measured on 8 x B200, C5 uniform-random graph: 1 piece 70.1 ms/step, 4 pieces 72.1 ms/step — cutting the gather
# into pieces costs more (shorter rows per pass, accumulator re-reads, NCCL sharing SMs/HBM) than the overlap wins
# back, because every remote row is needed and the exchange, not the gather, is the long pole.  Pipelining pays when
# the local gather is longer than the exchange (few ranks, high-degree graphs); it stays opt-in.
DEFAULT_HALO_CHUNKS = 1
# The dense exchange is pipelined over COLUMN slabs instead: the [rows, F] shard is all-gathered as `slabs` matrices of F/slabs
# columns (all started asynchronously) and slab c is gathered while slabs c+1.. are still in flight.  Unlike the row pieces
# every pass walks whole neighbour lists and writes its own columns of the output once (no accumulator re-reads), and a
# column's sum keeps its edge order: bit-identical to the unsplit exchange.  Measured on 2 B200s (uniform graph): 4 slabs 164.8
# ms per step against 151.2 ms unsplit — four gathers over 256-byte row segments take 8.6 ms per pass where one gather over 1 KB
# rows takes 5.2 ms, more than the overlap returns — so it is opt-in (GNNB200_HALO_SLABS=2 / 4).
DEFAULT_HALO_SLABS = 1

# Which rows travel per layer and direction ("halo = all remote rows for the uniform generator; only
# referenced blocks for the locality generator"):
#   'dense'  : all-gather of every rank's whole shard (each row crosses NVLink once, no index traffic, no packing);
#              right when nearly every remote row is referenced — on the uniform-random C5 graph a rank touches
#              ~96 % of all rows at 8 ranks.  The measured default.
#   'sparse' : each rank receives only the remote rows its edges reference: the needed ids are exchanged once per
#              graph (HaloPlan), then per layer the owners pack those rows (rows_gather kernel) and one
#              all_to_all_single with uneven splits delivers them behind the rank's own rows in one buffer
#              [n_local + halo, F] that the local CSR indexes.  Same per-row edge order => same bits as 'dense'.
#   'auto'   : 'sparse' iff the largest needed fraction of remote rows over all ranks is below
#              SPARSE_HALO_MAX_FRACTION (one small all-reduce per graph, so every rank takes the same branch).
#   'peer'   : no exchange step at all: every rank publishes its rows in a peer-mapped buffer (CUDA IPC over NVLink,
#              PeerRows) and the gather kernel reads neighbour rows straight from the owning GPU
#              (gnnb200_aggregate_peer_f32) — transfer and sums overlap inside ONE kernel, nothing is packed, staged
#              or all-gathered; per pass one local copy of the shard and one barrier.  A remote row crosses NVLink once
#              per referencing edge, so this is for graphs with locality; same per-row edge order => same bits.
#   'peercopy': the dense exchange without NCCL: every rank publishes its shard (PeerRows) and PULLS the other shards with
#              the copy engines (cudaMemcpyAsync from the peer-mapped buffers, no SM time, no NCCL kernel sharing HBM with
#              anything), then runs the ordinary gather on the assembled [P * rows, F] buffer.  Same bytes as 'dense'.
#   'sparse_overlap': the sparse exchange hidden behind the part of the gather that does not need it: the owned edges are
#              split into those with a LOCAL source (CSR over this rank's own rows) and those with a REMOTE source (CSR
#              over the halo buffer); the halo all-to-all is started asynchronously, the local gather (+ self term) runs
#              while the rows travel, and the halo gather then continues every row's sum (ACCUMULATE).  A row's sum is
#              "local neighbours in edge order, then remote neighbours in edge order": deterministic, but a different
#              association than the single-device edge order (fp32 rounding instead of bit identity).
#   'sparse_pull': the same split as 'sparse_overlap', but the halo rows are not SENT: every rank publishes its shard in a
#              peer-mapped buffer (PeerRows: one local copy + one barrier per pass) and PULLS the rows it references with a
#              gather kernel whose "neighbour lists" have one remote row each (gnnb200_aggregate_peer_f32 over NVLink, one
#              warp per row, every load independent) on a side stream while the local-source edges are summed — no pack
#              kernel, no uneven all-to-all.  Same association as 'sparse_overlap' => the same bits as that mode.  Measured
#              slower (generator (ii), 8 GPUs: 67.1 vs 48.1 ms per step): SM-issued remote reads, as in 'peer'.  Opt-in.
# All modes are bit-identical to each other and to the single-device kernel except 'sparse_overlap' (measured on 2 and 8
# B200s over NCCL / CUDA IPC: tests/test_gpu_partition.py, bench.py's selfcheck).  Measured C5 steps (profiles/r02):
#   uniform graph,   8 GPUs: dense 71.4 ms, peercopy 155.0 ms (the seven pulls of a rank serialise on one stream)
#   90 % intra-block, 8 GPUs: sparse 50.9 ms, peer 83.0 ms (remote reads inside the gather are latency-bound: 5.6 ms / pass)
#   90 % intra-block, 2 GPUs: peer 136.9 ms, sparse 149.7 ms, dense 156.4 ms
#   90 % intra-block, 2 GPUs (72 % of the remote rows needed): sparse 136.0 ms, dense 151.2 ms
# hence 'auto' = the default: the overlapped sparse exchange when the ranks need less than 85 % of the remote rows, the
# all-gather otherwise (uniform graph: ~100 % at 2 ranks, 96 % at 8).
DEFAULT_HALO = 'auto'
SPARSE_HALO_MAX_FRACTION = 0.85
HALO_MODES = ('dense', 'sparse', 'sparse_overlap', 'sparse_pull', 'auto', 'peer', 'peercopy')


def shard_bounds(num_nodes: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(lo, hi, rows_per_rank) of the contiguous equal split; the last rank may own fewer rows."""
    per = (num_nodes + world - 1) // world
    lo = min(num_nodes, rank * per)
    hi = min(num_nodes, lo + per)
    return lo, hi, per


def _pack_rows(x_local: Tensor, idx: Tensor) -> Tensor:
    """Rows of this rank's shard that the peers asked for, in request order (the send buffer of the sparse exchange)."""
    return ops.rows_gather.fn(x_local, idx)


class HaloPlan:
    """One direction's sparse halo of one rank: which remote rows its owned edges reference (`need`, ascending global
    ids, hence grouped by owner), which of its own rows every peer references (`serve_idx`, local ids, grouped by
    requesting rank), and the owned edges' column endpoints renumbered into the exchange buffer
    [own rows 0..n_local) | halo rows n_local..n_local+H)`.  Built once per graph with two small all-to-alls
    (counts, then ids)."""

    def __init__(self, other: Tensor, lo: int, hi: int, per: int, world: int, group=None):
        self.n_local, self.group = hi - lo, group
        remote = (other < lo) | (other >= hi)
        need = torch.unique(other[remote])                              # sorted => grouped by owner rank
        owner = torch.div(need, per, rounding_mode='floor')
        need_cnt = torch.bincount(owner, minlength=world)
        serve_cnt = torch.empty_like(need_cnt)
        dist.all_to_all_single(serve_cnt, need_cnt, group=group)
        self.need_cnt, self.serve_cnt = need_cnt.tolist(), serve_cnt.tolist()    # one host read each, per graph
        self.serve_idx = need.new_empty(sum(self.serve_cnt))
        dist.all_to_all_single(self.serve_idx, need - owner * per, output_split_sizes=self.serve_cnt,
                               input_split_sizes=self.need_cnt, group=group)
        self.halo_rows = int(need.numel())
        self.need = need                                   # ascending global ids of the referenced remote rows
        self.col = torch.where(remote, self.n_local + torch.searchsorted(need, other), other - lo)

    def exchange_async(self, x_local: Tensor):
        """(work, [H, F] halo rows in ascending global id): the all-to-all is started with async_op; the caller waits on
        `work` before reading the rows."""
        f = x_local.size(1)
        halo = x_local.new_empty(self.halo_rows, f)
        send = _pack_rows(x_local, self.serve_idx)
        work = dist.all_to_all_single(halo, send, output_split_sizes=self.need_cnt, input_split_sizes=self.serve_cnt,
                                      group=self.group, async_op=True)
        return work, halo

    def exchange(self, x_local: Tensor) -> Tensor:
        """[n_local + H, F]: this rank's rows followed by the remote rows it references (ascending global id)."""
        f = x_local.size(1)
        buf = x_local.new_empty(self.n_local + self.halo_rows, f)
        buf[: self.n_local].copy_(x_local[: self.n_local])
        send = _pack_rows(x_local, self.serve_idx)
        dist.all_to_all_single(buf[self.n_local:], send, output_split_sizes=self.need_cnt,
                               input_split_sizes=self.serve_cnt, group=self.group)
        return buf


def remote_fraction_needed(other: Tensor, lo: int, hi: int, num_nodes: int, group=None) -> float:
    """max over ranks of (distinct remote rows referenced) / (remote rows): the 'auto' criterion."""
    remote = (other < lo) | (other >= hi)
    frac = torch.unique(other[remote]).numel() / max(1, num_nodes - (hi - lo))
    t = torch.tensor([frac], dtype=torch.float32, device=other.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t)


def encode_peer_columns(other: Tensor, per: int) -> Tensor:
    """Global row ids -> (owner slot << 28) | row inside the owner's published buffer (int64 holding the int32 code
    that gnnb200_aggregate_peer_f32 decodes).  Needs per <= 2^28 rows per rank and <= 8 ranks (code < 2^31)."""
    owner = torch.div(other, per, rounding_mode='floor')
    return (owner << L.PEER_SHIFT) | (other - owner * per)


class PeerRows:
    """This rank's two published row buffers ([rows, feat] fp32 each, allocated by gnnb200_peer_alloc) and the other
    ranks' buffers mapped into this device's address space (CUDA IPC handles exchanged once over the process group).
    `publish(x_local)` copies the shard into the next buffer, orders every rank behind it (one 4-byte all-reduce on
    the stream = the barrier) and returns the device table of the P base pointers the gather kernel indexes.
    Two buffers + one barrier per pass are enough: a rank can only overwrite buffer b two passes later, i.e. after a
    barrier that every rank entered behind its own read of b."""

    _cache = {}

    @classmethod
    def get(cls, rows: int, feat: int, rank: int, world: int, group, device) -> 'PeerRows':
        key = (group, rows, feat, rank, world, str(device))      # the group object itself: an id() could be reused after a destroy
        if key not in cls._cache:
            cls._cache[key] = cls(rows, feat, rank, world, group, device)
        return cls._cache[key]

    def __init__(self, rows: int, feat: int, rank: int, world: int, group, device):
        import ctypes
        if world > 8 or rows > (1 << L.PEER_SHIFT):
            raise L.Gnnb200Error('peer halo: at most 8 ranks and 2^28 rows per rank')
        self.rows, self.feat, self.rank, self.world, self.group = rows, feat, rank, world, group
        self.turn = 0
        self._mine, self._mapped, handles = [], [], []
        for _ in range(2):
            ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * L.PEER_HANDLE_BYTES)()
            L.check(ops._invoke('gnnb200_peer_alloc', rows * feat * 4, ctypes.byref(ptr), handle), 'peer_alloc')
            self._mine.append(ptr.value or 0)
            handles.append(bytes(handle))
        everyone = [None] * world
        dist.all_gather_object(everyone, handles, group=group)
        tables = []
        for b in range(2):
            ptrs = []
            for r in range(world):
                if r == rank:
                    ptrs.append(self._mine[b])
                    continue
                ptr = ctypes.c_void_p()
                raw = (ctypes.c_ubyte * L.PEER_HANDLE_BYTES).from_buffer_copy(everyone[r][b])
                L.check(ops._invoke('gnnb200_peer_open', raw, ctypes.byref(ptr)), f'peer_open (rank {r})')
                self._mapped.append(ptr.value or 0)
                ptrs.append(ptr.value or 0)
            tables.append(ptrs)
        self.ptrs = tables                                                         # the same pointers on the host
        self.tables = torch.tensor(tables, dtype=torch.int64).to(device)          # [2, P] base pointers
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)

    def gather_all(self, x_local: Tensor) -> Tensor:
        """[P * rows, F]: every rank's published shard pulled by the copy engines (the 'peercopy' exchange)."""
        self.publish(x_local)
        b = (self.turn - 1) & 1
        full = x_local.new_empty(self.world * self.rows, self.feat)
        count, stream = self.rows * self.feat, ops._stream(x_local)
        for r in range(self.world):
            L.check(ops._invoke('gnnb200_peer_copy_f32', full.data_ptr() + 4 * r * count, self.ptrs[b][r], count, stream),
                    f'peer_copy (rank {r})')
        return full

    def publish(self, x_local: Tensor) -> Tensor:
        b = self.turn & 1
        self.turn += 1
        x_local = ops._rowmajor(x_local)
        L.check(ops._invoke('gnnb200_peer_publish_f32', x_local.data_ptr(), ops._ld(x_local), x_local.size(0), self.feat,
                            self._mine[b], self.feat, ops._stream(x_local)), 'peer_publish')
        dist.all_reduce(self._flag, group=self.group)          # every rank's copy precedes every rank's gather
        return self.tables[b]

    def close(self) -> None:
        """Collective: every rank unmaps its peers' buffers, then (after a barrier — freeing exported memory that a peer
        still has mapped is undefined) frees its own."""
        for p in self._mapped:
            ops._invoke('gnnb200_peer_close', p)
        self._mapped = []
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        for p in self._mine:
            ops._invoke('gnnb200_peer_free', p)
        self._mine = []
        PeerRows._cache = {k: v for k, v in PeerRows._cache.items() if v is not self}


class PartitionedGraph:
    """This rank's slice of the graph.  Passed wherever the backbone expects ``edge_index``; GINConv
    recognises it and takes the partitioned aggregation.

    The halo exchange is pipelined: every rank's row shard is cut into ``chunks`` pieces; piece c of all
    ranks is all-gathered (NCCL, asynchronously) while the gather over piece c-1 runs.  For that the
    local CSR (rows = own destinations) and CSC (rows = own sources) are each built as ``chunks``
    sub-structures in one stable sort with the composite key ``chunk * n_local + local_row``; their
    column ids index the gathered piece buffer ``[world, rows_per_chunk, F]``.  A row's sum is continued
    from pass to pass in a fixed order (piece 0 edges in edge order, then piece 1, ...), so results are
    deterministic; they differ from the single-device edge order only by fp32 re-association."""

    def __init__(self, edge_index: Tensor, num_nodes: int, rank: int, world: int, group=None, chunks: Optional[int] = None,
                 halo: Optional[str] = None):
        if chunks is None:
            chunks = int(os.environ.get('GNNB200_HALO_CHUNKS', DEFAULT_HALO_CHUNKS))
        if halo is None:
            halo = os.environ.get('GNNB200_HALO', DEFAULT_HALO)
        if halo not in HALO_MODES:
            raise ValueError(f'halo must be one of {HALO_MODES}, not {halo!r}')
        self.num_nodes, self.rank, self.world, self.group = int(num_nodes), rank, world, group
        self.lo, self.hi, self.per = shard_bounds(self.num_nodes, rank, world)
        self.n_local = self.hi - self.lo
        self.chunks = max(1, min(int(chunks), self.per)) if world > 1 else 1
        self.slabs = max(1, int(os.environ.get('GNNB200_HALO_SLABS', DEFAULT_HALO_SLABS)))
        self.rpc = (self.per + self.chunks - 1) // self.chunks            # rows per piece of one shard
        n_rows = max(self.n_local, 1)
        src, dst = edge_index[0], edge_index[1]
        if halo == 'peercopy':
            self.chunks = 1                                               # one exchange, one gather pass
            self.rpc = self.per
        if world > 1 and halo == 'auto':
            own = (dst >= self.lo) & (dst < self.hi)
            frac = remote_fraction_needed(src[own], self.lo, self.hi, self.num_nodes, group)
            halo = 'sparse_overlap' if frac < SPARSE_HALO_MAX_FRACTION else 'dense'
        self.halo = halo if world > 1 else 'dense'
        self.plan = self.plan_t = None
        self.long_rows = self.long_rows_t = None
        if self.halo == 'sparse':
            self.chunks = 1
            self.rowptr, self.col, self.local_edges, self.plan = self._build_sparse(src, dst, n_rows)
            self.rowptr_t, self.col_t, _, self.plan_t = self._build_sparse(dst, src, n_rows)
            self._find_long_rows()
            return
        if self.halo in ('sparse_overlap', 'sparse_pull'):
            self.chunks = 1
            self.split, self.local_edges = self._build_split(src, dst, n_rows)
            self.split_t, _ = self._build_split(dst, src, n_rows)
            self.rowptr = self.col = self.rowptr_t = self.col_t = None
            self._pull_stream = None
            return
        if self.halo == 'peer':
            self.chunks = 1
            self.rowptr, self.col, self.local_edges = self._build_peer(src, dst, n_rows)
            self.rowptr_t, self.col_t, _ = self._build_peer(dst, src, n_rows)
            return
        self.rowptr, self.col, self.local_edges = self._build(src, dst, n_rows)       # own destinations
        self.rowptr_t, self.col_t, _ = self._build(dst, src, n_rows)                  # own sources (transposed)
        if self.chunks == 1:
            self._find_long_rows()

    def _find_long_rows(self) -> None:
        """Hub rows of this rank's CSRs (gnnb200.graph.long_rows_of): they take the block-per-row kernel, as on one device."""
        from .graph import long_rows_of
        self.long_rows = long_rows_of(self.rowptr, int(self.col.numel()))
        self.long_rows_t = long_rows_of(self.rowptr_t, int(self.col_t.numel()))

    def _build_sparse(self, other: Tensor, mine: Tensor, n_rows: int):
        """CSR over the owned edges whose columns index the sparse exchange buffer (HaloPlan)."""
        own = (mine >= self.lo) & (mine < self.hi)
        plan = HaloPlan(other[own], self.lo, self.hi, self.per, self.world, self.group)
        m = mine[own] - self.lo
        rowptr, col, _ = ops.csr_build(torch.stack([plan.col, m], dim=0), n_rows, False)
        plan.col = None                                                   # only the CSR copy is kept
        return rowptr, col, int(m.numel()), plan

    def _build_split(self, other: Tensor, mine: Tensor, n_rows: int):
        """(local CSR over x_local, halo CSR over the exchanged rows, plan) for the overlapped sparse exchange."""
        own = (mine >= self.lo) & (mine < self.hi)
        o, m = other[own], mine[own] - self.lo
        plan = HaloPlan(o, self.lo, self.hi, self.per, self.world, self.group)
        remote = plan.col >= self.n_local
        near = ~remote
        from .graph import long_rows_of
        rowptr_l, col_l, _ = ops.csr_build(torch.stack([plan.col[near], m[near]], dim=0), n_rows, False)
        rowptr_h, col_h, _ = ops.csr_build(torch.stack([plan.col[remote] - self.n_local, m[remote]], dim=0), n_rows, False)
        plan.col = None
        if self.halo == 'sparse_pull':                     # one-entry "neighbour lists": row i of the halo buffer = remote row need[i]
            plan.pull_rowptr = torch.arange(plan.halo_rows + 1, dtype=torch.int32, device=o.device)
            plan.pull_col = encode_peer_columns(plan.need, self.per).to(torch.int32)
        plan.need = None
        plan.long_local = long_rows_of(rowptr_l, int(col_l.numel()))
        plan.long_halo = long_rows_of(rowptr_h, int(col_h.numel()))
        return (rowptr_l, col_l, rowptr_h, col_h, plan), int(m.numel())

    def _build_peer(self, other: Tensor, mine: Tensor, n_rows: int):
        """CSR over the owned edges whose columns address the owners' published buffers (encode_peer_columns)."""
        own = (mine >= self.lo) & (mine < self.hi)
        m = mine[own] - self.lo
        rowptr, col, _ = ops.csr_build(torch.stack([encode_peer_columns(other[own], self.per), m], dim=0), n_rows, False)
        return rowptr, col, int(m.numel())

    def _build(self, other: Tensor, mine: Tensor, n_rows: int):
        """Sub-CSRs over the edges whose `mine` endpoint this rank owns; columns = `other` endpoints."""
        own = (mine >= self.lo) & (mine < self.hi)
        o, m = other[own], mine[own] - self.lo
        if self.world == 1:
            rowptr, col, _ = ops.csr_build(torch.stack([o, m], dim=0), n_rows, False)
            return rowptr, col, int(m.numel())
        owner = torch.div(o, self.per, rounding_mode='floor')
        off = o - owner * self.per
        piece = torch.div(off, self.rpc, rounding_mode='floor')
        col_in_piece = owner * self.rpc + (off - piece * self.rpc)                   # index into [world, rpc, F]
        key = piece * n_rows + m
        rowptr, col, _ = ops.csr_build(torch.stack([col_in_piece, key], dim=0), self.chunks * n_rows, False)
        return rowptr, col, int(m.numel())

    def sub_rowptr(self, rowptr: Tensor, piece: int) -> Tensor:
        n_rows = max(self.n_local, 1)
        return rowptr[piece * n_rows: (piece + 1) * n_rows + 1]

    def all_gather_rows(self, x_local: Tensor) -> Tensor:
        """[N, F] from every rank's [n_local, F] shard (equal-size padded shards on the wire)."""
        if self.world == 1:
            return x_local
        f = x_local.size(1)
        if x_local.size(0) == self.per:
            send = x_local.contiguous()
        else:
            send = x_local.new_zeros(self.per, f)
            send[: x_local.size(0)] = x_local
        full = x_local.new_empty(self.per * self.world, f)
        dist.all_gather_into_tensor(full, send, group=self.group)
        return full[: self.num_nodes]

    def gather_pieces_async(self, x_local: Tensor):
        """Start one asynchronous all-gather per piece; returns [(work, buffer [world*rpc, F]), ...]."""
        f = x_local.size(1)
        padded = x_local
        if x_local.size(0) != self.chunks * self.rpc:
            padded = x_local.new_zeros(self.chunks * self.rpc, f)
            padded[: x_local.size(0)] = x_local
        out = []
        for c in range(self.chunks):
            buf = x_local.new_empty(self.world * self.rpc, f)
            work = dist.all_gather_into_tensor(buf, padded[c * self.rpc: (c + 1) * self.rpc], group=self.group,
                                               async_op=True)
            out.append((work, buf))
        return out

    def aggregate(self, x_local: Tensor, eps: Tensor, transposed: bool) -> Tensor:
        """sum over this rank's rows of the (transposed) graph + (1+eps) * x_local, pipelined with the
        halo exchange."""
        rowptr, col = (self.rowptr_t, self.col_t) if transposed else (self.rowptr, self.col)
        if self.world == 1:
            return ops._aggregate_raw(x_local, rowptr, col, L.AGG_SUM, x_local, eps, None)
        long_rows = self.long_rows_t if transposed else self.long_rows
        if self.halo == 'sparse':
            buf = (self.plan_t if transposed else self.plan).exchange(x_local)
            return ops._aggregate_raw(buf, rowptr, col, L.AGG_SUM, x_local, eps, None, long_rows=long_rows)
        if self.halo == 'sparse_overlap':
            rowptr_l, col_l, rowptr_h, col_h, plan = self.split_t if transposed else self.split
            work, halo = plan.exchange_async(x_local)                   # rows travel ...
            out = ops._aggregate_raw(x_local, rowptr_l, col_l, L.AGG_SUM, x_local, eps, None,     # ... while these are summed
                                     long_rows=plan.long_local)
            if work is not None:
                work.wait()
            if plan.halo_rows:
                out = ops._aggregate_raw(halo, rowptr_h, col_h, L.AGG_SUM, None, None, None, out, long_rows=plan.long_halo)
            return out
        if self.halo == 'sparse_pull':
            rowptr_l, col_l, rowptr_h, col_h, plan = self.split_t if transposed else self.split
            f = x_local.size(1)
            table = PeerRows.get(self.per, f, self.rank, self.world, self.group, x_local.device).publish(x_local)
            halo = done = None
            if plan.halo_rows and x_local.is_cuda:
                main = torch.cuda.current_stream(x_local.device)
                if self._pull_stream is None:
                    self._pull_stream = torch.cuda.Stream(device=x_local.device)
                published = torch.cuda.Event()
                published.record(main)                                  # behind the publish copy and the barrier
                with torch.cuda.stream(self._pull_stream):
                    self._pull_stream.wait_event(published)
                    halo = ops.aggregate_peer(table, f, plan.pull_rowptr, plan.pull_col, f, None, None)   # rows over NVLink ...
                    done = torch.cuda.Event()
                    done.record(self._pull_stream)
            elif plan.halo_rows:
                halo = ops.aggregate_peer(table, f, plan.pull_rowptr, plan.pull_col, f, None, None)
            out = ops._aggregate_raw(x_local, rowptr_l, col_l, L.AGG_SUM, x_local, eps, None,             # ... while these are summed
                                     long_rows=plan.long_local)
            if halo is not None:
                if done is not None:
                    main.wait_event(done)
                    halo.record_stream(main)
                out = ops._aggregate_raw(halo, rowptr_h, col_h, L.AGG_SUM, None, None, None, out, long_rows=plan.long_halo)
            return out
        if self.halo == 'peer':
            f = x_local.size(1)
            table = PeerRows.get(self.per, f, self.rank, self.world, self.group, x_local.device).publish(x_local)
            return ops.aggregate_peer(table, f, rowptr, col, f, x_local, eps)
        if self.halo == 'peercopy':
            f = x_local.size(1)
            full = PeerRows.get(self.per, f, self.rank, self.world, self.group, x_local.device).gather_all(x_local)
            return ops._aggregate_raw(full, rowptr, col, L.AGG_SUM, x_local, eps, None, long_rows=long_rows)
        f = x_local.size(1)
        if self.chunks == 1 and self.slabs > 1 and f % (4 * self.slabs) == 0 and ops.on_device(x_local):
            return self._aggregate_column_slabs(x_local, eps, rowptr, col, long_rows)
        pieces = self.gather_pieces_async(x_local)
        out = None
        for c, (work, buf) in enumerate(pieces):
            work.wait()                                   # current stream waits for piece c only
            last = c == self.chunks - 1
            out = ops._aggregate_raw(buf, self.sub_rowptr(rowptr, c), col, L.AGG_SUM,
                                     x_local if last else None, eps if last else None, None, out,
                                     long_rows=long_rows if self.chunks == 1 else None)
        return out


def _aggregate_column_slabs(self, x_local: Tensor, eps: Tensor, rowptr: Tensor, col: Tensor, long_rows) -> Tensor:
    """Dense halo exchange pipelined over column slabs (see DEFAULT_HALO_SLABS)."""
    f, n_loc = x_local.size(1), x_local.size(0)
    w = f // self.slabs
    out = x_local.new_empty(n_loc, f)
    in_flight = []
    for c in range(self.slabs):
        send = x_local.new_empty(self.per, w)
        send[:n_loc].copy_(x_local[:, c * w:(c + 1) * w])
        if n_loc < self.per:
            send[n_loc:].zero_()
        buf = x_local.new_empty(self.world * self.per, w)
        in_flight.append((dist.all_gather_into_tensor(buf, send, group=self.group, async_op=True), buf))
    for c, (work, buf) in enumerate(in_flight):
        work.wait()                                       # the current stream waits for slab c only
        ops._aggregate_raw(buf, rowptr, col, L.AGG_SUM, x_local[:, c * w:(c + 1) * w], eps, None,
                           dst=out[:, c * w:(c + 1) * w], long_rows=long_rows)
    return out


PartitionedGraph._aggregate_column_slabs = _aggregate_column_slabs


class _PartitionedGINAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h_local: Tensor, eps: Tensor, graph: PartitionedGraph) -> Tensor:
        h_local = h_local.contiguous()
        ctx.graph = graph
        ctx.save_for_backward(h_local, eps)
        return graph.aggregate(h_local, eps, transposed=False)

    @staticmethod
    def backward(ctx, g_local: Tensor):
        graph = ctx.graph
        h_local, eps = ctx.saved_tensors
        g_local = g_local.contiguous()
        gh = graph.aggregate(g_local, eps, transposed=True)
        geps = ops.dot(g_local, h_local) if ctx.needs_input_grad[1] else None      # per-rank partial sum
        return gh, geps, None


def partitioned_gin_aggregate(h_local: Tensor, eps: Tensor, graph: PartitionedGraph) -> Tensor:
    return _PartitionedGINAggregate.apply(h_local, eps, graph)


def merge_moments(per_rank: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Chan's parallel-variance merge of per-rank moments [P, 3, C] = (n, sum, centred m2) ->
    (total n, total sum, total m2), each [C].  Same inputs in the same rank order on every rank -> same bits."""
    cnt, sums, m2s = per_rank[:, 0], per_rank[:, 1], per_rank[:, 2]
    total_n = cnt.sum(0)
    total_sum = sums.sum(0)
    mean = total_sum / total_n.clamp(min=1.0)
    local_mean = sums / cnt.clamp(min=1.0)
    m2_tot = m2s.sum(0) + (cnt * (local_mean - mean) ** 2).sum(0)
    return total_n, total_sum, m2_tot


def gather_moments(n_rows: int, s: Tensor, m2: Tensor, group) -> Tensor:
    """All-gather this rank's (n, sum, m2) -> [P, 3, C]."""
    mine = torch.stack([torch.full_like(s, float(n_rows)), s, m2])
    world = dist.get_world_size(group)
    flat = mine.new_empty(world * 3, s.numel())            # concatenation along dim 0 (gloo and nccl agree on this form)
    dist.all_gather_into_tensor(flat, mine, group=group)
    return flat.view(world, 3, s.numel())


def synced_batch_stats(x_local: Tensor, graph_rows_total: int, group, running_mean: Optional[Tensor],
                       running_var: Optional[Tensor], momentum: float, eps: float, local_stats=None
                       ) -> Tuple[Tensor, Tensor]:
    """Global (mean, invstd) of a row-partitioned activation: per-rank (n, sum, m2) gathered and merged with
    Chan's formula in rank order; running buffers updated from the global moments (torch's rule)."""
    cols = x_local.size(1)
    if local_stats is None and ops.on_device(x_local) and x_local.is_cuda:
        # three launches + one collective: the column statistics land in rows 1-2 of the send buffer, row 0 is the row
        # count; after the all-gather one kernel merges the ranks' moments and finalises (gnnb200_bn_merge_finalize_f32)
        world = dist.get_world_size(group)
        x2 = ops._rowmajor(x_local.detach())
        mine = torch.empty(3, cols, dtype=torch.float32, device=x_local.device)
        mine[0].fill_(float(x_local.size(0)))
        ops._call_ws('gnnb200_colstats_f32', 'bn colstats', x2.device, x2.data_ptr(), ops._ld(x2), x2.size(0), cols,
                     mine[1].data_ptr(), mine[2].data_ptr(), stream=ops._stream(x2), key=(x2.size(0), cols))
        everyone = torch.empty(world * 3, cols, dtype=torch.float32, device=x_local.device)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        mean_out = torch.empty(cols, dtype=torch.float32, device=x_local.device)
        invstd = torch.empty(cols, dtype=torch.float32, device=x_local.device)
        L.check(ops._invoke('gnnb200_bn_merge_finalize_f32', everyone.data_ptr(), world, cols, eps, momentum,
                            ops._ptr(running_mean), ops._ptr(running_var), mean_out.data_ptr(), invstd.data_ptr(),
                            ops._stream(x2)), 'bn_merge_finalize')
        return mean_out, invstd
    s, m2 = local_stats if local_stats is not None else ops.colstats(x_local.detach())
    _, gsum, m2_tot = merge_moments(gather_moments(x_local.size(0), s, m2, group))
    mean_out = torch.empty_like(s)
    invstd = torch.empty_like(s)
    L.check(ops._invoke('gnnb200_bn_finalize_f32', gsum.data_ptr(), m2_tot.data_ptr(), graph_rows_total, s.numel(),
                        eps, momentum, ops._ptr(running_mean), ops._ptr(running_var), mean_out.data_ptr(),
                        invstd.data_ptr(), ops._stream(s)), 'bn_finalize (synced)')
    return mean_out, invstd


class partition_scope:
    """Context manager: inside it BatchNormAct layers treat their input as this rank's row shard."""

    def __init__(self, graph: PartitionedGraph):
        self.graph = graph

    def __enter__(self):
        self._old = (gnn._partition, ops.SYNC_GROUP)
        gnn._partition = self.graph if self.graph.world > 1 else None
        ops.SYNC_GROUP = (self.graph.group or dist.group.WORLD) if self.graph.world > 1 else None
        return self.graph

    def __exit__(self, *exc):
        gnn._partition, ops.SYNC_GROUP = self._old


def allreduce_gradients(module: torch.nn.Module, group=None) -> None:
    """One flat all-reduce (sum) over every parameter gradient (C1 of SURVEY §2.4).

    Every rank contributes a buffer with the SAME layout — all parameters in `module.parameters()` order, zeros
    where this rank has no gradient — because which parameters carry a gradient can differ between replicas (the
    reference's gradient surgery only writes the parameters of the first task of an unseeded shuffle,
    src/pretrain/gradient_surgery.py:43,60-68).  A parameter ends up with a gradient iff some rank had one (its
    presence flags travel in the same buffer); the others keep `.grad = None`, so the optimizer skips them exactly as
    it would on one device."""
    params = list(module.parameters())
    if not params:
        return
    dev = params[0].device
    pieces = [(p.grad.reshape(-1) if p.grad is not None else torch.zeros(p.numel(), dtype=torch.float32, device=dev))
              for p in params]
    flags = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], dtype=torch.float32, device=dev)
    flat = torch.cat(pieces + [flags])
    dist.all_reduce(flat, group=group)
    have = (flat[-len(params):] > 0).tolist()                    # one small device->host read
    off = 0
    for p, present in zip(params, have):
        n = p.numel()
        if present:
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
        off += n


class PartitionedBackboneStep:
    """The bench workload's training step (InputEncoder + GINBackbone, loss = sum(h), AdamW) on this
    rank's shard.  Every rank builds the same model from the same seed."""

    def __init__(self, models, device, feat_in: int, hidden: int, layers: int, num_nodes: int, rank: int,
                 world: int, group=None, seed: int = 0, lr: float = 1e-4, halo: Optional[str] = None):
        torch.manual_seed(seed)
        self.model = torch.nn.ModuleDict({'input_encoder': models.InputEncoder(feat_in, hidden),
                                          'gnn_backbone': models.GINBackbone(layers, hidden)}).to(device)
        self.model.train()
        self.opt = torch.optim.AdamW(self.model.parameters(), lr=lr)
        self.num_nodes, self.rank, self.world, self.group = num_nodes, rank, world, group
        self.halo, self.last_halo = halo, None
        self._graph_key, self._graph = None, None

    def graph_of(self, edge_index: Tensor) -> PartitionedGraph:
        """This rank's partition of `edge_index`, built once per tensor object and version (like gnnb200.graph.graph_of on one
        device): a static graph pays the O(E) ownership scan, the sorts and the halo plan once, a new or overwritten edge
        list pays them again."""
        import weakref
        alive = self._graph_key[0]() if self._graph_key is not None else None
        if self._graph is None or alive is not edge_index or self._graph_key[1] != edge_index._version:
            self._graph = PartitionedGraph(edge_index, self.num_nodes, self.rank, self.world, self.group, halo=self.halo)
            self._graph_key = (weakref.ref(edge_index), edge_index._version)     # the same OBJECT, unmodified since
        return self._graph

    def step(self, x: Tensor, edge_index: Tensor, x_is_local: bool = False) -> Tensor:
        """x: the full [N, F] feature matrix, or (x_is_local) just this rank's row shard."""
        graph = self.graph_of(edge_index)
        self.last_halo = graph.halo
        x_local = x if x_is_local else x[graph.lo:graph.hi]
        self.opt.zero_grad(set_to_none=True)
        with partition_scope(graph):
            h = self.model['gnn_backbone'](self.model['input_encoder'](x_local), graph)
            loss = h.sum()
            loss.backward()
        if self.world > 1:
            allreduce_gradients(self.model, self.group)
        self.opt.step()
        return loss
