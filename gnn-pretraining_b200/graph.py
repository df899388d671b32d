"""Sorted-CSR view of an ``edge_index`` tensor, built on the device once and reused by every layer
and every pass that sees the same tensor (the reference rebuilds nothing because it has no CSR;
SURVEY.md §8a row a4)."""
import os
from typing import Optional

import torch
from torch import Tensor

from . import _lib as L
from . import ops

_ATTR = '_gnnb200_graph'

# Graphs with at least LONG_ROW_MIN_EDGES edges are checked for rows with more than L.AGG_LONG_ROW neighbours (one nonzero()
# = one device sync per CSR build; batches of small graphs never pay it); those rows then take the block-per-row kernel
# instead of one warp's serial walk.  Measured on B200, C5 size with the power-law generator (skew 1.8, ~150 hubs, largest
# 17 k neighbours): 12.56 -> 9.95 ms per aggregation pass, 254.6 -> 237.0 ms per step (profiles/r02/a_variants.md).
# GNNB200_LONG_ROWS=0 turns the split off.
LONG_ROWS = os.environ.get('GNNB200_LONG_ROWS', '1') == '1'
LONG_ROW_MIN_EDGES = 1 << 16


def long_rows_of(rowptr: Tensor, num_edges: int) -> Optional[Tensor]:
    """int64 ids of the rows longer than L.AGG_LONG_ROW, or None (none, small graph, or the feature is off)."""
    if not LONG_ROWS or num_edges < LONG_ROW_MIN_EDGES:
        return None
    ids = torch.nonzero((rowptr[1:] - rowptr[:-1]) > L.AGG_LONG_ROW).view(-1)
    return ids if ids.numel() else None


# GNNB200_CHECK_INDICES=1: validate node ids at CSR build (one device sync per build).  The kernels dereference `col`
# without bounds checks, like a raw CUDA gather; the PyTorch path they replace raises an index error instead.
CHECK_INDICES = os.environ.get('GNNB200_CHECK_INDICES', '0') == '1'


class Graph:
    """rowptr/col grouped by destination (forward gather) and, lazily, rowptr_t/col_t grouped by
    source (backward gather).  Within a row, neighbours keep the original edge order, which makes
    the fp32 sums bit-identical to the CPU scatter_add_ order."""

    def __init__(self, edge_index: Tensor, num_nodes: int):
        # a second tensor object over the same storage: the cache attribute lives on the caller's object, so holding that
        # one here would close a reference cycle and keep the CSR buffers alive until the cyclic GC runs
        self.edge_index = edge_index.detach()
        self.num_nodes = int(num_nodes)
        self.num_edges = int(edge_index.size(1))
        if CHECK_INDICES and self.num_edges:
            lo, hi = int(edge_index.min()), int(edge_index.max())
            if lo < 0 or hi >= self.num_nodes:
                raise IndexError(f'edge_index holds node ids in [{lo}, {hi}] for a graph of {self.num_nodes} nodes')
        self.rowptr, self.col, _ = ops.csr_build(edge_index, self.num_nodes, False)   # the edge permutation is not kept
        self.long_rows = long_rows_of(self.rowptr, self.num_edges)
        self._t = None
        self._version = edge_index._version

    def transposed(self):
        if self._t is None:
            rowptr_t, col_t, _ = ops.csr_build(self.edge_index, self.num_nodes, True)
            self._t = (rowptr_t, col_t, long_rows_of(rowptr_t, self.num_edges))
        return self._t

    @property
    def rowptr_t(self) -> Tensor:
        return self.transposed()[0]

    @property
    def col_t(self) -> Tensor:
        return self.transposed()[1]

    @property
    def long_rows_t(self) -> Optional[Tensor]:
        return self.transposed()[2]

    def in_degree(self) -> Tensor:
        return (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)

    def out_degree(self) -> Tensor:
        r = self.rowptr_t
        return (r[1:] - r[:-1]).to(torch.int64)


def graph_of(edge_index: Tensor, num_nodes: int) -> Graph:
    """Cached Graph for this edge_index tensor object (invalidated by in-place edits)."""
    g: Optional[Graph] = getattr(edge_index, _ATTR, None)
    if g is not None and g.num_nodes == num_nodes and g._version == edge_index._version \
            and g.num_edges == edge_index.size(1):
        return g
    g = Graph(edge_index, num_nodes)
    try:
        setattr(edge_index, _ATTR, g)
    except AttributeError:
        pass
    return g


def segment_ptr_of(batch: Tensor, size: Optional[int] = None) -> Tensor:
    """int32 ptr [B+1] for a sorted ``batch`` vector, cached on the tensor object per segment count.  When ``size`` is
    not given it is read from the last element (one device->host sync, like PyG's ``int(batch.max()) + 1``) and
    remembered, so the sync happens once per tensor."""
    cached = getattr(batch, '_gnnb200_ptr', None)
    if cached is None or cached['version'] != batch._version:
        cached = {'version': batch._version, 'natural': None, 'ptr': {}}
    if size is None:
        if cached['natural'] is None:
            cached['natural'] = int(batch[-1]) + 1 if batch.numel() > 0 else 0
        size = cached['natural']
    ptr = cached['ptr'].get(size)
    if ptr is None:
        ptr = cached['ptr'][size] = ops.segment_ptr(batch, size)
        try:
            setattr(batch, '_gnnb200_ptr', cached)
        except AttributeError:
            pass
    return ptr
