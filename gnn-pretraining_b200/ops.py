"""torch.library custom ops (namespace ``gnnb200``) over the C ABI of libgnnb200.so.

Each op takes CUDA tensors, hands raw device pointers + the current CUDA stream to one
``gnnb200_*`` entry point (include/gnnb200.h) and returns fresh tensors.  Autograd formulas are
registered here too; every backward is itself made of gnnb200 ops (the transposed-CSR gather for
the aggregation, the amax-rule scatter for max pooling, GEMMs for the linears).  There is no CPU
or eager fallback: CPU tensors raise.
"""
import ctypes
import os
from ctypes import byref, c_size_t
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = L.load()
    return _lib


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)
_current_device = getattr(torch._C, '_cuda_getDevice', None) or torch.cuda.current_device


def _stream(t: Tensor) -> int:
    """Raw cudaStream_t of torch's current stream on t's device (the C call is ~10x cheaper than building a
    torch.cuda.Stream object; it matters when a step is a few hundred tiny launches)."""
    index = t.device.index
    if index is not None and index != _current_device():
        # the C entry points launch on the CURRENT device: a tensor of another GPU would be dereferenced in the wrong context
        raise L.Gnnb200Error(f'tensor lives on cuda:{index} but the current device is cuda:{_current_device()}: '
                             f'call under torch.cuda.device({index}) (one process per GPU sets it once)')
    if _raw_stream is not None:
        return _raw_stream(index if index is not None else _current_device())
    return torch.cuda.current_stream(t.device).cuda_stream


def on_device(t: Tensor) -> bool:
    """True when `t` lives where the kernels run.  The single place the package asks that question, so that the
    host-overhead profiler (scripts/host_profile.py) can run the Python layers with the kernels stubbed out."""
    return t.is_cuda


def _need_cuda(*tensors: Optional[Tensor]) -> None:
    for t in tensors:
        if t is not None and not on_device(t):
            raise L.Gnnb200Error('gnnb200 ops run on CUDA tensors only (no CPU fallback)')


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _rowmajor(t: Tensor) -> Tensor:
    """2-D fp32 with unit inner stride (any leading dimension)."""
    if t.dtype != torch.float32:
        raise L.Gnnb200Error(f'expected float32, got {t.dtype}')
    if t.dim() != 2:
        raise L.Gnnb200Error(f'expected a 2-D tensor, got {t.dim()}-D')
    if t.size(1) > 1 and t.stride(1) != 1:
        return t.contiguous()
    if t.size(0) > 1 and t.stride(0) < t.size(1):      # expanded / overlapping rows
        return t.contiguous()
    return t


def _ld(t: Tensor) -> int:
    return t.stride(0) if t.size(0) > 1 else max(t.size(1), 1)


BYPASS_OPERAND_CACHES = False     # gnnb200.cuda_graphs: recompute splits / re-pitched copies inside a captured graph


def _tma_rows(t: Tensor) -> Tensor:
    """`t` with a row pitch TMA can address (a multiple of 16 bytes, 16-byte aligned base): the tensor itself when it
    already is, else the [rows, cols] view of a zero-padded copy with leading dimension ceil4(cols) — same logical shape,
    so the tensor-core GEMM covers feature widths such as 1433 / 3703 / 21 / 7 / 37 (src/data/data_setup.py:31-41) instead
    of dropping to the FFMA kernel.  The copy is cached on the tensor object until it is modified in place (dataset
    features and weights are re-used every step)."""
    if t.size(1) % 4 == 0 and _ld(t) % 4 == 0 and t.data_ptr() % 16 == 0:
        return t
    if _ld(t) % 4 == 0 and t.data_ptr() % 16 == 0:
        return t                                           # ragged width on an aligned pitch: TMA clips the box
    key = (t._version, t.data_ptr(), tuple(t.shape))
    cached = getattr(t, '_gnnb200_tma_rows', None)
    if cached is not None and cached[0] == key and not BYPASS_OPERAND_CACHES:
        return cached[1]
    src = t.detach()
    pad = torch.zeros(src.size(0), (src.size(1) + 3) // 4 * 4, dtype=src.dtype, device=src.device)
    view = pad[:, : src.size(1)]
    view.copy_(src)
    if not BYPASS_OPERAND_CACHES:                         # (a graph capture's buffers are only valid after a replay)
        try:
            t._gnnb200_tma_rows = (key, view)
        except AttributeError:
            pass
    return view


# Own-kernel launches per entry-point call (library kernels such as CUB's sort are not counted).
KERNELS_PER_CALL = {
    'gnnb200_csr_build_i64': 2, 'gnnb200_segment_ptr_i64': 1, 'gnnb200_coalesce_i64': 3,
    'gnnb200_aggregate_f32': 1, 'gnnb200_aggregate_long_rows_f32': 1, 'gnnb200_dot_f32': 2, 'gnnb200_segment_pool_fwd_f32': 1,
    'gnnb200_segment_pool_bwd_f32': 1, 'gnnb200_rows_gather_f32': 1, 'gnnb200_rows_scatter_f32': 1,
    'gnnb200_rows_gather_bwd_f32': 1, 'gnnb200_gemm_f32': 1, 'gnnb200_linear_x3w_f32': 1, 'gnnb200_split_tf32_f32': 1,
    'gnnb200_colstats_f32': 2,
    'gnnb200_bn_finalize_f32': 1, 'gnnb200_bn_merge_finalize_f32': 1, 'gnnb200_bn_act_fwd_f32': 1, 'gnnb200_bn_act_bwd_f32': 3,
    'gnnb200_act_dropout_fwd_f32': 1, 'gnnb200_act_dropout_bwd_f32': 1, 'gnnb200_scale_f32': 1, 'gnnb200_sqdiff_sum_f32': 1,
    'gnnb200_sqdiff_bwd_f32': 1, 'gnnb200_sigmoid_bce_fwd_f32': 1, 'gnnb200_sigmoid_bce_bwd_f32': 1, 'gnnb200_ce_sum_fwd_f32': 1,
    'gnnb200_ce_bwd_f32': 1, 'gnnb200_negsample_count_i64': 1, 'gnnb200_negsample_write_i64': 1,
    'gnnb200_lp_features_f32': 1, 'gnnb200_lp_features_bwd_f32': 1, 'gnnb200_ntxent_fwd_f32': 3,
    'gnnb200_ntxent_bwd_f32': 1, 'gnnb200_pcgrad_f32': 2, 'gnnb200_normalize_rows_f32': 1,
    'gnnb200_normalize_rows_bwd_f32': 1, 'gnnb200_ntxent_sim_fwd_f32': 2, 'gnnb200_ntxent_sim_bwd_f32': 1,
    'gnnb200_aggregate_peer_f32': 1, 'gnnb200_peer_publish_f32': 0, 'gnnb200_peer_copy_f32': 0, 'gnnb200_peer_alloc': 0, 'gnnb200_peer_open': 0,
    'gnnb200_peer_close': 0, 'gnnb200_peer_free': 0,
    # composites (csrc/gin_layer.cu): base sequence + optional parts counted under their own keys by gnnb200/fused.py
    'gnnb200_gin_layer_fwd_f32': 5, 'gnnb200_gin_layer_fwd_f32:stats': 6,
    'gnnb200_gin_layer_bwd_f32': 10, 'gnnb200_gin_layer_bwd_f32:deps': 2, 'gnnb200_gin_layer_bwd_f32:dh': 1,
}
_calls = {}
AGG_TIMER = None      # bench.py sets this to a list to collect (start, stop) CUDA events per aggregation launch


def reset_counters() -> None:
    _calls.clear()


def call_counts() -> dict:
    return dict(_calls)


def launch_count() -> int:
    """Kernels of this library launched since reset_counters() (lower bound: split-K finish passes
    and chunked-pool finish passes are not added)."""
    return sum(KERNELS_PER_CALL.get(k, 1) * v for k, v in _calls.items())


def _invoke(name: str, *args) -> int:
    _calls[name] = _calls.get(name, 0) + 1
    return getattr(lib(), name)(*args)


_ws_sizes = {}


def _call_ws(name: str, what: str, device, *args, stream: int, key=None):
    """Two-phase call of an entry point whose trailing args are (workspace, &bytes, stream).  `key` (the
    shape-determining arguments) lets the size query be answered from a cache."""
    fn = getattr(lib(), name)
    nbytes = _ws_sizes.get((name, key)) if key is not None else None
    if nbytes is None:
        need = c_size_t(0)
        L.check(fn(*args, None, byref(need), stream), what + ' (workspace query)')
        nbytes = max(int(need.value), 256)
        if key is not None:
            _ws_sizes[(name, key)] = nbytes
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    have = c_size_t(ws.numel())
    _calls[name] = _calls.get(name, 0) + 1
    L.check(fn(*args, ws.data_ptr(), byref(have), stream), what)
    return ws


# ---------------------------------------------------------------------------------------------
# registration: every op is defined in the torch.library namespace ``gnnb200`` (schema inferred from the
# annotations, CUDA implementation only -> CPU tensors raise, fake/meta kernel, autograd formula), so
# ``torch.ops.gnnb200.<name>`` is the public operator surface.  The modules of this package call the same
# Python implementation through a thin fast path (direct call, or one torch.autograd.Function when a
# gradient is needed): the dispatcher round trip of a Python custom op costs ~60-80 us per call, which is
# what bounds the small-graph configs (hundreds of ops per step), not the kernels.
# ---------------------------------------------------------------------------------------------
import inspect

_LIBRARY = torch.library.Library('gnnb200', 'DEF')


class _Op:
    def __init__(self, name: str, fn, mutates):
        self.name, self.fn, self.qualname = name, fn, f'gnnb200::{name}'
        self.__doc__ = fn.__doc__
        schema = torch.library.infer_schema(fn, mutates_args=mutates)
        _LIBRARY.define(name + schema)
        _LIBRARY.impl(name, fn, 'CUDA')
        sig = inspect.signature(fn)
        self._params = list(sig.parameters.values())
        self._names = [p.name for p in self._params]
        self._autograd = None

    def register_fake(self, fake):
        torch.library.register_fake(self.qualname)(fake)
        return fake

    def register_autograd(self, backward, setup_context):
        torch.library.register_autograd(self.qualname, backward, setup_context=setup_context)
        impl = self.fn

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *args):
                out = impl(*args)
                setup_context(ctx, args, out)
                return out

            @staticmethod
            def backward(ctx, *grads):
                return backward(ctx, *grads)

        _Fn.__name__ = f'gnnb200_{self.name}'
        self._autograd = _Fn

    def _bind(self, args, kwargs):
        if not kwargs and len(args) == len(self._params):
            return args
        full = list(args)
        for p in self._params[len(args):]:
            if p.name in kwargs:
                full.append(kwargs[p.name])
            elif p.default is not inspect.Parameter.empty:
                full.append(p.default)
            else:
                raise TypeError(f'{self.qualname}: missing argument {p.name!r}')
        return tuple(full)

    def __call__(self, *args, **kwargs):
        args = self._bind(args, kwargs)
        if self._autograd is not None and torch.is_grad_enabled():
            for a in args:
                if isinstance(a, Tensor) and a.requires_grad:
                    return self._autograd.apply(*args)
        return self.fn(*args)


def _op(name: str, mutates=()):
    def wrap(fn):
        return _Op(name, fn, mutates)
    return wrap


# ---------------------------------------------------------------------------------------------
# structure ops (integer, bit-exact)
# ---------------------------------------------------------------------------------------------
@_op('csr_build')
def csr_build(edge_index: Tensor, num_nodes: int, by_src: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """(rowptr int32 [N+1], col int32 [E], eid int32 [E]); see gnnb200_csr_build_i64."""
    _need_cuda(edge_index)
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise L.Gnnb200Error('edge_index must be int64 [2, E]')
    ei = edge_index.contiguous()
    E = ei.size(1)
    dev = ei.device
    rowptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev)
    eid = torch.empty(E, dtype=torch.int32, device=dev)
    _call_ws('gnnb200_csr_build_i64', 'csr_build', dev, _ptr(ei), E, num_nodes, int(by_src),
             _ptr(rowptr), _ptr(col), _ptr(eid), stream=_stream(ei))
    return rowptr, col, eid


@csr_build.register_fake
def _(edge_index, num_nodes, by_src):
    E = edge_index.size(1)
    mk = lambda n: edge_index.new_empty(n, dtype=torch.int32)
    return mk(num_nodes + 1), mk(E), mk(E)


@_op('segment_ptr')
def segment_ptr(ids: Tensor, num_segments: int) -> Tensor:
    """Offsets int32 [S+1] of a sorted int64 id vector (Batch.batch -> ptr)."""
    _need_cuda(ids)
    ids = ids.contiguous()
    out = torch.empty(num_segments + 1, dtype=torch.int32, device=ids.device)
    L.check(_invoke('gnnb200_segment_ptr_i64', _ptr(ids), ids.numel(), num_segments, _ptr(out), _stream(ids)),
            'segment_ptr')
    return out


@segment_ptr.register_fake
def _(ids, num_segments):
    return ids.new_empty(num_segments + 1, dtype=torch.int32)


@_op('coalesce')
def coalesce(edge_index: Tensor, num_nodes: int) -> Tuple[Tensor, Tensor]:
    """(out [2, E] capacity buffer, count int64 [1]): columns sorted by row*N+col, duplicates dropped."""
    _need_cuda(edge_index)
    ei = edge_index.contiguous()
    E = ei.size(1)
    out = torch.empty_like(ei)
    count = torch.zeros(1, dtype=torch.int64, device=ei.device)
    _call_ws('gnnb200_coalesce_i64', 'coalesce', ei.device, _ptr(ei), E, num_nodes, _ptr(out), _ptr(count),
             stream=_stream(ei))
    return out, count


@coalesce.register_fake
def _(edge_index, num_nodes):
    return torch.empty_like(edge_index), edge_index.new_empty(1)


# ---------------------------------------------------------------------------------------------
# aggregation
# ---------------------------------------------------------------------------------------------
def _aggregate_raw(x: Tensor, rowptr: Tensor, col: Tensor, mode: int, self_x: Optional[Tensor],
                   eps: Optional[Tensor], dinv: Optional[Tensor], out: Optional[Tensor] = None,
                   long_rows: Optional[Tensor] = None, dst: Optional[Tensor] = None) -> Tensor:
    """out given => accumulate into it (GNNB200_AGG_ACCUMULATE): a later pass of a chunked aggregation.
    dst given => write the result there (e.g. a column slab of a wider matrix: any leading dimension), no accumulation.
    long_rows (int64 ids of the rows with more than L.AGG_LONG_ROW neighbours, SUM mode): the main launch skips them and a
    block-per-row kernel covers them (gnnb200_aggregate_long_rows_f32)."""
    _need_cuda(x, rowptr, col, self_x, eps, dinv)
    x = _rowmajor(x)
    n_rows = rowptr.numel() - 1
    if dst is not None:
        if out is not None or dst.shape != (n_rows, x.size(1)) or dst.stride(1) != 1:
            raise L.Gnnb200Error('aggregate: dst must be a [rows, F] view with unit inner stride (and no `out`)')
        out = dst
    elif out is None:
        out = torch.empty(n_rows, x.size(1), dtype=torch.float32, device=x.device)
    else:
        mode |= L.AGG_ACCUMULATE
    if self_x is not None:
        self_x = _rowmajor(self_x)
    split_long = (long_rows is not None and (mode & 7) == L.AGG_SUM and x.size(1) % 4 == 0 and x.size(1) <= 1024
                  and _ld(x) % 4 == 0 and x.data_ptr() % 16 == 0 and out.data_ptr() % 16 == 0 and _ld(out) % 4 == 0
                  and (self_x is None or (_ld(self_x) % 4 == 0 and self_x.data_ptr() % 16 == 0)))
    if split_long:
        mode |= L.AGG_SKIP_LONG
    timer = AGG_TIMER
    if timer is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    L.check(_invoke('gnnb200_aggregate_f32',
        _ptr(x), _ld(x), _ptr(rowptr), _ptr(col), n_rows, x.size(1), mode,
        _ptr(self_x), _ld(self_x) if self_x is not None else 0, _ptr(eps), _ptr(dinv),
        _ptr(out), _ld(out), _stream(x)), 'aggregate')
    if split_long:
        L.check(_invoke('gnnb200_aggregate_long_rows_f32', _ptr(x), _ld(x), _ptr(rowptr), _ptr(col), _ptr(long_rows),
                        long_rows.numel(), x.size(1), mode & ~L.AGG_SKIP_LONG, _ptr(self_x),
                        _ld(self_x) if self_x is not None else 0, _ptr(eps), _ptr(out), _ld(out), _stream(x)),
                'aggregate (long rows)')
    if timer is not None:
        ev[1].record()
        timer.append(ev)
    return out


def aggregate_peer(table: Tensor, ld: int, rowptr: Tensor, col: Tensor, feat: int, self_x: Optional[Tensor],
                   eps: Optional[Tensor]) -> Tensor:
    """Aggregation whose neighbour rows live in peer-mapped buffers of several GPUs (gnnb200_aggregate_peer_f32):
    `table` int64 [P] on this device = base pointers of the published [rows, ld] fp32 buffers, `col` int32 with the
    owner slot in bits 31..28.  SUM + optional (1+eps) self term; no autograd (see partition._PartitionedGINAggregate)."""
    _need_cuda(table, rowptr, col, self_x, eps)
    n_rows = rowptr.numel() - 1
    out = torch.empty(n_rows, feat, dtype=torch.float32, device=table.device)
    if self_x is not None:
        self_x = _rowmajor(self_x)
    timer = AGG_TIMER
    if timer is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    L.check(_invoke('gnnb200_aggregate_peer_f32', _ptr(table), table.numel(), ld, _ptr(rowptr), _ptr(col), n_rows, feat,
                    _ptr(self_x), _ld(self_x) if self_x is not None else 0, _ptr(eps), _ptr(out), _ld(out),
                    _stream(table)), 'aggregate_peer')
    if timer is not None:
        ev[1].record()
        timer.append(ev)
    return out


@_op('aggregate')
def aggregate(x: Tensor, rowptr: Tensor, col: Tensor, mode: int, self_x: Optional[Tensor] = None,
              eps: Optional[Tensor] = None, dinv: Optional[Tensor] = None) -> Tensor:
    """Raw CSR aggregation (no autograd): see gnnb200_aggregate_f32."""
    return _aggregate_raw(x, rowptr, col, mode, self_x, eps, dinv)


@aggregate.register_fake
def _(x, rowptr, col, mode, self_x=None, eps=None, dinv=None):
    return x.new_empty(rowptr.numel() - 1, x.size(1))


@_op('dot')
def dot(a: Tensor, b: Tensor) -> Tensor:
    """Deterministic sum(a*b) -> [1]."""
    _need_cuda(a, b)
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty(1, dtype=torch.float32, device=a.device)
    _call_ws('gnnb200_dot_f32', 'dot', a.device, _ptr(a), _ptr(b), a.numel(), _ptr(out), stream=_stream(a), key=())
    return out


@dot.register_fake
def _(a, b):
    return a.new_empty(1)


@_op('gin_aggregate')
def gin_aggregate(x: Tensor, eps: Tensor, rowptr: Tensor, col: Tensor, rowptr_t: Tensor, col_t: Tensor) -> Tensor:
    """z[i] = sum_{e: dst=i} x[src_e] + (1+eps) * x[i]   (GINConv before its MLP).
    rowptr/col: CSR grouped by dst; rowptr_t/col_t: grouped by src (used by the backward)."""
    return _aggregate_raw(x, rowptr, col, L.AGG_SUM, x, eps, None)


@gin_aggregate.register_fake
def _(x, eps, rowptr, col, rowptr_t, col_t):
    return torch.empty_like(x)


def _gin_setup(ctx, inputs, output):
    x, eps, rowptr, col, rowptr_t, col_t = inputs
    ctx.save_for_backward(x, eps, rowptr, col, rowptr_t, col_t)


def _gin_backward(ctx, g):
    x, eps, rowptr, col, rowptr_t, col_t = ctx.saved_tensors
    g = g.contiguous()
    gx = geps = None
    if ctx.needs_input_grad[0]:
        if rowptr_t.numel() == 0:
            raise L.Gnnb200Error('gin_aggregate backward needs the by-src CSR (build the graph with grad enabled)')
        gx = gin_aggregate(g, eps, rowptr_t, col_t, rowptr, col)   # transposed gather + (1+eps) g
    if ctx.needs_input_grad[1]:
        geps = dot(g, x)                                          # d/d eps = sum(g * x)
    return gx, geps, None, None, None, None


gin_aggregate.register_autograd(_gin_backward, setup_context=_gin_setup)


# ---------------------------------------------------------------------------------------------
# pooling
# ---------------------------------------------------------------------------------------------
@_op('segment_pool')
def segment_pool(x: Tensor, ptr: Tensor, mode: int) -> Tensor:
    """Segment sum/mean/max over row ranges ptr (int32 [S+1]) -> [S, F]."""
    _need_cuda(x, ptr)
    x = _rowmajor(x)
    S = ptr.numel() - 1
    out = torch.empty(S, x.size(1), dtype=torch.float32, device=x.device)
    _call_ws('gnnb200_segment_pool_fwd_f32', 'segment_pool', x.device, _ptr(x), _ld(x), _ptr(ptr), x.size(0), S,
             x.size(1), mode, _ptr(out), _ld(out), stream=_stream(x), key=(x.size(0), S, x.size(1)))
    return out


@segment_pool.register_fake
def _(x, ptr, mode):
    return x.new_empty(ptr.numel() - 1, x.size(1))


@_op('segment_pool_bwd')
def segment_pool_bwd(grad_out: Tensor, x: Tensor, out: Tensor, ptr: Tensor, mode: int) -> Tensor:
    _need_cuda(grad_out, x, out, ptr)
    g, x, out = _rowmajor(grad_out), _rowmajor(x), _rowmajor(out)
    S = ptr.numel() - 1
    gx = torch.empty(x.size(0), x.size(1), dtype=torch.float32, device=x.device)
    L.check(_invoke('gnnb200_segment_pool_bwd_f32', 
        _ptr(g), _ld(g), _ptr(x), _ld(x), _ptr(out), _ld(out), _ptr(ptr), x.size(0), S, x.size(1), mode,
        _ptr(gx), _ld(gx), _stream(x)), 'segment_pool_bwd')
    return gx


@segment_pool_bwd.register_fake
def _(grad_out, x, out, ptr, mode):
    return torch.empty_like(x)


def _pool_setup(ctx, inputs, output):
    x, ptr, mode = inputs
    ctx.mode = mode
    ctx.save_for_backward(x, output, ptr)


def _pool_backward(ctx, g):
    x, out, ptr = ctx.saved_tensors
    return segment_pool_bwd(g.contiguous(), x, out, ptr, ctx.mode), None, None


segment_pool.register_autograd(_pool_backward, setup_context=_pool_setup)


# ---------------------------------------------------------------------------------------------
# row gather / scatter
# ---------------------------------------------------------------------------------------------
@_op('rows_gather')
def rows_gather(x: Tensor, idx: Tensor) -> Tensor:
    """out[i] = x[idx[i]] (idx int64)."""
    _need_cuda(x, idx)
    x = _rowmajor(x)
    idx = idx.contiguous()
    out = torch.empty(idx.numel(), x.size(1), dtype=torch.float32, device=x.device)
    L.check(_invoke('gnnb200_rows_gather_f32', _ptr(x), _ld(x), _ptr(idx), idx.numel(), x.size(1), _ptr(out), _ld(out),
                                         _stream(x)), 'rows_gather')
    return out


@rows_gather.register_fake
def _(x, idx):
    return x.new_empty(idx.numel(), x.size(1))


@_op('rows_gather_bwd')
def rows_gather_bwd(grad_out: Tensor, idx: Tensor, num_rows: int) -> Tensor:
    """Deterministic transpose of rows_gather: grad_x[r] = sum_{i: idx[i]==r} grad_out[i]."""
    _need_cuda(grad_out, idx)
    g = _rowmajor(grad_out)
    pairs = torch.stack([idx, torch.arange(idx.numel(), device=idx.device)], dim=0)
    rowptr, _, eid = csr_build(pairs, num_rows, True)
    gx = torch.empty(num_rows, g.size(1), dtype=torch.float32, device=g.device)
    L.check(_invoke('gnnb200_rows_gather_bwd_f32', _ptr(g), _ld(g), _ptr(rowptr), _ptr(eid), num_rows, g.size(1),
                                             _ptr(gx), _ld(gx), _stream(g)), 'rows_gather_bwd')
    return gx


@rows_gather_bwd.register_fake
def _(grad_out, idx, num_rows):
    return grad_out.new_empty(num_rows, grad_out.size(1))


def _rg_setup(ctx, inputs, output):
    x, idx = inputs
    ctx.n = x.size(0)
    ctx.save_for_backward(idx)


def _rg_backward(ctx, g):
    (idx,) = ctx.saved_tensors
    return rows_gather_bwd(g.contiguous(), idx, ctx.n), None


rows_gather.register_autograd(_rg_backward, setup_context=_rg_setup)


@_op('rows_scatter')
def rows_scatter(base: Tensor, src: Tensor, idx: Tensor) -> Tensor:
    """Copy of ``base`` with rows idx[i] replaced by src[i] (src [M,F]) or by the single row src [F]."""
    _need_cuda(base, src, idx)
    out = _rowmajor(base).clone()
    broadcast = src.dim() == 1
    s = src.contiguous().view(1, -1) if broadcast else _rowmajor(src)
    idx = idx.contiguous()
    L.check(_invoke('gnnb200_rows_scatter_f32', _ptr(s), _ld(s), int(broadcast), _ptr(idx), idx.numel(), out.size(1),
                                          _ptr(out), _ld(out), _stream(out)), 'rows_scatter')
    return out


@rows_scatter.register_fake
def _(base, src, idx):
    return torch.empty_like(base)


def _rs_setup(ctx, inputs, output):
    base, src, idx = inputs
    ctx.broadcast = src.dim() == 1
    ctx.save_for_backward(idx)


def _rs_backward(ctx, g):
    (idx,) = ctx.saved_tensors
    g = g.contiguous()
    gsel = rows_gather(g, idx)                                     # gradient reaching the written rows
    zero = torch.zeros(g.size(1), dtype=g.dtype, device=g.device)
    gbase = rows_scatter(g, zero, idx)                             # overwritten rows get no gradient
    gsrc = gsel.sum(dim=0) if ctx.broadcast else gsel
    return gbase, gsrc, None


rows_scatter.register_autograd(_rs_backward, setup_context=_rs_setup)


# ---------------------------------------------------------------------------------------------
# dense transforms
# ---------------------------------------------------------------------------------------------
# 'tf32' = tensor cores wherever the layout rules hold, FFMA elsewhere (e.g. N == 1, ld % 4 != 0);
# 'tf32_strict' raises instead of taking the FFMA kernel.
# 'tf32_fwd3' (the default): nn.Linear forward passes in 3xTF32 with the weights split once per optimizer step
# (gnnb200_linear_x3w_f32), backward GEMMs in plain tf32.  The forward activations decide the ReLU masks of the backward pass:
# plain tf32 there flips enough of them to push end-to-end gradients past 2e-2 (scripts/precision_study.py: 2.4e-2 .. 1e-1),
# plain tf32 in dX / dW alone stays at 3e-3 .. 6e-3.
PRECISIONS = {'f32': L.GEMM_F32, 'tf32': L.GEMM_AUTO, 'tf32_strict': L.GEMM_TF32,
              'tf32x3': L.GEMM_AUTO_X3, 'tf32x3_strict': L.GEMM_TF32X3, 'tf32_fwd3': L.GEMM_AUTO_FWD3}
_PAD_PRECISIONS = (L.GEMM_AUTO, L.GEMM_AUTO_X3, L.GEMM_AUTO_FWD3)     # the non-strict tensor-core modes
# raw weights as the hi operand (kind::tf32 reads only the upper 19 bits of a word): the splitter warps then only write lo(A)
X3W_RAW_HI = os.environ.get('GNNB200_X3W_RAW_HI', '1') == '1'


def split_weight(w: Tensor) -> Tuple[Optional[Tensor], Tensor]:
    """(hi, lo) of a weight matrix for gnnb200_linear_x3w_f32, cached on the tensor until it is modified in place
    (an optimizer step bumps `_version`) or re-allocated."""
    key = (w._version, w.data_ptr(), X3W_RAW_HI)
    cached = getattr(w, '_gnnb200_split', None)
    if cached is not None and cached[0] == key and not BYPASS_OPERAND_CACHES:
        return cached[1], cached[2]
    src = w.detach()
    rows, pitch = src.size(0), _ld(src)
    if src.is_contiguous():
        hi, lo, count = torch.empty_like(src), torch.empty_like(src), src.numel()
    else:                                                   # a _tma_rows view: split the whole padded buffer, keep its pitch
        hi = torch.empty(rows, pitch, dtype=src.dtype, device=src.device)[:, : src.size(1)]
        lo = torch.empty(rows, pitch, dtype=src.dtype, device=src.device)[:, : src.size(1)]
        count = rows * pitch
    L.check(_invoke('gnnb200_split_tf32_f32', _ptr(src), count, _ptr(hi), _ptr(lo), _stream(src)), 'split_tf32')
    if X3W_RAW_HI:
        hi = None
    if not BYPASS_OPERAND_CACHES:
        w._gnnb200_split = (key, hi, lo)
    return hi, lo


def _linear_fwd_raw(x: Tensor, weight: Tensor, bias: Optional[Tensor], relu: bool, precision: int,
                    residual: Optional[Tensor] = None, want_stats: bool = False):
    """y = x W^T (+bias)(+residual)(ReLU), W [out, in]: the forward GEMM of every Linear.  Under GEMM_AUTO_FWD3 it runs the
    error-compensated kernel on the pre-split weights; any other precision is _gemm_raw."""
    if precision != L.GEMM_AUTO_FWD3:
        return _gemm_raw(x, False, weight, True, bias, relu, precision, residual, want_stats)
    _need_cuda(x, weight, bias, residual)
    x, weight = _rowmajor(x), _rowmajor(weight)
    M, K = x.shape
    N = weight.size(0)
    if K != weight.size(1):
        raise L.Gnnb200Error(f'linear inner dimensions differ: {K} vs {weight.size(1)}')
    if N % 4 == 0 and N >= 8 and M > 0 and K > 0:
        x, weight = _tma_rows(x), _tma_rows(weight)        # K % 4 != 0 (1433, 3703, 21, 7, 37): padded pitch, same logical K
    elif not weight.is_contiguous():
        weight = weight.contiguous()
    hi, lo = split_weight(weight)
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    if bias is not None:
        bias = bias.contiguous()
    if residual is not None:
        residual = _rowmajor(residual)
        if residual.shape != y.shape:
            raise L.Gnnb200Error(f'residual shape {tuple(residual.shape)} != output shape {tuple(y.shape)}')
    csum = cm2 = None
    if want_stats:
        csum = torch.empty(N, dtype=torch.float32, device=x.device)
        cm2 = torch.empty(N, dtype=torch.float32, device=x.device)
    _call_ws('gnnb200_linear_x3w_f32', 'linear (3xTF32, pre-split weights)', x.device, _ptr(x), _ld(x), _ptr(weight),
             _ptr(hi), _ptr(lo), _ld(weight), _ptr(y), _ld(y), M, N, K, _ptr(bias), _ptr(residual),
             _ld(residual) if residual is not None else 0, L.EPI_RELU if relu else L.EPI_NONE, int(X3W_RAW_HI), _ptr(csum),
             _ptr(cm2), stream=_stream(x),
             key=(M, N, K, _ld(x) % 4, _ld(weight) % 4, residual is None or _ld(residual) % 4 == 0, x.data_ptr() % 16, want_stats))
    return (y, csum, cm2) if want_stats else y


def _gemm_raw(a: Tensor, transa: bool, b: Tensor, transb: bool, bias: Optional[Tensor], relu: bool,
              precision: int, residual: Optional[Tensor] = None, want_stats: bool = False):
    """C = op(a) op(b) (+bias)(+residual)(ReLU); with want_stats also (column sums, centred second moments) of C."""
    _need_cuda(a, b, bias, residual)
    a, b = _rowmajor(a), _rowmajor(b)
    M, K = (a.size(1), a.size(0)) if transa else (a.size(0), a.size(1))
    Kb, N = (b.size(1), b.size(0)) if transb else (b.size(0), b.size(1))
    if K != Kb:
        raise L.Gnnb200Error(f'gemm inner dimensions differ: {K} vs {Kb}')
    if precision in _PAD_PRECISIONS and N % 4 == 0 and N >= 8 and M > 0 and K > 0:
        a, b = _tma_rows(a), _tma_rows(b)                  # e.g. x [N, 1433]: pitch 1436, logical width unchanged
    c = torch.empty(M, N, dtype=torch.float32, device=a.device)
    if bias is not None:
        bias = bias.contiguous()
    if residual is not None:
        residual = _rowmajor(residual)
        if residual.shape != c.shape:
            raise L.Gnnb200Error(f'residual shape {tuple(residual.shape)} != output shape {tuple(c.shape)}')
    csum = cm2 = None
    if want_stats:
        csum = torch.empty(N, dtype=torch.float32, device=a.device)
        cm2 = torch.empty(N, dtype=torch.float32, device=a.device)
    _call_ws('gnnb200_gemm_f32', 'gemm', a.device, _ptr(a), _ld(a), int(transa), _ptr(b), _ld(b), int(transb),
             _ptr(c), _ld(c), M, N, K, _ptr(bias), _ptr(residual), _ld(residual) if residual is not None else 0,
             L.EPI_RELU if relu else L.EPI_NONE, precision, _ptr(csum), _ptr(cm2), stream=_stream(a),
             key=(M, N, K, transa, transb, precision, _ld(a) % 4, _ld(b) % 4, residual is None or _ld(residual) % 4 == 0,
                  a.data_ptr() % 16, b.data_ptr() % 16, want_stats))
    return (c, csum, cm2) if want_stats else c


@_op('gemm')
def gemm(a: Tensor, transa: bool, b: Tensor, transb: bool, bias: Optional[Tensor] = None, relu: bool = False,
         precision: int = 0) -> Tensor:
    """C = op(a) op(b) (+bias)(ReLU), raw (no autograd): see gnnb200_gemm_f32."""
    return _gemm_raw(a, transa, b, transb, bias, relu, precision)


@gemm.register_fake
def _(a, transa, b, transb, bias=None, relu=False, precision=0):
    M = a.size(1) if transa else a.size(0)
    N = b.size(0) if transb else b.size(1)
    return a.new_empty(M, N)


@_op('colsum')
def colsum(x: Tensor) -> Tensor:
    """Deterministic per-column sum [F] (bias gradients)."""
    _need_cuda(x)
    x = _rowmajor(x)
    s = torch.empty(x.size(1), dtype=torch.float32, device=x.device)
    _call_ws('gnnb200_colstats_f32', 'colsum', x.device, _ptr(x), _ld(x), x.size(0), x.size(1), _ptr(s), None,
             stream=_stream(x), key=(x.size(0), x.size(1)))
    return s


@colsum.register_fake
def _(x):
    return x.new_empty(x.size(1))


@_op('colstats')
def colstats(x: Tensor) -> Tuple[Tensor, Tensor]:
    """(sum [F], centred second moment [F]) per column: BatchNorm batch statistics."""
    _need_cuda(x)
    x = _rowmajor(x)
    s = torch.empty(x.size(1), dtype=torch.float32, device=x.device)
    m2 = torch.empty(x.size(1), dtype=torch.float32, device=x.device)
    _call_ws('gnnb200_colstats_f32', 'colstats', x.device, _ptr(x), _ld(x), x.size(0), x.size(1), _ptr(s),
             _ptr(m2), stream=_stream(x))
    return s, m2


@colstats.register_fake
def _(x):
    return x.new_empty(x.size(1)), x.new_empty(x.size(1))


@_op('linear')
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], precision: int,
           residual: Optional[Tensor] = None, bias_feeds_norm: bool = False) -> Tensor:
    """y = x W^T + b (+ residual) with W [out, in] (nn.Linear layout); the residual add (GINLayer's
    `gin_conv(h) + h`) happens in the GEMM epilogue.  bias_feeds_norm=True declares that y goes straight
    into a training-mode BatchNorm: d(loss)/d(bias) is then identically zero (the batch mean absorbs any
    constant shift), so the backward writes zeros instead of reducing grad_y over its rows (the reference
    computes the same quantity numerically and gets rounding noise around 0)."""
    return _linear_fwd_raw(x, weight, bias, False, precision, residual)


@linear.register_fake
def _(x, weight, bias, precision, residual=None, bias_feeds_norm=False):
    return x.new_empty(x.size(0), weight.size(0))


def _lin_setup(ctx, inputs, output):
    x, weight, bias, precision, residual, bias_feeds_norm = inputs
    ctx.zero_bias_grad = bias_feeds_norm
    ctx.precision = precision
    ctx.has_bias = bias is not None
    ctx.has_residual = residual is not None
    ctx.save_for_backward(x, weight)


def _lin_backward(ctx, g):
    x, weight = ctx.saved_tensors
    g = g.contiguous()
    gx = gw = gb = None
    if ctx.needs_input_grad[0]:
        gx = gemm(g, False, weight, False, None, False, ctx.precision)      # [M,out] x [out,in]
    if ctx.needs_input_grad[1]:
        if x.size(1) % 4 != 0 and g.size(1) % 4 == 0 and g.size(1) >= 8 and ctx.precision in _PAD_PRECISIONS:
            # in-features not a multiple of 4 (encoders): the tensor-core kernel needs N % 4 == 0, so form dW^T = x^T g
            gw = gemm(x, True, g, False, None, False, ctx.precision).t().contiguous()
        else:
            gw = gemm(g, True, x, False, None, False, ctx.precision)        # [out,M] x [M,in]
    if ctx.has_bias and ctx.needs_input_grad[2]:
        gb = torch.zeros(g.size(1), dtype=g.dtype, device=g.device) if ctx.zero_bias_grad else colsum(g)
    gres = g if (ctx.has_residual and ctx.needs_input_grad[4]) else None
    return gx, gw, gb, None, gres, None


linear.register_autograd(_lin_backward, setup_context=_lin_setup)


@_op('linear_stats')
def linear_stats(x: Tensor, weight: Tensor, bias: Optional[Tensor], precision: int, residual: Optional[Tensor] = None,
                 bias_feeds_norm: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """(y, colsum(y), centred second moment of y's columns): `linear` whose GEMM epilogue also produces the batch
    statistics of the BatchNorm that consumes y (no separate read pass over y)."""
    return _linear_fwd_raw(x, weight, bias, False, precision, residual, True)


@linear_stats.register_fake
def _(x, weight, bias, precision, residual=None, bias_feeds_norm=False):
    n = weight.size(0)
    return x.new_empty(x.size(0), n), x.new_empty(n), x.new_empty(n)


def _lin_stats_backward(ctx, g, g_sum, g_m2):
    return _lin_backward(ctx, g)              # the statistics are consumed under no_grad (see bn_act's backward)


linear_stats.register_autograd(_lin_stats_backward, setup_context=_lin_setup)


# ---------------------------------------------------------------------------------------------
# fused BatchNorm1d (+ReLU)(+dropout)
# ---------------------------------------------------------------------------------------------
@_op('bn_batch_stats', mutates=('running_mean', 'running_var'))
def bn_batch_stats(x: Tensor, running_mean: Optional[Tensor], running_var: Optional[Tensor], momentum: float,
                   eps: float) -> Tuple[Tensor, Tensor]:
    """(mean, invstd) over the rows of x; the running buffers are updated in place with torch's rule
    (momentum, unbiased variance).  Not differentiable by itself: bn_act's backward carries the full
    batch-statistics Jacobian."""
    _need_cuda(x, running_mean, running_var)
    x = _rowmajor(x)
    rows, cols = x.shape
    dev = x.device
    st = _stream(x)
    s = torch.empty(cols, dtype=torch.float32, device=dev)
    m2 = torch.empty(cols, dtype=torch.float32, device=dev)
    _call_ws('gnnb200_colstats_f32', 'bn colstats', dev, _ptr(x), _ld(x), rows, cols, _ptr(s), _ptr(m2), stream=st,
             key=(rows, cols))
    mean = torch.empty(cols, dtype=torch.float32, device=dev)
    invstd = torch.empty(cols, dtype=torch.float32, device=dev)
    L.check(_invoke('gnnb200_bn_finalize_f32', _ptr(s), _ptr(m2), rows, cols, eps, momentum, _ptr(running_mean),
                    _ptr(running_var), _ptr(mean), _ptr(invstd), st), 'bn_finalize')
    return mean, invstd


@bn_batch_stats.register_fake
def _(x, running_mean, running_var, momentum, eps):
    return x.new_empty(x.size(1)), x.new_empty(x.size(1))


@_op('bn_stats_finalize', mutates=('running_mean', 'running_var'))
def bn_stats_finalize(col_sum: Tensor, col_m2: Tensor, rows: int, running_mean: Optional[Tensor],
                      running_var: Optional[Tensor], momentum: float, eps: float) -> Tuple[Tensor, Tensor]:
    """(mean, invstd) from column sums / centred second moments produced elsewhere (the GEMM epilogue); running
    buffers updated in place like bn_batch_stats."""
    _need_cuda(col_sum, col_m2, running_mean, running_var)
    cols = col_sum.numel()
    mean = torch.empty(cols, dtype=torch.float32, device=col_sum.device)
    invstd = torch.empty(cols, dtype=torch.float32, device=col_sum.device)
    L.check(_invoke('gnnb200_bn_finalize_f32', _ptr(col_sum), _ptr(col_m2), rows, cols, eps, momentum, _ptr(running_mean),
                    _ptr(running_var), _ptr(mean), _ptr(invstd), _stream(col_sum)), 'bn_finalize')
    return mean, invstd


@bn_stats_finalize.register_fake
def _(col_sum, col_m2, rows, running_mean, running_var, momentum, eps):
    return torch.empty_like(col_sum), torch.empty_like(col_sum)


@_op('bn_act')
def bn_act(x: Tensor, mean: Tensor, invstd: Tensor, gamma: Tensor, beta: Tensor, relu: bool, drop_p: float,
           seed: int, training: bool, rows_total: int = 0) -> Tensor:
    """y = drop(relu((x - mean) * invstd * gamma + beta)).  training=True means mean/invstd are the batch
    statistics of x (the backward then includes their Jacobian); False = constants (eval mode).
    rows_total > 0: x is this rank's row shard of a node-partitioned batch of rows_total rows whose
    statistics were reduced over SYNC_GROUP; the backward all-reduces dgamma/dbeta the same way."""
    _need_cuda(x, mean, invstd, gamma, beta)
    x = _rowmajor(x)
    rows, cols = x.shape
    y = torch.empty(rows, cols, dtype=torch.float32, device=x.device)
    L.check(_invoke('gnnb200_bn_act_fwd_f32', _ptr(x), _ld(x), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(beta),
                    int(relu), drop_p, seed, rows, cols, _ptr(y), _ld(y), _stream(x)), 'bn_act_fwd')
    return y


@bn_act.register_fake
def _(x, mean, invstd, gamma, beta, relu, drop_p, seed, training, rows_total=0):
    return torch.empty_like(x)


SYNC_GROUP = None     # torch.distributed process group of the node partition (set by gnnb200.partition)


def _bn_bwd_call(phase: int, g: Tensor, x: Tensor, mean, invstd, gamma, beta, relu, drop_p, seed, training,
                 rows_total: int, gx: Optional[Tensor], dgamma: Tensor, dbeta: Tensor) -> None:
    rows, cols = x.shape
    _call_ws('gnnb200_bn_act_bwd_f32', 'bn_act_bwd', x.device, _ptr(g), _ld(g), _ptr(x), _ld(x), _ptr(mean),
             _ptr(invstd), _ptr(gamma), _ptr(beta), int(relu), drop_p, seed, int(training), phase, rows,
             rows_total if rows_total > 0 else rows, cols, _ptr(gx), _ld(gx) if gx is not None else 0,
             _ptr(dgamma), _ptr(dbeta), stream=_stream(x), key=(rows, cols))


@_op('bn_act_bwd')
def bn_act_bwd(grad_y: Tensor, x: Tensor, mean: Tensor, invstd: Tensor, gamma: Tensor, beta: Tensor, relu: bool,
               drop_p: float, seed: int, training: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """Single-device backward: (grad_x, dgamma, dbeta)."""
    _need_cuda(grad_y, x, mean, invstd, gamma, beta)
    g, x = _rowmajor(grad_y), _rowmajor(x)
    rows, cols = x.shape
    gx = torch.empty(rows, cols, dtype=torch.float32, device=x.device)
    dgamma = torch.empty(cols, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(cols, dtype=torch.float32, device=x.device)
    _bn_bwd_call(0, g, x, mean, invstd, gamma, beta, relu, drop_p, seed, training, 0, gx, dgamma, dbeta)
    return gx, dgamma, dbeta


@bn_act_bwd.register_fake
def _(grad_y, x, mean, invstd, gamma, beta, relu, drop_p, seed, training):
    return torch.empty_like(x), x.new_empty(x.size(1)), x.new_empty(x.size(1))


@_op('bn_act_bwd_reduce')
def bn_act_bwd_reduce(grad_y: Tensor, x: Tensor, mean: Tensor, invstd: Tensor, gamma: Tensor, beta: Tensor,
                      relu: bool, drop_p: float, seed: int) -> Tensor:
    """Phase 1 of the partitioned backward: this shard's [2, C] = (dgamma, dbeta) partial sums."""
    _need_cuda(grad_y, x, mean, invstd, gamma, beta)
    g, x = _rowmajor(grad_y), _rowmajor(x)
    out = torch.empty(2, x.size(1), dtype=torch.float32, device=x.device)
    _bn_bwd_call(1, g, x, mean, invstd, gamma, beta, relu, drop_p, seed, True, 0, None, out[0], out[1])
    return out


@bn_act_bwd_reduce.register_fake
def _(grad_y, x, mean, invstd, gamma, beta, relu, drop_p, seed):
    return x.new_empty(2, x.size(1))


@_op('bn_act_bwd_apply')
def bn_act_bwd_apply(grad_y: Tensor, x: Tensor, mean: Tensor, invstd: Tensor, gamma: Tensor, beta: Tensor,
                     dgamma_dbeta: Tensor, relu: bool, drop_p: float, seed: int, rows_total: int) -> Tensor:
    """Phase 2: grad_x of this shard from the all-reduced (dgamma, dbeta) and the global row count."""
    _need_cuda(grad_y, x, mean, invstd, gamma, beta, dgamma_dbeta)
    g, x = _rowmajor(grad_y), _rowmajor(x)
    gx = torch.empty(x.size(0), x.size(1), dtype=torch.float32, device=x.device)
    dd = dgamma_dbeta.contiguous()
    _bn_bwd_call(2, g, x, mean, invstd, gamma, beta, relu, drop_p, seed, True, rows_total, gx, dd[0], dd[1])
    return gx


@bn_act_bwd_apply.register_fake
def _(grad_y, x, mean, invstd, gamma, beta, dgamma_dbeta, relu, drop_p, seed, rows_total):
    return torch.empty_like(x)


def _bn_setup(ctx, inputs, output):
    x, mean, invstd, gamma, beta, relu, drop_p, seed, training, rows_total = inputs
    ctx.cfg = (relu, drop_p, seed, training, rows_total)
    # the group the forward's statistics were reduced over: the backward must reduce dgamma/dbeta over the same one even
    # when it runs after partition_scope() has been left (e.g. loss.backward() outside the scope)
    ctx.group = SYNC_GROUP if rows_total > 0 else None
    ctx.save_for_backward(x, mean, invstd, gamma, beta)


def _bn_backward(ctx, gy):
    x, mean, invstd, gamma, beta = ctx.saved_tensors
    relu, drop_p, seed, training, rows_total = ctx.cfg
    gy = gy.contiguous()
    if rows_total > 0 and training:
        group = getattr(ctx, 'group', None)
        if group is None:
            raise L.Gnnb200Error('bn_act: the forward ran on a row shard (rows_total > 0) without a process group; '
                                 'the single-device backward would use local sums against global statistics')
        import torch.distributed as dist
        local = bn_act_bwd_reduce(gy, x, mean, invstd, gamma, beta, relu, drop_p, seed)
        total = local.clone()
        dist.all_reduce(total, group=group)
        gx = bn_act_bwd_apply(gy, x, mean, invstd, gamma, beta, total, relu, drop_p, seed, rows_total)
        # parameter gradients stay per-shard partial sums: the flat gradient all-reduce adds them up
        return gx, None, None, local[0], local[1], None, None, None, None, None
    gx, dgamma, dbeta = bn_act_bwd(gy, x, mean, invstd, gamma, beta, relu, drop_p, seed, training)
    return gx, None, None, dgamma, dbeta, None, None, None, None, None


bn_act.register_autograd(_bn_backward, setup_context=_bn_setup)


# ---------------------------------------------------------------------------------------------
# head tails and loss sums (csrc/heads.cu): one launch forward, one backward
# ---------------------------------------------------------------------------------------------
class _LinearAct(torch.autograd.Function):
    """y = dropout(relu(x W^T + b)): the hidden layers of MLPHead (reference src/models/heads.py:41-45).  Without dropout
    the ReLU runs in the GEMM epilogue (one launch); with it, GEMM + one ReLU/dropout launch.  The backward recovers the
    kept-and-positive mask from y (no second Philox pass), then the two GEMMs of the Linear."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], precision: int, drop_p: float, seed: int):
        if drop_p > 0.0:
            a = _linear_fwd_raw(x, weight, bias, False, precision)
            y = torch.empty_like(a)
            L.check(_invoke('gnnb200_act_dropout_fwd_f32', _ptr(a), a.numel(), 1, drop_p, seed, _ptr(y), _stream(a)),
                    'act_dropout_fwd')
        else:
            y = _linear_fwd_raw(x, weight, bias, True, precision)
        ctx.cfg = (precision, drop_p, bias is not None)
        ctx.save_for_backward(x, weight, y)
        return y

    @staticmethod
    def backward(ctx, g: Tensor):
        x, weight, y = ctx.saved_tensors
        precision, drop_p, has_bias = ctx.cfg
        g = g.contiguous()
        ga = torch.empty_like(y)
        L.check(_invoke('gnnb200_act_dropout_bwd_f32', _ptr(g), _ptr(y), y.numel(), drop_p, _ptr(ga), _stream(y)),
                'act_dropout_bwd')
        gx = gemm(ga, False, weight, False, None, False, precision) if ctx.needs_input_grad[0] else None
        gw = None
        if ctx.needs_input_grad[1]:
            if x.size(1) % 4 != 0 and ga.size(1) % 4 == 0 and ga.size(1) >= 8 and precision in _PAD_PRECISIONS:
                gw = gemm(x, True, ga, False, None, False, precision).t().contiguous()
            else:
                gw = gemm(ga, True, x, False, None, False, precision)
        gb = colsum(ga) if has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None, None, None


def linear_act(x: Tensor, weight: Tensor, bias: Optional[Tensor], precision: int, drop_p: float, seed: int) -> Tensor:
    return _LinearAct.apply(x, weight, bias, precision, drop_p, seed)


class _ScaleGrad(torch.autograd.Function):
    """Gradient reversal (reference src/models/heads.py:16-24): identity forward, grad * (-lambda) backward in one launch."""

    @staticmethod
    def forward(ctx, x: Tensor, lambda_val: float):
        ctx.alpha = -float(lambda_val)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g: Tensor):
        g = g.contiguous()
        out = torch.empty_like(g)
        L.check(_invoke('gnnb200_scale_f32', _ptr(g), g.numel(), ctx.alpha, _ptr(out), _stream(g)), 'scale')
        return out, None


def gradient_reversal(x: Tensor, lambda_val: float) -> Tensor:
    return _ScaleGrad.apply(x, lambda_val)


@_op('mse_sum')
def mse_sum(pred: Tensor, target: Tensor) -> Tensor:
    """sum((pred - target)^2) as a 0-d tensor = F.mse_loss(pred, target, reduction='sum'); gradient for `pred` only."""
    _need_cuda(pred, target)
    if pred.shape != target.shape:
        raise L.Gnnb200Error(f'mse_sum: shapes differ: {tuple(pred.shape)} vs {tuple(target.shape)}')
    a, b = pred.contiguous(), target.contiguous()
    out = torch.empty(1, dtype=torch.float32, device=a.device)
    _call_ws('gnnb200_sqdiff_sum_f32', 'mse_sum', a.device, _ptr(a), _ptr(b), a.numel(), _ptr(out), stream=_stream(a), key=())
    return out.view(())


@mse_sum.register_fake
def _(pred, target):
    return pred.new_empty(())


def _mse_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _mse_backward(ctx, g):
    pred, target = ctx.saved_tensors
    a, b = pred.contiguous(), target.contiguous()
    ga = torch.empty_like(a)
    L.check(_invoke('gnnb200_sqdiff_bwd_f32', _ptr(a), _ptr(b), _ptr(g.contiguous().view(1)), a.numel(), _ptr(ga), _stream(a)),
            'mse_sum backward')
    return ga.view_as(pred), None


mse_sum.register_autograd(_mse_backward, setup_context=_mse_setup)


@_op('sigmoid_bce_sum')
def sigmoid_bce_sum(logits: Tensor, labels: Tensor) -> Tuple[Tensor, Tensor]:
    """(probs = sigmoid(logits), loss = F.binary_cross_entropy(probs, labels, reduction='sum') as a 0-d tensor): the tail
    of the link-prediction task (reference src/models/heads.py:67 + src/pretrain/tasks.py:120) in one launch."""
    _need_cuda(logits, labels)
    z, t = logits.contiguous().view(-1), labels.contiguous().view(-1)
    if z.numel() != t.numel() or t.dtype != torch.float32:
        raise L.Gnnb200Error('sigmoid_bce_sum: logits and float32 labels of the same size')
    probs = torch.empty_like(z)
    loss = torch.empty(1, dtype=torch.float32, device=z.device)
    _call_ws('gnnb200_sigmoid_bce_fwd_f32', 'sigmoid_bce', z.device, _ptr(z), _ptr(t), z.numel(), _ptr(probs), _ptr(loss),
             stream=_stream(z), key=())
    return probs.view_as(logits), loss.view(())


@sigmoid_bce_sum.register_fake
def _(logits, labels):
    return torch.empty_like(logits), logits.new_empty(())


def _bce_setup(ctx, inputs, output):
    ctx.save_for_backward(output[0], inputs[1])


def _bce_backward(ctx, g_probs, g_loss):
    probs, labels = ctx.saved_tensors
    if g_loss is None:
        g_loss = torch.zeros((), dtype=torch.float32, device=probs.device)
    p, t = probs.contiguous().view(-1), labels.contiguous().view(-1)
    gz = torch.empty_like(p)
    L.check(_invoke('gnnb200_sigmoid_bce_bwd_f32', _ptr(p), _ptr(t), _ptr(g_loss.contiguous().view(1)), p.numel(), _ptr(gz),
                    _stream(p)), 'sigmoid_bce backward')
    return gz.view_as(probs), None        # (probs is an output for metrics only: no gradient flows through it here)


sigmoid_bce_sum.register_autograd(_bce_backward, setup_context=_bce_setup)


@_op('cross_entropy_sum')
def cross_entropy_sum(logits: Tensor, target: Tensor) -> Tuple[Tensor, Tensor]:
    """(loss = F.cross_entropy(logits, target, reduction='sum') as a 0-d tensor, lse [rows])."""
    _need_cuda(logits, target)
    z = _rowmajor(logits)
    t = target.contiguous()
    if t.dtype != torch.int64 or t.numel() != z.size(0):
        raise L.Gnnb200Error('cross_entropy_sum: int64 class indices, one per row')
    lse = torch.empty(z.size(0), dtype=torch.float32, device=z.device)
    loss = torch.empty(1, dtype=torch.float32, device=z.device)
    _call_ws('gnnb200_ce_sum_fwd_f32', 'cross_entropy_sum', z.device, _ptr(z), _ld(z), _ptr(t), z.size(0), z.size(1), _ptr(lse),
             _ptr(loss), stream=_stream(z), key=())
    return loss.view(()), lse


@cross_entropy_sum.register_fake
def _(logits, target):
    return logits.new_empty(()), logits.new_empty(logits.size(0))


def _ce_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1], output[1])


def _ce_backward(ctx, g_loss, g_lse):
    logits, target, lse = ctx.saved_tensors
    if g_loss is None:
        g_loss = torch.zeros((), dtype=torch.float32, device=logits.device)
    z = _rowmajor(logits)
    gz = torch.empty(z.size(0), z.size(1), dtype=torch.float32, device=z.device)
    L.check(_invoke('gnnb200_ce_bwd_f32', _ptr(z), _ld(z), _ptr(target.contiguous()), _ptr(lse), _ptr(g_loss.contiguous().view(1)),
                    z.size(0), z.size(1), _ptr(gz), _ld(gz), _stream(z)), 'cross_entropy_sum backward')
    return gz, None


cross_entropy_sum.register_autograd(_ce_backward, setup_context=_ce_setup)


# ---------------------------------------------------------------------------------------------
# link-prediction decoder features
# ---------------------------------------------------------------------------------------------
@_op('lp_features')
def lp_features(h: Tensor, edges: Tensor) -> Tensor:
    """[E, 3H] = [h_u + h_v, h_u * h_v, |h_u - h_v|] for edges int64 [2, E]."""
    _need_cuda(h, edges)
    h = _rowmajor(h)
    edges = edges.contiguous()
    E, H = edges.size(1), h.size(1)
    feat = torch.empty(E, 3 * H, dtype=torch.float32, device=h.device)
    L.check(_invoke('gnnb200_lp_features_f32', _ptr(h), _ld(h), _ptr(edges), E, H, _ptr(feat), _ld(feat), _stream(h)),
            'lp_features')
    return feat


@lp_features.register_fake
def _(h, edges):
    return h.new_empty(edges.size(1), 3 * h.size(1))


@_op('lp_features_bwd')
def lp_features_bwd(grad_feat: Tensor, h: Tensor, edges: Tensor) -> Tensor:
    _need_cuda(grad_feat, h, edges)
    g, h = _rowmajor(grad_feat), _rowmajor(h)
    edges = edges.contiguous()
    N, H, E = h.size(0), h.size(1), edges.size(1)
    u_ptr, _, u_eid = csr_build(edges, N, True)
    v_ptr, _, v_eid = csr_build(edges, N, False)
    gh = torch.empty(N, H, dtype=torch.float32, device=h.device)
    L.check(_invoke('gnnb200_lp_features_bwd_f32', 
        _ptr(h), _ld(h), _ptr(edges), E, H, _ptr(g), _ld(g), _ptr(u_ptr), _ptr(u_eid), _ptr(v_ptr), _ptr(v_eid),
        N, _ptr(gh), _ld(gh), _stream(h)), 'lp_features_bwd')
    return gh


@lp_features_bwd.register_fake
def _(grad_feat, h, edges):
    return torch.empty_like(h)


def _lpf_setup(ctx, inputs, output):
    h, edges = inputs
    ctx.save_for_backward(h, edges)


def _lpf_backward(ctx, g):
    h, edges = ctx.saved_tensors
    return lp_features_bwd(g.contiguous(), h, edges), None


lp_features.register_autograd(_lpf_backward, setup_context=_lpf_setup)


# ---------------------------------------------------------------------------------------------
# NT-Xent
# ---------------------------------------------------------------------------------------------
@_op('ntxent_fwd')
def ntxent_fwd(z: Tensor, temperature: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(loss [1], zn [2M,D], lse [2M], norm [2M]) for z = cat(view1, view2) [2M, D]."""
    _need_cuda(z)
    z = _rowmajor(z)
    R, D = z.size(0), z.size(1)
    dev = z.device
    zn = torch.empty(R, D, dtype=torch.float32, device=dev)
    lse = torch.empty(R, dtype=torch.float32, device=dev)
    norm = torch.empty(R, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    _call_ws('gnnb200_ntxent_fwd_f32', 'ntxent_fwd', dev, _ptr(z), _ld(z), R, D, ctypes.c_float(temperature),
             _ptr(zn), _ptr(lse), _ptr(norm), _ptr(loss), stream=_stream(z))
    return loss, zn, lse, norm


@ntxent_fwd.register_fake
def _(z, temperature):
    R = z.size(0)
    return z.new_empty(1), torch.empty_like(z), z.new_empty(R), z.new_empty(R)


@_op('ntxent_bwd')
def ntxent_bwd(grad_loss: Tensor, zn: Tensor, lse: Tensor, norm: Tensor, temperature: float) -> Tensor:
    _need_cuda(grad_loss, zn, lse, norm)
    R, D = zn.size(0), zn.size(1)
    gz = torch.empty(R, D, dtype=torch.float32, device=zn.device)
    gl = grad_loss.contiguous().view(1)
    L.check(_invoke('gnnb200_ntxent_bwd_f32', _ptr(zn), _ptr(lse), _ptr(norm), _ptr(gl), R, D,
                                        ctypes.c_float(temperature), _ptr(gz), _ld(gz), _stream(zn)), 'ntxent_bwd')
    return gz


@ntxent_bwd.register_fake
def _(grad_loss, zn, lse, norm, temperature):
    return torch.empty_like(zn)


def _ntx_setup(ctx, inputs, output):
    z, temperature = inputs
    loss, zn, lse, norm = output
    ctx.temperature = temperature
    ctx.save_for_backward(zn, lse, norm)


def _ntx_backward(ctx, g_loss, g_zn, g_lse, g_norm):
    zn, lse, norm = ctx.saved_tensors
    return ntxent_bwd(g_loss.contiguous(), zn, lse, norm, ctx.temperature), None


ntxent_fwd.register_autograd(_ntx_backward, setup_context=_ntx_setup)


# ---- tensor-core NT-Xent: similarity matrix on tcgen05, row reductions around it ----------------------------
@_op('normalize_rows')
def normalize_rows(z: Tensor) -> Tuple[Tensor, Tensor]:
    """(z / max(|z|, 1e-12) row-wise, |z|) = F.normalize(z, dim=1) and the norms."""
    _need_cuda(z)
    z = _rowmajor(z)
    zn = torch.empty(z.size(0), z.size(1), dtype=torch.float32, device=z.device)
    norm = torch.empty(z.size(0), dtype=torch.float32, device=z.device)
    L.check(_invoke('gnnb200_normalize_rows_f32', _ptr(z), _ld(z), z.size(0), z.size(1), _ptr(zn), _ptr(norm), _stream(z)),
            'normalize_rows')
    return zn, norm


@normalize_rows.register_fake
def _(z):
    return torch.empty_like(z), z.new_empty(z.size(0))


@_op('normalize_rows_bwd')
def normalize_rows_bwd(zn: Tensor, grad_zn: Tensor, norm: Tensor) -> Tensor:
    _need_cuda(zn, grad_zn, norm)
    g = _rowmajor(grad_zn)
    out = torch.empty_like(zn)
    L.check(_invoke('gnnb200_normalize_rows_bwd_f32', _ptr(zn), _ptr(g), _ld(g), _ptr(norm), zn.size(0), zn.size(1),
                    _ptr(out), _ld(out), _stream(zn)), 'normalize_rows_bwd')
    return out


@normalize_rows_bwd.register_fake
def _(zn, grad_zn, norm):
    return torch.empty_like(zn)


@_op('ntxent_sim_fwd')
def ntxent_sim_fwd(sim: Tensor, temperature: float) -> Tuple[Tensor, Tensor]:
    """(loss [1], lse [2M]) from the similarity matrix sim = zn zn^T [2M, 2M] (diagonal excluded, positives i <-> i+M)."""
    _need_cuda(sim)
    sim = _rowmajor(sim)
    R = sim.size(0)
    lse = torch.empty(R, dtype=torch.float32, device=sim.device)
    row_loss = torch.empty(max(R, 1), dtype=torch.float32, device=sim.device)
    loss = torch.empty(1, dtype=torch.float32, device=sim.device)
    L.check(_invoke('gnnb200_ntxent_sim_fwd_f32', _ptr(sim), _ld(sim), R, temperature, _ptr(lse), _ptr(row_loss), _ptr(loss),
                    _stream(sim)), 'ntxent_sim_fwd')
    return loss, lse


@ntxent_sim_fwd.register_fake
def _(sim, temperature):
    return sim.new_empty(1), sim.new_empty(sim.size(0))


@_op('ntxent_sim_bwd_', mutates=('sim',))
def ntxent_sim_bwd_(sim: Tensor, temperature: float, lse: Tensor, grad_loss: Tensor) -> None:
    """Overwrite sim in place with dL/dsim."""
    _need_cuda(sim, lse, grad_loss)
    if sim.stride(1) != 1:
        raise L.Gnnb200Error('sim must be row-major')
    gl = grad_loss.contiguous().view(1)
    L.check(_invoke('gnnb200_ntxent_sim_bwd_f32', _ptr(sim), _ld(sim), sim.size(0), temperature, _ptr(lse), _ptr(gl),
                    _stream(sim)), 'ntxent_sim_bwd')


@ntxent_sim_bwd_.register_fake
def _(sim, temperature, lse, grad_loss):
    return None


class _NtXentTensorCore(torch.autograd.Function):
    """NT-Xent with the [2M, 2M] similarity matrix formed on the tensor cores (three tcgen05 GEMMs per step: sim,
    dsim*zn, dsim^T*zn); the matrix lives only between forward and backward and is overwritten by its own gradient."""

    @staticmethod
    def forward(ctx, z: Tensor, temperature: float, precision: int):
        zn, norm = normalize_rows.fn(z)
        sim = _gemm_raw(zn, False, zn, True, None, False, precision)
        loss, lse = ntxent_sim_fwd.fn(sim, float(temperature))
        ctx.saved = (zn, norm, sim, lse)
        ctx.cfg = (float(temperature), precision)
        return loss

    @staticmethod
    def backward(ctx, g_loss: Tensor):
        if ctx.saved is None:
            raise L.Gnnb200Error('NT-Xent (tensor-core path): the similarity matrix was overwritten by its own gradient in '
                                 'the first backward pass; a second backward through the same graph is not supported')
        zn, norm, sim, lse = ctx.saved
        temperature, precision = ctx.cfg
        ntxent_sim_bwd_.fn(sim, temperature, lse, g_loss)
        gzn = _gemm_raw(sim, False, zn, False, None, False, precision)
        gzn = _gemm_raw(sim, True, zn, False, None, False, precision, gzn)       # + dsim^T zn in the epilogue
        ctx.saved = None
        return normalize_rows_bwd.fn(zn, gzn, norm), None, None


def ntxent_tensor_core(z: Tensor, temperature: float, precision: int) -> Tensor:
    return _NtXentTensorCore.apply(z, temperature, precision)


# ---------------------------------------------------------------------------------------------
# gradient surgery on a flat buffer
# ---------------------------------------------------------------------------------------------
@_op('pcgrad')
def pcgrad(task_grads: Tensor, seg_offsets: Tensor, present: Tensor, order: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """(out [P], has_out uint8 [S], counters int32 [T, S, 2]); see gnnb200_pcgrad_f32.
    task_grads [T, P] fp32, seg_offsets int64 [S+1], present uint8 [T, S], order int32 [T] (all CUDA)."""
    _need_cuda(task_grads, seg_offsets, present, order)
    tg = task_grads.contiguous()
    T, Pn = tg.shape
    S = seg_offsets.numel() - 1
    dev = tg.device
    rank_of = torch.empty_like(order)
    rank_of[order.long()] = torch.arange(T, dtype=order.dtype, device=dev)
    work = torch.empty_like(tg)
    out = torch.zeros(Pn, dtype=torch.float32, device=dev)
    has_out = torch.zeros(S, dtype=torch.uint8, device=dev)
    counters = torch.zeros(T, S, 2, dtype=torch.int32, device=dev)
    L.check(_invoke('gnnb200_pcgrad_f32', _ptr(tg), _ptr(work), T, Pn, _ptr(seg_offsets.contiguous()), S,
                    _ptr(present.contiguous()), _ptr(order.contiguous()), _ptr(rank_of), _ptr(out), _ptr(has_out),
                    _ptr(counters), _stream(tg)), 'pcgrad')
    return out, has_out, counters


@pcgrad.register_fake
def _(task_grads, seg_offsets, present, order):
    S = seg_offsets.numel() - 1
    return (task_grads.new_empty(task_grads.size(1)), task_grads.new_empty(S, dtype=torch.uint8),
            task_grads.new_empty(task_grads.size(0), S, 2, dtype=torch.int32))
