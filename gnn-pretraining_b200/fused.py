"""One autograd node per GIN layer (reference src/models/gnn.py:26-43).

The op-level path (ops.py) gives every kernel its own autograd node; that is flexible but (i) the gradient
of the layer input reaches it over two edges (the neighbour gather / self term and the `+ h` residual) and the
autograd engine adds them with an extra elementwise pass, and (ii) ~13 Python autograd nodes per layer cost
more host time than the kernels of a small graph batch.  This function runs the same kernels in the same
order with a hand-written backward:

  forward : z = A h + (1+eps) h -> a1 = z W1^T + b1 -> r1 = relu(bn1(a1)) -> s = r1 W2^T + b2 + h -> out = drop(relu(bn2(s)))
  backward: ds = bn2'(g) -> dW2 = ds^T r1, dr1 = ds W2 -> da1 = bn1'(dr1) -> dW1 = da1^T z, dz = da1 W1
            -> dh = ds + A^T dz + (1+eps) dz   (the transposed gather accumulates straight into ds's buffer)
            -> d(eps) = <dz, h>;  db1 = db2 = 0 in training mode (a bias feeding BatchNorm has zero gradient)
"""
import os
from ctypes import byref

import torch
from torch import Tensor

from . import _lib as L
from . import graph as graph_mod
from . import ops
from .graph import Graph

# A layer pass's launches are issued from C++ (gnnb200_gin_layer_fwd/bwd_f32, csrc/gin_layer.cu) instead of one ctypes call
# per kernel.  Same entry points, same arguments, same order (tests/test_native_layer_trace.py compares the two call
# sequences on CPU); what changes is host time — a small-graph step is launch-bound and ~80 % of its launches are inside
# GIN layers.  Measured on B200 (profiles/r02/a_variants.md): C2 fine-tune step 271 -> 361 steps/s, C3 s4 step 10.7 -> 14.1,
# C4 s5 step 6.9 -> 8.5.  GNNB200_NATIVE_LAYER=0 selects the per-kernel Python path.
NATIVE_LAYER = os.environ.get('GNNB200_NATIVE_LAYER', '1') == '1'


def _native_usable(h: Tensor, tensors, bn1, bn2, graph, need_t: bool, precision: int = 0) -> bool:
    """Layouts the composite assumes: dense parameters, fp32 rows, standard BatchNorm buffers, no per-launch timing."""
    if ops.AGG_TIMER is not None or h.dtype != torch.float32 or bn1.running_mean is None or bn2.running_mean is None:
        return False
    if graph_mod.LONG_ROWS and (getattr(graph, 'long_rows', None) is not None or
                                (need_t and getattr(graph, 'long_rows_t', None) is not None)):
        return False                  # hub rows take the block-per-row kernel: Python path
    return all(t is None or t.is_contiguous() for t in tensors)


_p = ops._ptr


def _native_forward(h, eps, w1, b1, g1, be1, w2, b2, g2, be2, graph, bn1, bn2, training, drop_p, seed, precision):
    n, C = h.shape
    H = w1.size(0)
    dev = h.device
    new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)      # noqa: E731
    z, a1, r1, s, out = new(n, C), new(n, H), new(n, H), new(n, C), new(n, C)
    a = L.GinLayerArgs(num_rows=n, hidden=C, mid=H, rowptr=_p(graph.rowptr), col=_p(graph.col), h=_p(h), ldh=ops._ld(h),
                       eps=_p(eps), w1=_p(w1), b1=_p(b1), gamma1=_p(g1), beta1=_p(be1), w2=_p(w2), b2=_p(b2), gamma2=_p(g2),
                       beta2=_p(be2), z=_p(z), a1=_p(a1), r1=_p(r1), s=_p(s), out=_p(out), seed=seed, drop_p=drop_p,
                       training=int(training), precision=precision)
    if precision == L.GEMM_AUTO_FWD3:
        (hi1, lo1), (hi2, lo2) = ops.split_weight(w1), ops.split_weight(w2)      # cached until the optimizer touches them
        a.w1_hi, a.w1_lo, a.w2_hi, a.w2_lo, a.x3w_raw_hi = _p(hi1), _p(lo1), _p(hi2), _p(lo2), int(ops.X3W_RAW_HI)
    if training:
        mean1, invstd1, mean2, invstd2 = new(H), new(H), new(C), new(C)
        for bn, k in ((bn1, '1'), (bn2, '2')):
            if bn.track_running_stats:
                if bn.num_batches_tracked is not None:
                    bn.num_batches_tracked.add_(1)
                setattr(a, 'running_mean' + k, _p(bn.running_mean))
                setattr(a, 'running_var' + k, _p(bn.running_var))
            setattr(a, 'momentum' + k, float(bn.momentum if bn.momentum is not None else 0.0))
            setattr(a, 'bn_eps' + k, float(bn.eps))
        ops._calls['gnnb200_gin_layer_fwd_f32:stats'] = ops._calls.get('gnnb200_gin_layer_fwd_f32:stats', 0) + 1
    else:
        mean1, invstd1 = bn1.running_mean, torch.rsqrt(bn1.running_var + bn1.eps)
        mean2, invstd2 = bn2.running_mean, torch.rsqrt(bn2.running_var + bn2.eps)
    a.mean1, a.invstd1, a.mean2, a.invstd2 = _p(mean1), _p(invstd1), _p(mean2), _p(invstd2)
    ops._call_ws('gnnb200_gin_layer_fwd_f32', 'gin_layer_fwd', dev, byref(a), stream=ops._stream(h),
                 key=(n, C, H, bool(training), precision, ops._ld(h) % 4, h.data_ptr() % 16))
    return out, (z, a1, r1, s, mean1, invstd1, mean2, invstd2)


def _native_backward(g_out, h, eps, w1, g1, be1, w2, g2, be2, z, a1, r1, s, mean1, invstd1, mean2, invstd2, graph,
                     drop_p, seed, precision, need_dh, need_deps):
    n, C = h.shape
    H = w1.size(0)
    dev = h.device
    new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)      # noqa: E731
    ds, dr1, da1, dz = new(n, C), new(n, H), new(n, H), new(n, C)
    dw1, dw2, dg1, dbe1, dg2, dbe2 = new(H, C), new(C, H), new(H), new(H), new(C), new(C)
    dense_h = ops._ld(h) == C
    deps = new(1) if (need_deps and dense_h) else None
    a = L.GinLayerArgs(num_rows=n, hidden=C, mid=H, h=_p(h), ldh=ops._ld(h), eps=_p(eps), w1=_p(w1), gamma1=_p(g1),
                       beta1=_p(be1), w2=_p(w2), gamma2=_p(g2), beta2=_p(be2), mean1=_p(mean1), invstd1=_p(invstd1),
                       mean2=_p(mean2), invstd2=_p(invstd2), z=_p(z), a1=_p(a1), r1=_p(r1), s=_p(s), grad_out=_p(g_out),
                       ds=_p(ds), dr1=_p(dr1), da1=_p(da1), dz=_p(dz), dw1=_p(dw1), dw2=_p(dw2), dgamma1=_p(dg1),
                       dbeta1=_p(dbe1), dgamma2=_p(dg2), dbeta2=_p(dbe2), deps=_p(deps), seed=seed, drop_p=drop_p,
                       training=1, precision=precision, need_dh=int(need_dh))
    if need_dh:
        a.rowptr, a.col = _p(graph.rowptr_t), _p(graph.col_t)
    for flag, tag in ((deps is not None, ':deps'), (need_dh, ':dh')):
        if flag:
            ops._calls['gnnb200_gin_layer_bwd_f32' + tag] = ops._calls.get('gnnb200_gin_layer_bwd_f32' + tag, 0) + 1
    ops._call_ws('gnnb200_gin_layer_bwd_f32', 'gin_layer_bwd', dev, byref(a), stream=ops._stream(h),
                 key=(n, C, H, precision, deps is not None, ops._ld(h) % 4, h.data_ptr() % 16))
    if need_deps and deps is None:
        deps = ops.dot.fn(dz, h)
    return (ds if need_dh else None), deps, dw1, dg1, dbe1, dw2, dg2, dbe2


class GINLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h: Tensor, eps: Tensor, w1: Tensor, b1: Tensor, g1: Tensor, be1: Tensor, w2: Tensor, b2: Tensor,
                g2: Tensor, be2: Tensor, graph: Graph, bn1, bn2, training: bool, drop_p: float, seed: int, precision: int):
        h = ops._rowmajor(h)
        ctx.graph, ctx.cfg = graph, (training, drop_p, seed, precision)
        ctx.native = NATIVE_LAYER and _native_usable(h, (eps, w1, b1, g1, be1, w2, b2, g2, be2), bn1, bn2, graph,
                                                     training and h.requires_grad, precision)
        if ctx.native:
            out, saved = _native_forward(h, eps, w1, b1, g1, be1, w2, b2, g2, be2, graph, bn1, bn2, training, drop_p, seed,
                                         precision)
            ctx.save_for_backward(h, eps, w1, g1, be1, w2, g2, be2, *saved)
            return out
        z = ops._aggregate_raw(h, graph.rowptr, graph.col, L.AGG_SUM, h, eps, None, long_rows=getattr(graph, 'long_rows', None))
        a1 = ops._linear_fwd_raw(z, w1, b1, False, precision)
        mean1, invstd1 = _stats(bn1, a1, training)
        r1 = ops.bn_act.fn(a1, mean1, invstd1, g1, be1, True, 0.0, 0, training, 0)
        s = ops._linear_fwd_raw(r1, w2, b2, False, precision, h)
        mean2, invstd2 = _stats(bn2, s, training)
        out = ops.bn_act.fn(s, mean2, invstd2, g2, be2, True, drop_p, seed, training, 0)
        ctx.save_for_backward(h, eps, w1, g1, be1, w2, g2, be2, z, a1, r1, s, mean1, invstd1, mean2, invstd2)
        return out

    @staticmethod
    def backward(ctx, g_out: Tensor):
        h, eps, w1, g1, be1, w2, g2, be2, z, a1, r1, s, mean1, invstd1, mean2, invstd2 = ctx.saved_tensors
        training, drop_p, seed, precision = ctx.cfg
        graph = ctx.graph
        g_out = g_out.contiguous()
        if ctx.native and training:
            dh, deps, dw1, dg1, dbe1, dw2, dg2, dbe2 = _native_backward(
                g_out, h, eps, w1, g1, be1, w2, g2, be2, z, a1, r1, s, mean1, invstd1, mean2, invstd2, graph, drop_p, seed,
                precision, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
            return (dh, deps, dw1, torch.zeros_like(be1), dg1, dbe1, dw2, torch.zeros_like(be2), dg2, dbe2,
                    None, None, None, None, None, None, None)
        ds, dg2, dbe2 = ops.bn_act_bwd.fn(g_out, s, mean2, invstd2, g2, be2, True, drop_p, seed, training)
        db2 = torch.zeros_like(be2) if training else ops.colsum.fn(ds)
        dw2 = ops._gemm_raw(ds, True, r1, False, None, False, precision)
        dr1 = ops._gemm_raw(ds, False, w2, False, None, False, precision)
        da1, dg1, dbe1 = ops.bn_act_bwd.fn(dr1, a1, mean1, invstd1, g1, be1, True, 0.0, 0, training)
        db1 = torch.zeros_like(be1) if training else ops.colsum.fn(da1)
        dw1 = ops._gemm_raw(da1, True, z, False, None, False, precision)
        dz = ops._gemm_raw(da1, False, w1, False, None, False, precision)
        deps = ops.dot.fn(dz, h) if ctx.needs_input_grad[1] else None
        dh = None
        if ctx.needs_input_grad[0]:
            rowptr_t, col_t = graph.rowptr_t, graph.col_t
            dh = ops._aggregate_raw(dz, rowptr_t, col_t, L.AGG_SUM, dz, eps, None, out=ds,    # ds is dead: reuse it
                                    long_rows=getattr(graph, 'long_rows_t', None))
        return (dh, deps, dw1, db1, dg1, dbe1, dw2, db2, dg2, dbe2, None, None, None, None, None, None, None)


def _stats(bn, x: Tensor, training: bool):
    """(mean, invstd) for a BatchNormAct module: batch statistics + running-buffer update in training mode."""
    if training or bn.running_mean is None:
        upd = training and bn.track_running_stats
        if upd and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        return ops.bn_batch_stats.fn(x, bn.running_mean if upd else None, bn.running_var if upd else None,
                                     float(bn.momentum if bn.momentum is not None else 0.0), float(bn.eps))
    return bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)
