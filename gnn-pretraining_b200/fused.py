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
import torch
from torch import Tensor

from . import _lib as L
from . import ops
from .graph import Graph


class GINLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h: Tensor, eps: Tensor, w1: Tensor, b1: Tensor, g1: Tensor, be1: Tensor, w2: Tensor, b2: Tensor,
                g2: Tensor, be2: Tensor, graph: Graph, bn1, bn2, training: bool, drop_p: float, seed: int, precision: int):
        h = ops._rowmajor(h)
        z = ops._aggregate_raw(h, graph.rowptr, graph.col, L.AGG_SUM, h, eps, None)
        a1 = ops._gemm_raw(z, False, w1, True, b1, False, precision)
        mean1, invstd1 = _stats(bn1, a1, training)
        r1 = ops.bn_act.fn(a1, mean1, invstd1, g1, be1, True, 0.0, 0, training, 0)
        s = ops._gemm_raw(r1, False, w2, True, b2, False, precision, h)
        mean2, invstd2 = _stats(bn2, s, training)
        out = ops.bn_act.fn(s, mean2, invstd2, g2, be2, True, drop_p, seed, training, 0)
        ctx.save_for_backward(h, eps, w1, g1, be1, w2, g2, be2, z, a1, r1, s, mean1, invstd1, mean2, invstd2)
        ctx.graph, ctx.cfg = graph, (training, drop_p, seed, precision)
        return out

    @staticmethod
    def backward(ctx, g_out: Tensor):
        h, eps, w1, g1, be1, w2, g2, be2, z, a1, r1, s, mean1, invstd1, mean2, invstd2 = ctx.saved_tensors
        training, drop_p, seed, precision = ctx.cfg
        graph = ctx.graph
        g_out = g_out.contiguous()
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
            dh = ops._aggregate_raw(dz, rowptr_t, col_t, L.AGG_SUM, dz, eps, None, out=ds)   # ds is dead: reuse it
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
