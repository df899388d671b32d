"""The C++ layer composites (gnnb200_gin_layer_fwd/bwd_f32, csrc/gin_layer.cu) against the Python launch sequence of
gnnb200/fused.py — on CPU, without launching anything.

Both paths are sequences of C-ABI calls.  The Python path's calls are captured by replacing ops._invoke / ops._call_ws
with recorders; the composite runs for real in the library's trace mode (gnnb200_dev_trace_begin/_end), where it records
the calls it would make.  After dropping workspace / stream arguments and renaming every pointer by order of first
appearance, the two traces must be IDENTICAL: same entry points, same order, same scalars bit for bit, same dataflow
between buffers.  That is the whole correctness argument of the composite: the kernels themselves are the ones the
GPU suite already covers."""
import ctypes
import struct
from types import SimpleNamespace

import pytest
import torch

import gnnb200  # noqa: F401
from gnnb200 import _lib as L
from gnnb200 import fused, ops

COMPOSITES = ('gnnb200_gin_layer_fwd_f32', 'gnnb200_gin_layer_bwd_f32')


def _words(name, args, raw=False):
    """One call -> list of ('p', address) / ('v', 64-bit word), workspace / size / stream arguments dropped.
    raw: the arguments are already the 64-bit words of the library's trace (floats as their bit patterns)."""
    types = L.SIGNATURES[name]
    has_ws = L.SZP in types
    types = types[:-3] if has_ws else types[:-1]
    assert len(args) >= len(types), (name, len(args), len(types))
    out = []
    for t, v in zip(types, args):
        if t is L.P:
            out.append(('p', int(v or 0)))
        elif t is ctypes.c_float and not raw:
            out.append(('v', struct.unpack('I', struct.pack('f', float(v)))[0]))
        else:
            out.append(('v', int(v) & 0xFFFFFFFFFFFFFFFF))
    return out


def _canonical(trace):
    """[(name, words)] with pointers renamed by first appearance."""
    names = {0: 'NULL'}
    canon = []
    for name, words in trace:
        row = []
        for kind, v in words:
            if kind == 'p':
                row.append(names.setdefault(v, f'buf{len(names)}'))
            else:
                row.append(v)
        canon.append((name, tuple(row)))
    return canon


class _Recorder:
    def __init__(self, monkeypatch):
        self.trace, self.keep = [], []
        self.lib = L.load()
        real_empty = torch.empty

        def empty(*a, **k):                                    # no address is reused inside a trace
            t = real_empty(*a, **k).zero_()
            self.keep.append(t)
            return t

        monkeypatch.setattr(ops, '_invoke', self.invoke)
        monkeypatch.setattr(ops, '_call_ws', self.call_ws)
        monkeypatch.setattr(ops, 'on_device', lambda t: True)
        monkeypatch.setattr(ops, '_stream', lambda t: 0)
        monkeypatch.setattr(ops.torch, 'empty', empty)

    def invoke(self, name, *args):
        self.trace.append((name, _words(name, args)))
        return 0

    def call_ws(self, name, what, device, *args, stream, key=None):
        if name not in COMPOSITES:
            self.trace.append((name, _words(name, args + (None, None, None))))
            return None
        fn = getattr(self.lib, name)
        need = ctypes.c_size_t(0)
        assert fn(*args, None, ctypes.byref(need), None) == 0          # the size query is host arithmetic
        ws = ctypes.create_string_buffer(max(int(need.value), 256))
        buf = (ctypes.c_uint64 * 4096)()
        assert self.lib.gnnb200_dev_trace_begin(buf, len(buf)) == 0
        try:
            have = ctypes.c_size_t(len(ws))
            rc = fn(*args, ctypes.cast(ws, ctypes.c_void_p), ctypes.byref(have), None)
        finally:
            n = self.lib.gnnb200_dev_trace_end()
        assert rc == 0 and n >= 0, (rc, n)
        i = 0
        while i < n:
            fname, argc = L.TRACE_FUNCTIONS[buf[i]], int(buf[i + 1])
            assert argc == len(L.SIGNATURES[fname])
            self.trace.append((fname, _words(fname, list(buf[i + 2: i + 2 + argc]), raw=True)))
            i += 2 + argc
        return None


def _layer_inputs(n, C, e, h_requires_grad, eps_requires_grad, strided_h):
    g = torch.Generator().manual_seed(n + C)
    H = 2 * C
    if strided_h:
        h = torch.randn(n, C + 8, generator=g)[:, :C]
    else:
        h = torch.randn(n, C, generator=g)
    h.requires_grad_(h_requires_grad)
    eps = torch.zeros(1, requires_grad=eps_requires_grad)
    w1, b1 = torch.randn(H, C, generator=g).requires_grad_(), torch.randn(H, generator=g).requires_grad_()
    w2, b2 = torch.randn(C, H, generator=g).requires_grad_(), torch.randn(C, generator=g).requires_grad_()
    bn1, bn2 = torch.nn.BatchNorm1d(H), torch.nn.BatchNorm1d(C)
    rowptr = torch.zeros(n + 1, dtype=torch.int32)
    col = torch.zeros(e, dtype=torch.int32)
    graph = SimpleNamespace(rowptr=rowptr, col=col, rowptr_t=rowptr.clone(), col_t=col.clone())
    return h, eps, w1, b1, bn1, w2, b2, bn2, graph


def _run(monkeypatch, native, training, drop_p, backward, precision='tf32', **kw):
    rec = _Recorder(monkeypatch)
    monkeypatch.setattr(fused, 'NATIVE_LAYER', native)
    h, eps, w1, b1, bn1, w2, b2, bn2, graph = _layer_inputs(**kw)
    bn1.train(training)
    bn2.train(training)
    if precision == 'tf32_fwd3':          # the weight split is cached per optimizer step: warm it so that neither trace holds it
        ops.split_weight(w1), ops.split_weight(w2)
        assert [n for n, _ in rec.trace] == ['gnnb200_split_tf32_f32'] * 2
        rec.trace.clear()
    out = fused.GINLayerFn.apply(h, eps, w1, b1, bn1.weight, bn1.bias, w2, b2, bn2.weight, bn2.bias, graph, bn1, bn2,
                                 training, drop_p, 1234567890123 if drop_p else 0, ops.PRECISIONS[precision])
    if backward:
        out.sum().backward()
    grads = [None if t.grad is None else tuple(t.grad.shape) for t in (h, eps, w1, b1, w2, b2, bn1.weight, bn1.bias,
                                                                       bn2.weight, bn2.bias)]
    counts = (int(bn1.num_batches_tracked), int(bn2.num_batches_tracked))
    return _canonical(rec.trace), grads, counts


CASES = [
    dict(training=True, drop_p=0.2, backward=True, n=300, C=256, e=900, h_requires_grad=True, eps_requires_grad=True, strided_h=False),
    dict(training=True, drop_p=0.0, backward=True, n=77, C=64, e=10, h_requires_grad=False, eps_requires_grad=True, strided_h=False),
    dict(training=True, drop_p=0.2, backward=True, n=64, C=128, e=0, h_requires_grad=True, eps_requires_grad=False, strided_h=True),
    dict(training=False, drop_p=0.0, backward=False, n=300, C=256, e=900, h_requires_grad=False, eps_requires_grad=False, strided_h=False),
    dict(training=True, drop_p=0.2, backward=True, n=5000, C=256, e=20000, h_requires_grad=True, eps_requires_grad=True, strided_h=False),
]


@pytest.mark.parametrize('precision,raw_hi', [('tf32', False), ('tf32_fwd3', False), ('tf32_fwd3', True)])
@pytest.mark.parametrize('case', CASES, ids=[f'case{i}' for i in range(len(CASES))])
def test_composite_makes_the_same_calls_as_the_python_path(monkeypatch, case, precision, raw_hi):
    monkeypatch.setattr(ops, 'X3W_RAW_HI', raw_hi)
    want, grads_py, counts_py = _run(monkeypatch, False, precision=precision, **case)
    got, grads_c, counts_c = _run(monkeypatch, True, precision=precision, **case)
    assert [n for n, _ in got] == [n for n, _ in want]
    for (name, a), (_, b) in zip(got, want):
        assert a == b, f'{name}:\n composite {a}\n python    {b}'
    assert grads_c == grads_py and counts_c == counts_py
    # the composite really replaced the per-kernel calls: the expected sequence for a training step
    if case['training'] and case['backward'] and case['h_requires_grad'] and case['eps_requires_grad']:
        fwd_gemm = 'linear_x3w' if precision == 'tf32_fwd3' else 'gemm'
        assert [n[8:-4] for n, _ in got] == [
            'aggregate', fwd_gemm, 'colstats', 'bn_finalize', 'bn_act_fwd', fwd_gemm, 'colstats', 'bn_finalize', 'bn_act_fwd',
            'bn_act_bwd', 'gemm', 'gemm', 'bn_act_bwd', 'gemm', 'gemm', 'dot', 'aggregate']


def test_strided_input_with_eps_gradient_takes_the_dot_product_outside(monkeypatch):
    """h with a leading dimension: the composite cannot run the dense dot product, the Python side adds it afterwards —
    same calls as the Python path, the (independent) dot product and the transposed gather swapped."""
    case = dict(training=True, drop_p=0.0, backward=True, n=50, C=64, e=30, h_requires_grad=True, eps_requires_grad=True,
                strided_h=True)
    want, _, _ = _run(monkeypatch, False, **case)
    got, _, _ = _run(monkeypatch, True, **case)
    assert sorted(n for n, _ in got) == sorted(n for n, _ in want)
    assert [n for n, _ in got][-2:] == ['gnnb200_aggregate_f32', 'gnnb200_dot_f32']


def test_composite_argument_checks():
    lib = L.load()
    need = ctypes.c_size_t(0)
    a = L.GinLayerArgs(num_rows=10, hidden=0, mid=512)
    assert lib.gnnb200_gin_layer_fwd_f32(ctypes.byref(a), None, ctypes.byref(need), None) == L.EINVAL
    a = L.GinLayerArgs(num_rows=10, hidden=256, mid=512, ldh=256, training=0)
    assert lib.gnnb200_gin_layer_bwd_f32(ctypes.byref(a), None, ctypes.byref(need), None) == L.EUNSUPPORTED
    a.training = 1
    assert lib.gnnb200_gin_layer_fwd_f32(ctypes.byref(a), None, ctypes.byref(need), None) == 0 and need.value > 0
    small = ctypes.c_size_t(8)
    assert lib.gnnb200_gin_layer_fwd_f32(ctypes.byref(a), ctypes.c_void_p(64), ctypes.byref(small), None) == L.EWORKSPACE
    big = ctypes.c_size_t(need.value)
    assert lib.gnnb200_gin_layer_fwd_f32(ctypes.byref(a), ctypes.c_void_p(64), ctypes.byref(big), None) == L.EINVAL   # null tensors
    assert lib.gnnb200_gin_layer_fwd_f32(None, None, ctypes.byref(need), None) == L.EINVAL
