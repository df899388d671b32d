"""GEMM microbenchmark at the C5 shapes: every GEMM of a GIN layer (forward, dX, dW), plain tf32 against the
error-compensated kernels.  L2 is flushed between timed launches (a 512 MB fill), CUDA events on the launching stream.
  python scripts/bench_gemm.py [scale] > profiles/...    (dev tool; needs a GPU)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import gnnb200  # noqa: E402,F401
from gnnb200 import ops  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
N = int(2_449_029 * scale)
dev = 'cuda'
PEAK = 6550.4
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
if os.path.isfile(path):
    PEAK = float(json.load(open(path))['hbm_gbs'])
flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)


def timed(fn, iters=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    ms.sort()
    return ms[len(ms) // 2]


def report(name, ms, M, Nn, K, res):
    byt = 4 * (M * K + K * Nn + M * Nn + (M * Nn if res else 0))
    fl = 2.0 * M * Nn * K
    print(f'{name:44s} M={M:8d} N={Nn:4d} K={K:8d}: {ms:7.3f} ms  {byt / ms / 1e6:7.0f} GB/s ({byt / ms / 1e6 / PEAK:.2f} of HBM peak)  '
          f'{fl / ms / 1e9:7.1f} TFLOP/s', flush=True)


def run(name, a, ta, b, tb, bias=None, res=None, precs=('tf32_strict', 'tf32x3_strict')):
    M, K = (a.size(1), a.size(0)) if ta else (a.size(0), a.size(1))
    Nn = b.size(0) if tb else b.size(1)
    for prec in precs:
        P = ops.PRECISIONS[prec]
        report(f'{name} [{prec}]', timed(lambda: ops._gemm_raw(a, ta, b, tb, bias, False, P, res)), M, Nn, K, res is not None)
    if not ta and tb:                              # nn.Linear forward layout: the pre-split-weight kernel, both split modes
        for raw in (False, True):
            ops.X3W_RAW_HI = raw
            fn = lambda: ops._linear_fwd_raw(a, b, bias, False, ops.PRECISIONS['tf32_fwd3'], res)   # noqa: E731
            report(f'{name} [x3w raw_hi={int(raw)}]', timed(fn), M, Nn, K, res is not None)


x256 = torch.randn(N, 256, device=dev)
x512 = torch.randn(N, 512, device=dev)
x100 = torch.randn(N, 100, device=dev)
w1 = torch.randn(512, 256, device=dev)
w2 = torch.randn(256, 512, device=dev)
we = torch.randn(256, 100, device=dev)
b512 = torch.randn(512, device=dev)
b256 = torch.randn(256, device=dev)
run('fwd lin1  x[N,256] W1^T', x256, False, w1, True, b512)
run('fwd lin2  r[N,512] W2^T +res', x512, False, w2, True, b256, x256)
run('fwd enc   x[N,100] We^T', x100, False, we, True, b256)
run('dX lin2   g[N,256] W2', x256, False, w2, False, precs=('tf32_strict',))
run('dX lin1   g[N,512] W1', x512, False, w1, False, precs=('tf32_strict',))
run('dW lin2   g[N,256]^T r[N,512]', x256, True, x512, False, precs=('tf32_strict',))
run('dW lin1   g[N,512]^T x[N,256]', x512, True, x256, False, precs=('tf32_strict',))
run('dW enc    g[N,256]^T x[N,100]', x256, True, x100, False, precs=('tf32_strict',))
