"""GEMM microbenchmark at the C5 shapes (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnnb200
from gnnb200 import ops
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
N = int(2_449_029 * scale)
dev = 'cuda'
P = ops.PRECISIONS[os.environ.get('PREC', 'tf32_strict')]
def run(name, a, ta, b, tb, bias=None, res=None, iters=5):
    for _ in range(2): ops._gemm_raw(a, ta, b, tb, bias, False, P, res)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): c = ops._gemm_raw(a, ta, b, tb, bias, False, P, res)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    M, K = (a.size(1), a.size(0)) if ta else (a.size(0), a.size(1))
    Nn = b.size(0) if tb else b.size(1)
    byt = 4 * (M * K + K * Nn + M * Nn + (M * Nn if res is not None else 0))
    fl = 2.0 * M * Nn * K
    print(f'{name:28s} M={M:8d} N={Nn:4d} K={K:8d}: {ms:7.3f} ms  {byt/ms/1e6:7.0f} GB/s ({byt/ms/1e6/6550.4:.2f} of HBM peak)  {fl/ms/1e9:7.1f} TFLOP/s', flush=True)
x256 = torch.randn(N, 256, device=dev); x512 = torch.randn(N, 512, device=dev); x100 = torch.randn(N, 100, device=dev)
w1 = torch.randn(512, 256, device=dev); w2 = torch.randn(256, 512, device=dev); we = torch.randn(256, 100, device=dev)
b512 = torch.randn(512, device=dev); b256 = torch.randn(256, device=dev)
run('fwd lin1  x[N,256] W1^T', x256, False, w1, True, b512)
run('fwd lin2  r[N,512] W2^T +res', x512, False, w2, True, b256, x256)
run('fwd enc   x[N,100] We^T', x100, False, we, True, b256)
run('dX lin2   g[N,256] W2', x256, False, w2, False)
run('dX lin1   g[N,512] W1', x512, False, w1, False)
run('dW lin2   g[N,256]^T r[N,512]', x256, True, x512, False)
run('dW lin1   g[N,512]^T x[N,256]', x512, True, x256, False)
run('dW enc    g[N,256]^T x[N,100]', x256, True, x100, False)
