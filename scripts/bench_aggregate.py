"""Microbenchmark of aggregation kernel variants at C5 scale (dev tool; not part of the product path)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnnb200
from gnnb200 import ops, synthetic, _lib
from gnnb200.graph import Graph

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
loc = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
N, E = int(2_449_029 * scale), int(61_859_140 * scale)
dev = torch.device('cuda')
d = synthetic.products_like(N, E, 4, seed=42, locality=loc, device=dev)
g = Graph(d['edge_index'], N)
x = torch.randn(N, 256, device=dev)
out = torch.empty_like(x)
eps = torch.zeros(1, device=dev)
lib = _lib.load()
fn = lib.gnnb200_dev_aggregate_variant
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64] + [ctypes.c_void_p] * 2 + [ctypes.c_int, ctypes.c_void_p]
fn.restype = ctypes.c_int
st = torch.cuda.current_stream().cuda_stream
bytes_alg = E * 256 * 4 + 2 * N * 256 * 4 + E * 4 + (N + 1) * 4
ref = None
for v in [0, 3, 8, 9, 10, 11, 12, 13]:
    for _ in range(2):
        rc = fn(x.data_ptr(), g.rowptr.data_ptr(), g.col.data_ptr(), N, eps.data_ptr(), out.data_ptr(), v, st)
        assert rc == 0, rc
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn(x.data_ptr(), g.rowptr.data_ptr(), g.col.data_ptr(), N, eps.data_ptr(), out.data_ptr(), v, st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    if ref is None: ref = out.clone()
    ok = torch.equal(ref, out)
    print(f'variant {v}: {ms:8.3f} ms  {bytes_alg / ms / 1e6:8.1f} GB/s  {bytes_alg / ms / 1e6 / 6550.4:.3f} of peak  bitwise_same={ok}', flush=True)
