"""ORACLE (test infrastructure only): ctypes loader of the plain-C restatement in oracle/c/oracle_c.c.

`build()` compiles it with gcc (-O2 -ffp-contract=off: no fused multiply-add, so fp32 results are the ones a scalar
mul-then-add gives); `lib()` loads it, building on first use.  Wrappers take / return CPU torch tensors."""
import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_int64, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'c', 'oracle_c.c')
LIB = os.path.join(HERE, 'c', 'liboracle_c.so')
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-std=c11', '-Wall', '-Wextra', '-shared', '-fPIC', SRC, '-o', LIB,
                        '-lm'], check=True)
    return LIB


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        P, I = c_void_p, c_int64
        sig = {
            'oc_csr_build': ([P, I, I, c_int, P, P, P], c_int),
            'oc_segment_ptr': ([P, I, I, P], None),
            'oc_coalesce': ([P, I, I, c_int, P], c_int64),
            'oc_gin_aggregate': ([P, I, I, P, I, c_float, c_int, c_int, P], None),
            'oc_segment_pool': ([P, P, I, I, I, c_int, P], None),
            'oc_segment_max_bwd': ([P, P, P, P, I, I, I, P], None),
            'oc_lp_features': ([P, I, P, I, P], None),
            'oc_all_non_edges': ([P, I, I, I, I, P], c_int64),
            'oc_rows_gather': ([P, I, P, I, P], None),
            'oc_rows_fill': ([P, I, P, I, P], None),
        }
        for name, (argtypes, restype) in sig.items():
            fn = getattr(_lib, name)
            fn.argtypes, fn.restype = argtypes, restype
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == 'cpu'
    return t.data_ptr()


def _i(t):
    assert t.dtype == torch.int64 and t.is_contiguous() and t.device.type == 'cpu'
    return t.data_ptr()


def csr_build(edge_index, num_nodes, by_src):
    ei = edge_index.contiguous()
    E = ei.size(1)
    rowptr = torch.empty(num_nodes + 1, dtype=torch.int32)
    col, eid = torch.empty(E, dtype=torch.int32), torch.empty(E, dtype=torch.int32)
    rc = lib().oc_csr_build(_i(ei), E, num_nodes, int(by_src), rowptr.data_ptr(), col.data_ptr(), eid.data_ptr())
    assert rc == 0, rc
    return rowptr, col, eid


def segment_ptr(batch, num_segments):
    ptr = torch.empty(num_segments + 1, dtype=torch.int32)
    lib().oc_segment_ptr(_i(batch.contiguous()), batch.numel(), num_segments, ptr.data_ptr())
    return ptr


def coalesce(edge_index, num_nodes, symmetrise=False):
    ei = edge_index.contiguous()
    E = ei.size(1)
    cap = 2 * E if symmetrise else E
    buf = torch.empty(2 * max(cap, 1), dtype=torch.int64)
    n = lib().oc_coalesce(_i(ei), E, num_nodes, int(symmetrise), buf.data_ptr())
    return buf[:2 * n].view(2, n).clone()


def gin_aggregate(x, edge_index, eps, transposed=False, with_self=True):
    x, ei = x.contiguous(), edge_index.contiguous()
    out = torch.empty_like(x)
    lib().oc_gin_aggregate(_f(x), x.size(0), x.size(1), _i(ei), ei.size(1), float(eps), int(transposed), int(with_self), _f(out))
    return out


def segment_pool(x, batch, num_graphs, mode):
    x = x.contiguous()
    out = torch.empty(num_graphs, x.size(1))
    lib().oc_segment_pool(_f(x), _i(batch.contiguous()), x.size(0), x.size(1), num_graphs, {'sum': 0, 'mean': 1, 'max': 2}[mode], _f(out))
    return out


def segment_max_bwd(grad_out, x, out, batch):
    gx = torch.empty_like(x)
    lib().oc_segment_max_bwd(_f(grad_out.contiguous()), _f(x.contiguous()), _f(out.contiguous()), _i(batch.contiguous()),
                             x.size(0), x.size(1), out.size(0), _f(gx))
    return gx


def lp_features(h, edges):
    h, edges = h.contiguous(), edges.contiguous()
    feat = torch.empty(edges.size(1), 3 * h.size(1))
    lib().oc_lp_features(_f(h), h.size(1), _i(edges), edges.size(1), _f(feat))
    return feat


def all_non_edges(edges, n, want):
    edges = edges.contiguous()
    cap = max(min(want, n * n - n), 1)
    buf = torch.empty(2 * cap, dtype=torch.int64)
    k = lib().oc_all_non_edges(_i(edges), edges.size(1), n, want, cap, buf.data_ptr())
    return buf.view(2, cap)[:, :k].clone()


def rows_gather(x, idx):
    x = x.contiguous()
    out = torch.empty(idx.numel(), x.size(1))
    lib().oc_rows_gather(_f(x), x.size(1), _i(idx.contiguous()), idx.numel(), _f(out))
    return out


def rows_fill(x, idx, row):
    x = x.clone().contiguous()
    lib().oc_rows_fill(_f(x), x.size(1), _i(idx.contiguous()), idx.numel(), _f(row.contiguous()))
    return x
