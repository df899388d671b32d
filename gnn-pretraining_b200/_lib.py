"""ctypes binding of libgnnb200.so (the C ABI declared in include/gnnb200.h).

There is no CPU or eager fallback: if the shared library is missing, loading fails loudly.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_longlong, c_size_t, c_uint64, c_void_p, POINTER, Structure

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libgnnb200.so')

OK, EINVAL, ERANGE, EWORKSPACE, EUNSUPPORTED = 0, -1, -2, -3, -4
AGG_SUM, AGG_MEAN, AGG_GCN = 0, 1, 2
AGG_ACCUMULATE, AGG_SKIP_LONG, AGG_LONG_ROW = 8, 16, 1024
POOL_SUM, POOL_MEAN, POOL_MAX = 0, 1, 2
GEMM_F32, GEMM_TF32, GEMM_AUTO, GEMM_TF32X3, GEMM_AUTO_X3, GEMM_AUTO_FWD3 = 0, 1, 2, 3, 4, 5
EPI_NONE, EPI_RELU = 0, 1

P = c_void_p
I64 = c_int64
SZP = POINTER(c_size_t)



class GinLayerArgs(Structure):
    """gnnb200_gin_layer_t of include/gnnb200.h (field for field)."""
    _fields_ = ([('num_rows', I64), ('hidden', I64), ('mid', I64), ('rowptr', P), ('col', P), ('h', P), ('ldh', I64),
                 ('eps', P)] +
                [(k, P) for k in ('w1', 'b1', 'gamma1', 'beta1', 'w2', 'b2', 'gamma2', 'beta2',
                                  'running_mean1', 'running_var1', 'running_mean2', 'running_var2',
                                  'mean1', 'invstd1', 'mean2', 'invstd2', 'z', 'a1', 'r1', 's', 'out', 'grad_out',
                                  'ds', 'dr1', 'da1', 'dz', 'dw1', 'dw2', 'dgamma1', 'dbeta1', 'dgamma2', 'dbeta2', 'deps')] +
                [('seed', c_uint64), ('drop_p', c_float), ('momentum1', c_float), ('bn_eps1', c_float),
                 ('momentum2', c_float), ('bn_eps2', c_float), ('training', c_int), ('precision', c_int),
                 ('need_dh', c_int), ('x3w_raw_hi', c_int)] +
                [(k, P) for k in ('w1_hi', 'w1_lo', 'w2_hi', 'w2_lo')])


# name -> argtypes, in the order of include/gnnb200.h
SIGNATURES = {
    'gnnb200_version': [],
    'gnnb200_error_string': [c_int],
    'gnnb200_csr_build_i64': [P, I64, I64, c_int, P, P, P, P, SZP, P],
    'gnnb200_segment_ptr_i64': [P, I64, I64, P, P],
    'gnnb200_coalesce_i64': [P, I64, I64, P, P, P, SZP, P],
    'gnnb200_aggregate_f32': [P, I64, P, P, I64, I64, c_int, P, I64, P, P, P, I64, P],
    'gnnb200_aggregate_long_rows_f32': [P, I64, P, P, P, I64, I64, c_int, P, I64, P, P, I64, P],
    'gnnb200_dot_f32': [P, P, I64, P, P, SZP, P],
    'gnnb200_segment_pool_fwd_f32': [P, I64, P, I64, I64, I64, c_int, P, I64, P, SZP, P],
    'gnnb200_segment_pool_bwd_f32': [P, I64, P, I64, P, I64, P, I64, I64, I64, c_int, P, I64, P],
    'gnnb200_rows_gather_f32': [P, I64, P, I64, I64, P, I64, P],
    'gnnb200_rows_scatter_f32': [P, I64, c_int, P, I64, I64, P, I64, P],
    'gnnb200_rows_gather_bwd_f32': [P, I64, P, P, I64, I64, P, I64, P],
    'gnnb200_gemm_f32': [P, I64, c_int, P, I64, c_int, P, I64, I64, I64, I64, P, P, I64, c_int, c_int, P, P, P, SZP, P],
    'gnnb200_split_tf32_f32': [P, I64, P, P, P],
    'gnnb200_linear_x3w_f32': [P, I64, P, P, P, I64, P, I64, I64, I64, I64, P, P, I64, c_int, c_int, P, P, P, SZP, P],
    'gnnb200_colstats_f32': [P, I64, I64, I64, P, P, P, SZP, P],
    'gnnb200_bn_finalize_f32': [P, P, I64, I64, c_float, c_float, P, P, P, P, P],
    'gnnb200_bn_merge_finalize_f32': [P, I64, I64, c_float, c_float, P, P, P, P, P],
    'gnnb200_bn_act_fwd_f32': [P, I64, P, P, P, P, c_int, c_float, c_uint64, I64, I64, P, I64, P],
    'gnnb200_bn_act_bwd_f32': [P, I64, P, I64, P, P, P, P, c_int, c_float, c_uint64, c_int, c_int, I64, I64, I64, P, I64, P, P, P, SZP, P],
    'gnnb200_lp_features_f32': [P, I64, P, I64, I64, P, I64, P],
    'gnnb200_lp_features_bwd_f32': [P, I64, P, I64, I64, P, I64, P, P, P, P, I64, P, I64, P],
    'gnnb200_ntxent_fwd_f32': [P, I64, I64, I64, c_float, P, P, P, P, P, SZP, P],
    'gnnb200_ntxent_bwd_f32': [P, P, P, P, I64, I64, c_float, P, I64, P],
    'gnnb200_normalize_rows_f32': [P, I64, I64, I64, P, P, P],
    'gnnb200_normalize_rows_bwd_f32': [P, P, I64, P, I64, I64, P, I64, P],
    'gnnb200_ntxent_sim_fwd_f32': [P, I64, I64, c_float, P, P, P, P],
    'gnnb200_ntxent_sim_bwd_f32': [P, I64, I64, c_float, P, P, P],
    'gnnb200_act_dropout_fwd_f32': [P, I64, c_int, c_float, c_uint64, P, P],
    'gnnb200_act_dropout_bwd_f32': [P, P, I64, c_float, P, P],
    'gnnb200_scale_f32': [P, I64, c_float, P, P],
    'gnnb200_sqdiff_sum_f32': [P, P, I64, P, P, SZP, P],
    'gnnb200_sqdiff_bwd_f32': [P, P, P, I64, P, P],
    'gnnb200_sigmoid_bce_fwd_f32': [P, P, I64, P, P, P, SZP, P],
    'gnnb200_sigmoid_bce_bwd_f32': [P, P, P, I64, P, P],
    'gnnb200_ce_sum_fwd_f32': [P, I64, P, I64, I64, P, P, P, SZP, P],
    'gnnb200_ce_bwd_f32': [P, I64, P, P, P, I64, I64, P, I64, P],
    'gnnb200_negsample_count_i64': [P, I64, P, P, I64, P, I64, P, P, P],
    'gnnb200_negsample_write_i64': [P, I64, P, P, I64, P, P, I64, P, P],
    'gnnb200_host_py_sample_range': [P, P, I64, I64, P],
    'gnnb200_pcgrad_f32': [P, P, I64, I64, P, I64, P, P, P, P, P, P, P],
    'gnnb200_aggregate_peer_f32': [P, c_int, I64, P, P, I64, I64, P, I64, P, P, I64, P],
    'gnnb200_peer_publish_f32': [P, I64, I64, I64, P, I64, P],
    'gnnb200_peer_copy_f32': [P, P, I64, P],
    'gnnb200_peer_alloc': [c_size_t, POINTER(c_void_p), P],
    'gnnb200_peer_open': [P, POINTER(c_void_p)],
    'gnnb200_peer_close': [P],
    'gnnb200_peer_free': [P],
    'gnnb200_gin_layer_fwd_f32': [POINTER(GinLayerArgs), P, SZP, P],
    'gnnb200_gin_layer_bwd_f32': [POINTER(GinLayerArgs), P, SZP, P],
    'gnnb200_dev_trace_begin': [P, c_size_t],
    'gnnb200_dev_trace_end': [],
}
TRACE_FUNCTIONS = ('gnnb200_aggregate_f32', 'gnnb200_gemm_f32', 'gnnb200_colstats_f32', 'gnnb200_bn_finalize_f32',
                   'gnnb200_bn_act_fwd_f32', 'gnnb200_bn_act_bwd_f32', 'gnnb200_dot_f32', 'gnnb200_linear_x3w_f32')   # ids of gnnb200_dev_trace_*
PEER_SHIFT, MAX_PEERS, PEER_HANDLE_BYTES = 28, 16, 64

_lib = None


class Gnnb200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise Gnnb200Error(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(there is no CPU / eager fallback for the gnnb200 hot path)')
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError -> header and library disagree
        fn.argtypes = argtypes
        fn.restype = (c_char_p if name == 'gnnb200_error_string' else
                      c_longlong if name == 'gnnb200_dev_trace_end' else c_int)
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().gnnb200_error_string(code).decode()
        raise Gnnb200Error(f'{what} failed: {msg} (code {code})')
