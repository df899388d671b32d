"""Error contract of the C ABI (include/gnnb200.h), checked WITHOUT a GPU: every entry point validates its
arguments on the host before it touches the CUDA runtime, so bad sizes / modes / null pointers / sizes beyond the
int32 CSR come back as GNNB200_EINVAL / GNNB200_ERANGE / GNNB200_EUNSUPPORTED and never reach a launch, and the
workspace-size queries of the kernels that need no library (CUB) query are pure host arithmetic.

The device pointers below are a dummy non-null address: a validation path never dereferences them.  No call here
passes validation with a non-null workspace, so nothing is launched (a launch needs a GPU: tests/test_gpu_*.py)."""
from ctypes import byref, c_size_t

import pytest

import gnnb200  # noqa: F401
from gnnb200 import _lib as L

D = 0x7F0000000000          # "device pointer": never dereferenced on the host
BIG = 2 ** 31               # first size that does not fit the int32 CSR


@pytest.fixture(scope='module')
def lib():
    return L.load()


def _call(lib, name, *args):
    need = c_size_t(0xDEAD)
    rc = getattr(lib, name)(*[byref(need) if a == 'SZ' else a for a in args])
    return rc, need.value


# (entry point, arguments, expected return code) — one row per documented failure mode
BAD = [
    # ---- structure ----
    ('gnnb200_csr_build_i64', (D, -1, 10, 0, D, D, None, None, 'SZ', None), L.EINVAL),          # negative E
    ('gnnb200_csr_build_i64', (D, 10, -1, 0, D, D, None, None, 'SZ', None), L.EINVAL),          # negative N
    ('gnnb200_csr_build_i64', (D, 10, 10, 0, D, D, None, None, None, None), L.EINVAL),          # no size out-parameter
    ('gnnb200_csr_build_i64', (D, BIG, 10, 0, D, D, None, None, 'SZ', None), L.ERANGE),         # E >= 2^31
    ('gnnb200_csr_build_i64', (D, 10, BIG, 0, D, D, None, None, 'SZ', None), L.ERANGE),         # N >= 2^31
    ('gnnb200_segment_ptr_i64', (D, -1, 4, D, None), L.EINVAL),
    ('gnnb200_segment_ptr_i64', (D, 4, 4, None, None), L.EINVAL),                               # ptr == NULL
    ('gnnb200_segment_ptr_i64', (None, 4, 4, D, None), L.EINVAL),                               # ids == NULL with n > 0
    ('gnnb200_segment_ptr_i64', (D, BIG, 4, D, None), L.ERANGE),
    ('gnnb200_coalesce_i64', (D, -1, 10, D, D, None, 'SZ', None), L.EINVAL),
    ('gnnb200_coalesce_i64', (D, 10, 10, D, D, None, None, None), L.EINVAL),
    ('gnnb200_coalesce_i64', (D, BIG, 10, D, D, None, 'SZ', None), L.ERANGE),
    # ---- aggregation ----
    ('gnnb200_aggregate_f32', (D, 256, D, D, -1, 256, L.AGG_SUM, None, 0, None, None, D, 256, None), L.EINVAL),
    ('gnnb200_aggregate_f32', (D, 256, D, D, 10, 256, 3, None, 0, None, None, D, 256, None), L.EINVAL),      # bad mode
    ('gnnb200_aggregate_f32', (D, 256, D, D, 10, 256, L.AGG_MEAN | L.AGG_ACCUMULATE, None, 0, None, None, D, 256, None),
     L.EINVAL),                                                                                 # ACCUMULATE is SUM-only
    ('gnnb200_dot_f32', (D, D, -1, D, None, 'SZ', None), L.EINVAL),
    ('gnnb200_dot_f32', (D, D, 10, D, None, None, None), L.EINVAL),
    # ---- pooling ----
    ('gnnb200_segment_pool_fwd_f32', (D, 256, D, -1, 4, 256, L.POOL_MEAN, D, 256, None, 'SZ', None), L.EINVAL),
    ('gnnb200_segment_pool_fwd_f32', (D, 256, D, 10, 4, 256, 7, D, 256, None, 'SZ', None), L.EINVAL),        # bad mode
    ('gnnb200_segment_pool_fwd_f32', (D, 256, D, 10, 4, 256, L.POOL_MEAN, D, 256, None, None, None), L.EINVAL),
    ('gnnb200_segment_pool_fwd_f32', (D, 256, D, 10, 4, 1 << 24, L.POOL_MEAN, D, 256, None, 'SZ', None), L.ERANGE),
    ('gnnb200_segment_pool_bwd_f32', (D, 256, D, 256, D, 256, D, 10, -4, 256, L.POOL_MAX, D, 256, None), L.EINVAL),
    ('gnnb200_segment_pool_bwd_f32', (None, 256, D, 256, D, 256, D, 10, 4, 256, L.POOL_MAX, D, 256, None), L.EINVAL),
    # ---- rows ----
    ('gnnb200_rows_gather_f32', (D, 256, D, -1, 256, D, 256, None), L.EINVAL),
    ('gnnb200_rows_gather_f32', (None, 256, D, 4, 256, D, 256, None), L.EINVAL),
    ('gnnb200_rows_scatter_f32', (D, 256, 0, D, -1, 256, D, 256, None), L.EINVAL),
    ('gnnb200_rows_scatter_f32', (D, 256, 0, None, 4, 256, D, 256, None), L.EINVAL),
    ('gnnb200_rows_gather_bwd_f32', (D, 256, D, D, -1, 256, D, 256, None), L.EINVAL),
    ('gnnb200_rows_gather_bwd_f32', (D, 256, None, D, 4, 256, D, 256, None), L.EINVAL),
    # ---- dense ----
    ('gnnb200_gemm_f32', (D, 8, 0, D, 8, 1, D, 8, -1, 8, 8, None, None, 0, 0, L.GEMM_AUTO, None, None, None, 'SZ', None),
     L.EINVAL),
    ('gnnb200_gemm_f32', (D, 8, 0, D, 8, 1, D, 8, 8, 8, 8, None, None, 0, 0, L.GEMM_AUTO, None, None, None, None, None),
     L.EINVAL),
    ('gnnb200_gemm_f32', (D, 8, 0, D, 8, 1, D, 8, BIG, 8, 8, None, None, 0, 0, L.GEMM_AUTO, None, None, None, 'SZ', None),
     L.ERANGE),
    ('gnnb200_gemm_f32', (D, 8, 0, D, 8, 1, D, 8, 8, 8, 8, None, None, 0, 0, 99, None, None, None, 'SZ', None),
     L.EINVAL),                                                                                 # unknown precision
    ('gnnb200_gemm_f32', (D, 7, 0, D, 8, 1, D, 8, 8, 8, 7, None, None, 0, 0, L.GEMM_TF32, None, None, None, 'SZ', None),
     L.EUNSUPPORTED),                                                                           # lda % 4 != 0: not TMA-legal
    ('gnnb200_split_tf32_f32', (D, -1, D, D, None), L.EINVAL),
    ('gnnb200_split_tf32_f32', (D, 16, None, D, None), L.EINVAL),                               # hi == NULL
    ('gnnb200_linear_x3w_f32', (D, 8, D, D, D, 8, D, 8, -1, 8, 8, None, None, 0, 0, 0, None, None, None, 'SZ', None), L.EINVAL),
    ('gnnb200_linear_x3w_f32', (D, 8, D, D, D, 8, D, 8, 8, 8, 8, None, None, 0, 0, 0, None, None, None, None, None), L.EINVAL),
    ('gnnb200_linear_x3w_f32', (D, 8, D, D, D, 8, D, 8, BIG, 8, 8, None, None, 0, 0, 0, None, None, None, 'SZ', None), L.ERANGE),
    ('gnnb200_linear_x3w_f32', (D, 8, D, None, D, 8, D, 8, 8, 8, 8, None, None, 0, 0, 0, None, None, D, 'SZ', None),
     L.EINVAL),                                                                                 # no W_hi although raw_hi == 0
    ('gnnb200_colstats_f32', (D, 256, -1, 256, D, D, None, 'SZ', None), L.EINVAL),
    ('gnnb200_colstats_f32', (D, 256, 10, 1 << 24, D, D, None, 'SZ', None), L.ERANGE),
    # ---- BatchNorm ----
    ('gnnb200_bn_finalize_f32', (D, D, 0, 256, 1e-5, 0.1, None, None, D, D, None), L.EINVAL),  # rows must be > 0
    ('gnnb200_bn_finalize_f32', (None, D, 10, 256, 1e-5, 0.1, None, None, D, D, None), L.EINVAL),
    ('gnnb200_bn_merge_finalize_f32', (D, 0, 256, 1e-5, 0.1, None, None, D, D, None), L.EINVAL),               # no parts
    ('gnnb200_bn_merge_finalize_f32', (None, 8, 256, 1e-5, 0.1, None, None, D, D, None), L.EINVAL),
    ('gnnb200_bn_act_fwd_f32', (D, 256, D, D, D, D, 1, 1.0, 0, 10, 256, D, 256, None), L.EINVAL),            # p must be < 1
    ('gnnb200_bn_act_fwd_f32', (D, 256, D, D, D, D, 1, -0.1, 0, 10, 256, D, 256, None), L.EINVAL),
    ('gnnb200_bn_act_fwd_f32', (D, 256, None, D, D, D, 1, 0.0, 0, 10, 256, D, 256, None), L.EINVAL),
    ('gnnb200_bn_act_fwd_f32', (D, 30, D, D, D, D, 1, 0.0, 0, 10, 30, D, 30, None), L.EUNSUPPORTED),         # cols % 4 != 0
    ('gnnb200_bn_act_bwd_f32', (D, 256, D, 256, D, D, D, D, 1, 0.0, 0, 1, 3, 10, 10, 256, D, 256, D, D, None, 'SZ', None),
     L.EINVAL),                                                                                 # phase must be 0..2
    ('gnnb200_bn_act_bwd_f32', (D, 256, D, 256, D, D, D, D, 1, 0.0, 0, 1, 2, 10, 5, 256, D, 256, D, D, None, 'SZ', None),
     L.EINVAL),                                                                                 # rows_total < rows
    ('gnnb200_bn_act_bwd_f32', (D, 256, D, 256, D, D, D, D, 1, 1.5, 0, 1, 0, 10, 10, 256, D, 256, D, D, None, 'SZ', None),
     L.EINVAL),
    # ---- heads ----
    ('gnnb200_lp_features_f32', (D, 256, D, -1, 256, D, 768, None), L.EINVAL),
    ('gnnb200_lp_features_f32', (None, 256, D, 4, 256, D, 768, None), L.EINVAL),
    ('gnnb200_lp_features_bwd_f32', (D, 256, D, 4, 256, D, 768, D, D, D, D, -1, D, 256, None), L.EINVAL),
    ('gnnb200_lp_features_bwd_f32', (D, 256, None, 4, 256, D, 768, D, D, D, D, 10, D, 256, None), L.EINVAL),
    ('gnnb200_ntxent_fwd_f32', (D, 128, 7, 128, 0.5, D, D, D, D, None, 'SZ', None), L.EINVAL),               # odd 2M
    ('gnnb200_ntxent_fwd_f32', (D, 128, 1 << 24, 128, 0.5, D, D, D, D, None, 'SZ', None), L.ERANGE),
    ('gnnb200_ntxent_bwd_f32', (D, D, D, D, 8, 100, 0.5, D, 100, None), L.EUNSUPPORTED),                     # D % 16 != 0
    ('gnnb200_ntxent_bwd_f32', (D, D, D, D, 8, 128, 0.0, D, 128, None), L.EINVAL),                           # T must be > 0
    ('gnnb200_normalize_rows_f32', (D, 128, -1, 128, D, D, None), L.EINVAL),
    ('gnnb200_normalize_rows_f32', (None, 128, 4, 128, D, D, None), L.EINVAL),
    ('gnnb200_normalize_rows_bwd_f32', (D, D, 128, D, 4, -128, D, 128, None), L.EINVAL),
    ('gnnb200_ntxent_sim_fwd_f32', (D, 8, 7, 0.5, D, D, D, None), L.EINVAL),
    ('gnnb200_ntxent_sim_fwd_f32', (D, 8, 8, 0.5, D, D, None, None), L.EINVAL),                              # loss == NULL
    ('gnnb200_ntxent_sim_fwd_f32', (D, 1 << 24, 1 << 24, 0.5, D, D, D, None), L.ERANGE),
    ('gnnb200_ntxent_sim_bwd_f32', (D, 8, 8, -1.0, D, D, None), L.EINVAL),
    ('gnnb200_ntxent_sim_bwd_f32', (None, 8, 8, 0.5, D, D, None), L.EINVAL),
    # ---- head tails / loss sums ----
    ('gnnb200_act_dropout_fwd_f32', (D, 16, 1, 1.0, 0, D, None), L.EINVAL),                                  # p must be < 1
    ('gnnb200_act_dropout_fwd_f32', (None, 16, 1, 0.5, 0, D, None), L.EINVAL),
    ('gnnb200_act_dropout_bwd_f32', (D, None, 16, 0.5, D, None), L.EINVAL),
    ('gnnb200_scale_f32', (D, -1, 2.0, D, None), L.EINVAL),
    ('gnnb200_sqdiff_sum_f32', (D, D, -1, D, None, 'SZ', None), L.EINVAL),
    ('gnnb200_sqdiff_sum_f32', (D, D, 16, D, None, None, None), L.EINVAL),
    ('gnnb200_sqdiff_bwd_f32', (D, D, None, 16, D, None), L.EINVAL),
    ('gnnb200_sigmoid_bce_fwd_f32', (D, D, -1, D, D, None, 'SZ', None), L.EINVAL),
    ('gnnb200_sigmoid_bce_bwd_f32', (D, D, D, 16, None, None), L.EINVAL),
    ('gnnb200_ce_sum_fwd_f32', (D, 4, D, 16, 0, D, D, None, 'SZ', None), L.EINVAL),                          # no classes
    ('gnnb200_ce_sum_fwd_f32', (D, 4, D, 16, 1 << 20, D, D, None, 'SZ', None), L.ERANGE),
    ('gnnb200_ce_bwd_f32', (D, 2, D, D, D, 16, 4, D, 4, None), L.EINVAL),                                    # ldz < cols
    # ---- negative sampling ----
    ('gnnb200_negsample_count_i64', (D, 10, D, D, 4, D, -1, D, D, None), L.EINVAL),                            # negative quota
    ('gnnb200_negsample_count_i64', (D, 10, D, D, 4, None, 10, D, D, None), L.EINVAL),
    ('gnnb200_negsample_count_i64', (D, BIG, D, D, 4, D, 10, D, D, None), L.ERANGE),
    ('gnnb200_negsample_write_i64', (D, 10, D, D, 4, D, D, -1, D, None), L.EINVAL),
    ('gnnb200_negsample_write_i64', (D, 10, D, D, 4, D, None, 10, D, None), L.EINVAL),
    ('gnnb200_host_py_sample_range', (None, D, 10, 2, D), L.EINVAL),
    ('gnnb200_host_py_sample_range', (D, None, 10, 2, D), L.EINVAL),
    # ---- gradient surgery ----
    ('gnnb200_pcgrad_f32', (D, D, -1, 10, D, 2, D, D, D, D, D, D, None), L.EINVAL),
    ('gnnb200_pcgrad_f32', (D, None, 2, 10, D, 2, D, D, D, D, D, D, None), L.EINVAL),
    ('gnnb200_pcgrad_f32', (D, D, 70000, 10, D, 2, D, D, D, D, D, D, None), L.ERANGE),                       # > 65535 tasks
    # ---- long rows ----
    ('gnnb200_aggregate_long_rows_f32', (D, 256, D, D, None, 3, 256, 0, None, 0, None, D, 256, None), L.EINVAL),      # no row list
    ('gnnb200_aggregate_long_rows_f32', (D, 256, D, D, D, 3, 256, 1, None, 0, None, D, 256, None), L.EINVAL),         # MEAN
    ('gnnb200_aggregate_long_rows_f32', (D, 254, D, D, D, 3, 254, 0, None, 0, None, D, 256, None), L.EUNSUPPORTED),   # not 128-bit
    ('gnnb200_aggregate_f32', (D, 254, D, D, 4, 254, 16, None, 0, None, None, D, 256, None), L.EUNSUPPORTED),          # SKIP_LONG needs it too
    ('gnnb200_aggregate_f32', (D, 256, D, D, 4, 256, 17, None, 0, None, None, D, 256, None), L.EINVAL),               # SKIP_LONG with MEAN
    # ---- peer-memory aggregation ----
    ('gnnb200_aggregate_peer_f32', (None, 2, 256, D, D, 4, 256, None, 0, None, D, 256, None), L.EINVAL),     # no pointer table
    ('gnnb200_aggregate_peer_f32', (D, 17, 256, D, D, 4, 256, None, 0, None, D, 256, None), L.EINVAL),       # > GNNB200_MAX_PEERS
    ('gnnb200_aggregate_peer_f32', (D, 2, 256, D, D, 4, 256, None, 0, None, None, 256, None), L.EINVAL),     # out == NULL
    ('gnnb200_aggregate_peer_f32', (D, 2, 256, D, D, BIG, 256, None, 0, None, D, 256, None), L.ERANGE),
    ('gnnb200_aggregate_peer_f32', (D, 2, 255, D, D, 4, 255, None, 0, None, D, 256, None), L.EUNSUPPORTED),  # rows not 16-byte
    ('gnnb200_peer_publish_f32', (D, 128, 4, 256, D, 256, None), L.EINVAL),                                  # lds < feat
    ('gnnb200_peer_publish_f32', (None, 256, 4, 256, D, 256, None), L.EINVAL),
    ('gnnb200_peer_copy_f32', (D, None, 16, None), L.EINVAL),
    ('gnnb200_peer_copy_f32', (D, D, -1, None), L.EINVAL),
    ('gnnb200_peer_alloc', (0, None, None), L.EINVAL),
    ('gnnb200_peer_open', (None, None), L.EINVAL),
    ('gnnb200_peer_close', (None,), L.EINVAL),
    ('gnnb200_peer_free', (None,), L.EINVAL),
]


@pytest.mark.parametrize('name,args,want', BAD, ids=[f'{n[8:]}-{i}' for i, (n, _, _) in enumerate(BAD)])
def test_bad_arguments_are_refused_on_the_host(lib, name, args, want):
    assert len(args) == len(L.SIGNATURES[name]), 'test row out of date with include/gnnb200.h'
    rc, _ = _call(lib, name, *args)
    assert rc == want, f'{name}: got {rc} ({lib.gnnb200_error_string(rc).decode()}), want {want}'


def test_every_compute_entry_point_has_an_error_row():
    # the struct-based layer composites and the trace hook have their argument checks in tests/test_native_layer_trace.py
    elsewhere = {'gnnb200_version', 'gnnb200_error_string', 'gnnb200_gin_layer_fwd_f32', 'gnnb200_gin_layer_bwd_f32',
                 'gnnb200_dev_trace_begin', 'gnnb200_dev_trace_end'}
    assert {n for n, _, _ in BAD} == set(L.SIGNATURES) - elsewhere


def test_workspace_queries_are_host_arithmetic(lib):
    """workspace == NULL => the size comes back and nothing is launched (include/gnnb200.h conventions)."""
    rc, n = _call(lib, 'gnnb200_dot_f32', D, D, 1 << 20, D, None, 'SZ', None)
    assert rc == 0 and 0 < n < (1 << 20)
    rc, n1 = _call(lib, 'gnnb200_colstats_f32', D, 256, 100_000, 256, D, D, None, 'SZ', None)
    rc2, n2 = _call(lib, 'gnnb200_colstats_f32', D, 256, 200_000, 256, D, D, None, 'SZ', None)
    assert rc == 0 and rc2 == 0 and 0 < n1 < n2                      # partials grow with the row count
    rc, n = _call(lib, 'gnnb200_ntxent_fwd_f32', D, 128, 200, 128, 0.5, D, D, D, D, None, 'SZ', None)
    assert rc == 0 and n > 0
    # GEMM: the query works with C == NULL (the support predicate reads no memory).  The tensor path also needs the
    # driver's cuTensorMapEncodeTiled (resolved at run time): without a driver AUTO falls back to the FFMA kernel and
    # the strict TF32 modes answer EUNSUPPORTED instead of crashing.
    for prec in (L.GEMM_F32, L.GEMM_AUTO, L.GEMM_AUTO_X3):
        rc, _ = _call(lib, 'gnnb200_gemm_f32', D, 256, 0, D, 256, 1, None, 512, 100_000, 512, 256, None, None, 0, 0, prec,
                      None, None, None, 'SZ', None)
        assert rc == 0, prec
    for prec in (L.GEMM_TF32, L.GEMM_TF32X3):
        rc, _ = _call(lib, 'gnnb200_gemm_f32', D, 256, 0, D, 256, 1, None, 512, 100_000, 512, 256, None, None, 0, 0, prec,
                      None, None, None, 'SZ', None)
        assert rc in (0, L.EUNSUPPORTED), prec


def test_too_small_workspace_is_refused(lib):
    have = c_size_t(16)
    rc = lib.gnnb200_dot_f32(D, D, 1 << 20, D, D, byref(have), None)
    assert rc == L.EWORKSPACE


def test_error_strings_cover_every_code(lib):
    for code in (L.OK, L.EINVAL, L.ERANGE, L.EWORKSPACE, L.EUNSUPPORTED, -5):
        s = lib.gnnb200_error_string(code)
        assert s and b'unknown' not in s
    assert b'unknown' in lib.gnnb200_error_string(-99)
