// C-ABI glue: version/error strings and the GEMM dispatcher (include/gnnb200.h).
#include <stdlib.h>
#include "common.cuh"

namespace gnnb200 {
int gemm_simt(const float* A, int64_t lda, int transa, const float* B, int64_t ldb, int transb, float* C,
              int64_t ldc, int64_t M, int64_t N, int64_t K, const float* bias, const float* residual, int64_t ldr,
              int epilogue, void* workspace, size_t* workspace_bytes, cudaStream_t stream);
int gemm_tf32_supported(const float* A, int64_t lda, int transa, const float* B, int64_t ldb, int transb,
                        const float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const float* residual,
                        int64_t ldr);
int gemm_tf32(const float* A, int64_t lda, int transa, const float* B, const float* B_lo, int64_t ldb, int transb, float* C,
              int64_t ldc, int64_t M, int64_t N, int64_t K, const float* bias, const float* residual, int64_t ldr,
              int epilogue, int x3, float* col_sum, float* col_m2, void* workspace, size_t* workspace_bytes,
              cudaStream_t stream);
int split_tf32(const float* x, int64_t n, float* hi, float* lo, cudaStream_t stream);
}  // namespace gnnb200

extern "C" int gnnb200_version(void) { return 100; }

extern "C" const char* gnnb200_error_string(int code) {
  switch (code) {
    case GNNB200_OK: return "ok";
    case GNNB200_EINVAL: return "invalid argument";
    case GNNB200_ERANGE: return "size does not fit the int32 CSR / grid limits";
    case GNNB200_EWORKSPACE: return "workspace too small";
    case GNNB200_EUNSUPPORTED: return "shape or alignment not supported by this kernel";
    case GNNB200_ETMA: return "TMA tensor-map encode failed";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown gnnb200 error";
  }
}

extern "C" int gnnb200_gemm_f32(const float* A, int64_t lda, int transa, const float* B, int64_t ldb, int transb,
                                float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const float* bias,
                                const float* residual, int64_t ldr, int epilogue, int precision, float* col_sum,
                                float* col_m2, void* workspace, size_t* workspace_bytes, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (M < 0 || N < 0 || K < 0 || !workspace_bytes) return GNNB200_EINVAL;
  if (M >= (int64_t)INT32_MAX || N >= (int64_t)INT32_MAX) return GNNB200_ERANGE;
  if (workspace && M > 0 && N > 0 && (!C || (K > 0 && (!A || !B)))) return GNNB200_EINVAL;
  const bool want_stats = col_sum != nullptr || col_m2 != nullptr;
  // FFMA path: the column statistics are a second pass over C (gnnb200_colstats_f32) sharing the caller's workspace
  auto simt_with_stats = [&]() -> int {
    size_t gemm_bytes = 0, stat_bytes = 0;
    int rc = gnnb200::gemm_simt(A, lda, transa, B, ldb, transb, C, ldc, M, N, K, bias, residual, ldr, epilogue, nullptr,
                                &gemm_bytes, stream);
    if (rc) return rc;
    if (want_stats) {
      rc = gnnb200_colstats_f32(C, ldc, M, N, col_sum, col_m2, nullptr, &stat_bytes, stream_);
      if (rc) return rc;
    }
    const size_t need = gemm_bytes > stat_bytes ? gemm_bytes : stat_bytes;
    if (!workspace) {
      *workspace_bytes = need;
      return GNNB200_OK;
    }
    if (*workspace_bytes < need) return GNNB200_EWORKSPACE;
    size_t b = *workspace_bytes;
    rc = gnnb200::gemm_simt(A, lda, transa, B, ldb, transb, C, ldc, M, N, K, bias, residual, ldr, epilogue, workspace, &b,
                            stream);
    if (rc || !want_stats) return rc;
    b = *workspace_bytes;
    return gnnb200_colstats_f32(C, ldc, M, N, col_sum, col_m2, workspace, &b, stream_);
  };
  if (precision == GNNB200_GEMM_AUTO_FWD3) precision = GNNB200_GEMM_AUTO;
  if (precision == GNNB200_GEMM_F32) return simt_with_stats();
  if (precision == GNNB200_GEMM_TF32 || precision == GNNB200_GEMM_AUTO || precision == GNNB200_GEMM_TF32X3 ||
      precision == GNNB200_GEMM_AUTO_X3) {
    const int x3 = (precision == GNNB200_GEMM_TF32X3 || precision == GNNB200_GEMM_AUTO_X3) ? 1 : 0;
    // the support predicate only reads sizes, leading dimensions and pointer alignment; during the
    // workspace query C may be NULL (NULL is 16-byte aligned), so both phases take the same branch
    if (!gnnb200::gemm_tf32_supported(A, lda, transa, B, ldb, transb, C, ldc, M, N, K, residual, ldr)) {
      if (precision == GNNB200_GEMM_TF32 || precision == GNNB200_GEMM_TF32X3) return GNNB200_EUNSUPPORTED;
      return simt_with_stats();
    }
    return gnnb200::gemm_tf32(A, lda, transa, B, nullptr, ldb, transb, C, ldc, M, N, K, bias, residual, ldr, epilogue, x3,
                              col_sum, col_m2, workspace, workspace_bytes, stream);
  }
  return GNNB200_EINVAL;
}

extern "C" int gnnb200_split_tf32_f32(const float* x, int64_t n, float* hi, float* lo, gnnb200_stream_t stream) {
  if (n < 0 || (n > 0 && (!x || !hi || !lo))) return GNNB200_EINVAL;
  return gnnb200::split_tf32(x, n, hi, lo, (cudaStream_t)stream);
}

extern "C" int gnnb200_linear_x3w_f32(const float* X, int64_t ldx, const float* W, const float* W_hi, const float* W_lo,
                                      int64_t ldw, float* Y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                                      const float* bias, const float* residual, int64_t ldr, int epilogue, int raw_hi,
                                      float* col_sum, float* col_m2, void* workspace, size_t* workspace_bytes,
                                      gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (M < 0 || N < 0 || K < 0 || !workspace_bytes) return GNNB200_EINVAL;
  if (M >= (int64_t)INT32_MAX || N >= (int64_t)INT32_MAX) return GNNB200_ERANGE;
  if (workspace && M > 0 && N > 0 && (!Y || (K > 0 && (!X || !W || !W_lo || (!raw_hi && !W_hi))))) return GNNB200_EINVAL;
  // layouts TMA cannot express (ld % 4 != 0, N = 1, ...) take the fp32 FFMA kernel on the unsplit weights
  if (!gnnb200::gemm_tf32_supported(X, ldx, 0, W, ldw, 1, Y, ldy, M, N, K, residual, ldr) ||
      ((uintptr_t)W_lo & 15) || (!raw_hi && ((uintptr_t)W_hi & 15)))
    return gnnb200_gemm_f32(X, ldx, 0, W, ldw, 1, Y, ldy, M, N, K, bias, residual, ldr, epilogue, GNNB200_GEMM_F32, col_sum,
                            col_m2, workspace, workspace_bytes, stream_);
  return gnnb200::gemm_tf32(X, ldx, 0, raw_hi ? W : W_hi, W_lo, ldw, 1, Y, ldy, M, N, K, bias, residual, ldr, epilogue,
                            raw_hi ? 3 : 2, col_sum, col_m2, workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// Host helper (no device work): random.sample(range(n), k) of CPython 3.x on a COPY of the interpreter's Mersenne
// Twister state, so that negative sampling's random branch (PyG negative_sampling -> random.sample)
// draws exactly the reference's numbers without ~1 us of interpreter time per draw.  Follows Lib/random.py:
//   _randbelow_with_getrandbits(m): b = m.bit_length(); r = getrandbits(b); while r >= m: r = getrandbits(b)
//   getrandbits(b <= 32) = genrand_uint32() >> (32 - b)
//   sample(): setsize = 21 (+ 4 ** ceil(log(3k, 4)) if k > 5); n <= setsize -> pool with swap-removal,
//             else rejection against the set of already selected values.
// mt[624] / *pos are Python's random.getstate()[1]; both are advanced in place.  Returns 0, or GNNB200_EINVAL /
// GNNB200_ERANGE (n >= 2^32: the caller keeps using the interpreter).
// ------------------------------------------------------------------------------------------------------------------
namespace {
struct MT {
  uint32_t* mt;
  int pos;
  uint32_t next() {
    if (pos >= 624) {
      static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
      int kk;
      uint32_t y;
      for (kk = 0; kk < 624 - 397; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 0x1u];
      }
      for (; kk < 623; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 0x1u];
      }
      y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
      mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 0x1u];
      pos = 0;
    }
    uint32_t y = mt[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  uint64_t below(uint64_t m) {               // m >= 1, m < 2^32
    int bits = 0;
    for (uint64_t t = m; t; t >>= 1) ++bits;
    uint32_t r = next() >> (32 - bits);
    while (r >= m) r = next() >> (32 - bits);
    return r;
  }
};
}  // namespace

extern "C" int gnnb200_host_py_sample_range(uint32_t* mt624, int32_t* pos, int64_t n, int64_t k, int64_t* out) {
  if (!mt624 || !pos || n < 0 || k < 0 || k > n || (k > 0 && !out) || *pos < 0 || *pos > 624) return GNNB200_EINVAL;
  if (n >= (1LL << 32)) return GNNB200_ERANGE;
  if (k == 0) return GNNB200_OK;
  MT g{mt624, *pos};
  double setsize = 21.0;
  if (k > 5) {
    // 4 ** ceil(log(3k, 4)) with Python's float log: the smallest power of 4 that is >= 3k, except where the
    // float quotient log(3k)/log(4) lands above an exact integer (it does not for 3k < 2^53: checked by the tests)
    int e = 0;
    long double p = 1.0L;
    while (p < (long double)(3 * k)) { p *= 4.0L; ++e; }
    (void)e;
    setsize += (double)p;
  }
  if ((double)n <= setsize) {
    int64_t* pool = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    if (!pool) return GNNB200_EINVAL;
    for (int64_t i = 0; i < n; ++i) pool[i] = i;
    for (int64_t i = 0; i < k; ++i) {
      const uint64_t j = g.below((uint64_t)(n - i));
      out[i] = pool[j];
      pool[j] = pool[n - i - 1];
    }
    free(pool);
  } else {
    const size_t words = ((size_t)n + 63) / 64;
    uint64_t* seen = (uint64_t*)calloc(words, sizeof(uint64_t));
    if (!seen) return GNNB200_EINVAL;
    for (int64_t i = 0; i < k; ++i) {
      uint64_t j = g.below((uint64_t)n);
      while (seen[j >> 6] & (1ull << (j & 63))) j = g.below((uint64_t)n);
      seen[j >> 6] |= 1ull << (j & 63);
      out[i] = (int64_t)j;
    }
    free(seen);
  }
  *pos = g.pos;
  return GNNB200_OK;
}
