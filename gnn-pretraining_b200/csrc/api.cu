// C-ABI glue: version/error strings and the GEMM dispatcher (include/gnnb200.h).
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
