// One C call per GIN layer pass (reference src/models/gnn.py:26-43): the host-side sequence of kernel launches that
// gnnb200/fused.py otherwise issues from Python — gather(+self term) -> GEMM -> column statistics -> BN+ReLU ->
// GEMM(+residual) -> column statistics -> BN+ReLU+dropout, and its hand-written backward — issued from C++ on the
// caller's stream.  Nothing new runs on the device: every step is one of the public entry points of this library
// with exactly the arguments the Python path passes (tests/test_native_layer_trace.py compares the two call
// sequences argument by argument on CPU through the trace hook below).  What changes is the host time per layer:
// a step on small graph batches is ~3,000 launches of ~4 us kernels and the Python between them costs ~25 us per
// launch (scripts/host_profile.py); here a layer pass is one ctypes call.
#include <string.h>
#include <type_traits>
#include "common.cuh"

namespace gnnb200 {
namespace {

enum FnId { kAggregate = 0, kGemm, kColstats, kBnFinalize, kBnFwd, kBnBwd, kDot, kLinearX3w };

// Development trace: between gnnb200_dev_trace_begin/_end the composites RECORD their calls (function id, argument
// count, arguments as 64-bit words) instead of making them.  Per thread; lets a CPU-only test check the plumbing.
struct Trace {
  uint64_t* buf;
  size_t cap, len;
  bool on, overflow;
};
thread_local Trace g_trace = {nullptr, 0, 0, false, false};

template <class T>
uint64_t word(T v) {
  if constexpr (std::is_pointer_v<T>) {
    return (uint64_t)(uintptr_t)v;
  } else if constexpr (std::is_same_v<T, float>) {
    uint32_t b;
    memcpy(&b, &v, 4);
    return b;
  } else {
    return (uint64_t)(int64_t)v;
  }
}

template <class... P, class... A>
int call(int id, int (*fn)(P...), A... args) {
  static_assert(sizeof...(P) == sizeof...(A), "argument count");
  if (g_trace.on) {
    const uint64_t rec[] = {word<P>((P)args)...};
    const size_t n = sizeof...(P);
    if (g_trace.len + 2 + n > g_trace.cap) {
      g_trace.overflow = true;
      return GNNB200_OK;
    }
    g_trace.buf[g_trace.len++] = (uint64_t)id;
    g_trace.buf[g_trace.len++] = (uint64_t)n;
    for (size_t i = 0; i < n; ++i) g_trace.buf[g_trace.len++] = rec[i];
    return GNNB200_OK;
  }
  return fn((P)args...);
}

#define GNNB200_TRY(expr)        \
  do {                           \
    const int _rc = (expr);      \
    if (_rc) return _rc;         \
  } while (0)

struct Scratch {   // scratch of the inner entry points: all launches are on one stream, so one region serves them in turn
  size_t bytes = 0;
  void need(size_t b) { bytes = b > bytes ? b : bytes; }
};

int gemm_bytes(const gnnb200_gin_layer_t* a, const float* A, int64_t lda, int ta, const float* B, int64_t ldb, int tb,
               int64_t ldc, int64_t M, int64_t N, int64_t K, const float* residual, int64_t ldr, size_t* out) {
  return gnnb200_gemm_f32(A, lda, ta, B, ldb, tb, nullptr, ldc, M, N, K, nullptr, residual, ldr, GNNB200_EPI_NONE,
                          a->precision, nullptr, nullptr, nullptr, out, nullptr);
}

// forward GEMMs under GNNB200_GEMM_AUTO_FWD3 run on the pre-split weights
bool fwd3(const gnnb200_gin_layer_t* a) { return a->precision == GNNB200_GEMM_AUTO_FWD3; }

int linear_bytes(const gnnb200_gin_layer_t* a, const float* X, int64_t ldx, const float* W, int64_t ldw, int64_t ldy, int64_t M,
                 int64_t N, int64_t K, const float* residual, int64_t ldr, size_t* out) {
  if (fwd3(a))
    return gnnb200_linear_x3w_f32(X, ldx, W, nullptr, nullptr, ldw, nullptr, ldy, M, N, K, nullptr, residual, ldr,
                                  GNNB200_EPI_NONE, a->x3w_raw_hi, nullptr, nullptr, nullptr, out, nullptr);
  return gemm_bytes(a, X, ldx, 0, W, ldw, 1, ldy, M, N, K, residual, ldr, out);
}

bool bad_common(const gnnb200_gin_layer_t* a, const size_t* ws_bytes) {
  return !a || !ws_bytes || a->num_rows < 0 || a->hidden <= 0 || a->mid <= 0 || a->ldh < a->hidden;
}

}  // namespace
}  // namespace gnnb200

using namespace gnnb200;

extern "C" int gnnb200_dev_trace_begin(uint64_t* buf, size_t capacity_words) {
  if (!buf || capacity_words == 0) return GNNB200_EINVAL;
  g_trace = {buf, capacity_words, 0, true, false};
  return GNNB200_OK;
}

// number of words written, or -1 if the buffer was too small
extern "C" long long gnnb200_dev_trace_end(void) {
  const long long n = g_trace.overflow ? -1 : (long long)g_trace.len;
  g_trace = {nullptr, 0, 0, false, false};
  return n;
}

extern "C" int gnnb200_gin_layer_fwd_f32(const gnnb200_gin_layer_t* a, void* workspace, size_t* workspace_bytes,
                                         gnnb200_stream_t stream) {
  if (bad_common(a, workspace_bytes)) return GNNB200_EINVAL;
  const int64_t n = a->num_rows, C = a->hidden, H = a->mid;
  const bool train = a->training != 0;
  // ---- workspace: [sum1 H | m2_1 H | sum2 C | m2_2 C] (training) + the largest inner scratch
  Workspace ws(workspace);
  float* sum1 = train ? ws.take<float>((size_t)H) : nullptr;
  float* m21 = train ? ws.take<float>((size_t)H) : nullptr;
  float* sum2 = train ? ws.take<float>((size_t)C) : nullptr;
  float* m22 = train ? ws.take<float>((size_t)C) : nullptr;
  Scratch sc;
  size_t b = 0;
  GNNB200_TRY(linear_bytes(a, a->z, C, a->w1, C, H, n, H, C, nullptr, 0, &b));
  sc.need(b);
  GNNB200_TRY(linear_bytes(a, a->r1, H, a->w2, H, C, n, C, H, a->h, a->ldh, &b));
  sc.need(b);
  if (train) {
    GNNB200_TRY(gnnb200_colstats_f32(nullptr, H, n, H, nullptr, nullptr, nullptr, &b, nullptr));
    sc.need(b);
    GNNB200_TRY(gnnb200_colstats_f32(nullptr, C, n, C, nullptr, nullptr, nullptr, &b, nullptr));
    sc.need(b);
  }
  char* scratch = reinterpret_cast<char*>(ws.take<char>(sc.bytes));
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;
  if (!a->rowptr || !a->h || !a->w1 || !a->w2 || !a->gamma1 || !a->beta1 || !a->gamma2 || !a->beta2 || !a->z || !a->a1 ||
      !a->r1 || !a->s || !a->out || !a->mean1 || !a->invstd1 || !a->mean2 || !a->invstd2)
    return GNNB200_EINVAL;
  if (fwd3(a) && (!a->w1_lo || !a->w2_lo || (!a->x3w_raw_hi && (!a->w1_hi || !a->w2_hi)))) return GNNB200_EINVAL;
  if (n == 0) return GNNB200_OK;

  // z = A h + (1 + eps) h
  GNNB200_TRY(call(kAggregate, gnnb200_aggregate_f32, a->h, a->ldh, a->rowptr, a->col, n, C, GNNB200_AGG_SUM, a->h, a->ldh,
                   a->eps, nullptr, a->z, C, stream));
  // a1 = z W1^T + b1
  b = sc.bytes;
  if (fwd3(a))
    GNNB200_TRY(call(kLinearX3w, gnnb200_linear_x3w_f32, a->z, C, a->w1, a->w1_hi, a->w1_lo, C, a->a1, H, n, H, C, a->b1, nullptr,
                     0, GNNB200_EPI_NONE, a->x3w_raw_hi, nullptr, nullptr, scratch, &b, stream));
  else
  GNNB200_TRY(call(kGemm, gnnb200_gemm_f32, a->z, C, 0, a->w1, C, 1, a->a1, H, n, H, C, a->b1, nullptr, 0, GNNB200_EPI_NONE,
                   a->precision, nullptr, nullptr, scratch, &b, stream));
  if (train) {
    b = sc.bytes;
    GNNB200_TRY(call(kColstats, gnnb200_colstats_f32, a->a1, H, n, H, sum1, m21, scratch, &b, stream));
    GNNB200_TRY(call(kBnFinalize, gnnb200_bn_finalize_f32, sum1, m21, n, H, a->bn_eps1, a->momentum1, a->running_mean1,
                     a->running_var1, a->mean1, a->invstd1, stream));
  }
  // r1 = relu(bn1(a1))
  GNNB200_TRY(call(kBnFwd, gnnb200_bn_act_fwd_f32, a->a1, H, a->mean1, a->invstd1, a->gamma1, a->beta1, 1, 0.0f, 0, n, H,
                   a->r1, H, stream));
  // s = r1 W2^T + b2 + h
  b = sc.bytes;
  if (fwd3(a))
    GNNB200_TRY(call(kLinearX3w, gnnb200_linear_x3w_f32, a->r1, H, a->w2, a->w2_hi, a->w2_lo, H, a->s, C, n, C, H, a->b2, a->h,
                     a->ldh, GNNB200_EPI_NONE, a->x3w_raw_hi, nullptr, nullptr, scratch, &b, stream));
  else
  GNNB200_TRY(call(kGemm, gnnb200_gemm_f32, a->r1, H, 0, a->w2, H, 1, a->s, C, n, C, H, a->b2, a->h, a->ldh,
                   GNNB200_EPI_NONE, a->precision, nullptr, nullptr, scratch, &b, stream));
  if (train) {
    b = sc.bytes;
    GNNB200_TRY(call(kColstats, gnnb200_colstats_f32, a->s, C, n, C, sum2, m22, scratch, &b, stream));
    GNNB200_TRY(call(kBnFinalize, gnnb200_bn_finalize_f32, sum2, m22, n, C, a->bn_eps2, a->momentum2, a->running_mean2,
                     a->running_var2, a->mean2, a->invstd2, stream));
  }
  // out = dropout(relu(bn2(s)))
  GNNB200_TRY(call(kBnFwd, gnnb200_bn_act_fwd_f32, a->s, C, a->mean2, a->invstd2, a->gamma2, a->beta2, 1, a->drop_p, a->seed,
                   n, C, a->out, C, stream));
  return GNNB200_OK;
}

extern "C" int gnnb200_gin_layer_bwd_f32(const gnnb200_gin_layer_t* a, void* workspace, size_t* workspace_bytes,
                                         gnnb200_stream_t stream) {
  if (bad_common(a, workspace_bytes)) return GNNB200_EINVAL;
  if (!a->training) return GNNB200_EUNSUPPORTED;     // eval-mode backward (bias gradients by column sums): Python path
  const int64_t n = a->num_rows, C = a->hidden, H = a->mid;
  Workspace ws(workspace);
  Scratch sc;
  size_t b = 0;
  GNNB200_TRY(gnnb200_bn_act_bwd_f32(nullptr, C, nullptr, C, nullptr, nullptr, nullptr, nullptr, 1, a->drop_p, a->seed, 1, 0, n,
                                     n, C, nullptr, C, nullptr, nullptr, nullptr, &b, nullptr));
  sc.need(b);
  GNNB200_TRY(gnnb200_bn_act_bwd_f32(nullptr, H, nullptr, H, nullptr, nullptr, nullptr, nullptr, 1, 0.0f, 0, 1, 0, n, n, H,
                                     nullptr, H, nullptr, nullptr, nullptr, &b, nullptr));
  sc.need(b);
  GNNB200_TRY(gemm_bytes(a, a->ds, C, 1, a->r1, H, 0, H, C, H, n, nullptr, 0, &b));      // dW2 = ds^T r1
  sc.need(b);
  GNNB200_TRY(gemm_bytes(a, a->ds, C, 0, a->w2, H, 0, H, n, H, C, nullptr, 0, &b));      // dr1 = ds W2
  sc.need(b);
  GNNB200_TRY(gemm_bytes(a, a->da1, H, 1, a->z, C, 0, C, H, C, n, nullptr, 0, &b));      // dW1 = da1^T z
  sc.need(b);
  GNNB200_TRY(gemm_bytes(a, a->da1, H, 0, a->w1, C, 0, C, n, C, H, nullptr, 0, &b));     // dz = da1 W1
  sc.need(b);
  if (a->deps) {
    GNNB200_TRY(gnnb200_dot_f32(nullptr, nullptr, n * C, nullptr, nullptr, &b, nullptr));
    sc.need(b);
  }
  char* scratch = reinterpret_cast<char*>(ws.take<char>(sc.bytes));
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;
  if (!a->grad_out || !a->s || !a->a1 || !a->r1 || !a->z || !a->w1 || !a->w2 || !a->gamma1 || !a->beta1 || !a->gamma2 ||
      !a->beta2 || !a->mean1 || !a->invstd1 || !a->mean2 || !a->invstd2 || !a->ds || !a->dr1 || !a->da1 || !a->dz ||
      !a->dw1 || !a->dw2 || !a->dgamma1 || !a->dbeta1 || !a->dgamma2 || !a->dbeta2)
    return GNNB200_EINVAL;
  if ((a->deps && (!a->h || a->ldh != C)) || (a->need_dh && !a->rowptr && n > 0)   /* col may be NULL: a graph without edges */) return GNNB200_EINVAL;
  if (n == 0) return GNNB200_OK;

  // ds = bn2'(grad_out)   (recomputes x_hat, the ReLU sign and the dropout mask from s)
  b = sc.bytes;
  GNNB200_TRY(call(kBnBwd, gnnb200_bn_act_bwd_f32, a->grad_out, C, a->s, C, a->mean2, a->invstd2, a->gamma2, a->beta2, 1,
                   a->drop_p, a->seed, 1, 0, n, n, C, a->ds, C, a->dgamma2, a->dbeta2, scratch, &b, stream));
  b = sc.bytes;
  GNNB200_TRY(call(kGemm, gnnb200_gemm_f32, a->ds, C, 1, a->r1, H, 0, a->dw2, H, C, H, n, nullptr, nullptr, 0,
                   GNNB200_EPI_NONE, a->precision, nullptr, nullptr, scratch, &b, stream));
  b = sc.bytes;
  GNNB200_TRY(call(kGemm, gnnb200_gemm_f32, a->ds, C, 0, a->w2, H, 0, a->dr1, H, n, H, C, nullptr, nullptr, 0,
                   GNNB200_EPI_NONE, a->precision, nullptr, nullptr, scratch, &b, stream));
  // da1 = bn1'(dr1)
  b = sc.bytes;
  GNNB200_TRY(call(kBnBwd, gnnb200_bn_act_bwd_f32, a->dr1, H, a->a1, H, a->mean1, a->invstd1, a->gamma1, a->beta1, 1, 0.0f, 0,
                   1, 0, n, n, H, a->da1, H, a->dgamma1, a->dbeta1, scratch, &b, stream));
  b = sc.bytes;
  GNNB200_TRY(call(kGemm, gnnb200_gemm_f32, a->da1, H, 1, a->z, C, 0, a->dw1, C, H, C, n, nullptr, nullptr, 0,
                   GNNB200_EPI_NONE, a->precision, nullptr, nullptr, scratch, &b, stream));
  b = sc.bytes;
  GNNB200_TRY(call(kGemm, gnnb200_gemm_f32, a->da1, H, 0, a->w1, C, 0, a->dz, C, n, C, H, nullptr, nullptr, 0,
                   GNNB200_EPI_NONE, a->precision, nullptr, nullptr, scratch, &b, stream));
  if (a->deps) {   // d(eps) = <dz, h>
    b = sc.bytes;
    GNNB200_TRY(call(kDot, gnnb200_dot_f32, a->dz, a->h, n * C, a->deps, scratch, &b, stream));
  }
  if (a->need_dh) {   // dh = ds + A^T dz + (1 + eps) dz, accumulated into ds's buffer
    GNNB200_TRY(call(kAggregate, gnnb200_aggregate_f32, a->dz, C, a->rowptr, a->col, n, C,
                     GNNB200_AGG_SUM | GNNB200_AGG_ACCUMULATE, a->dz, C, a->eps, nullptr, a->ds, C, stream));
  }
  return GNNB200_OK;
}
