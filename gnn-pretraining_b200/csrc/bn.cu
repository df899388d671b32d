// Fused BatchNorm1d (+ReLU)(+dropout) over node rows (K4/K5 of SURVEY §2.4): the elementwise tail of
// InputEncoder and GINLayer (reference src/models/gnn.py:19-22, :31-32, :41-43).  The reference runs
// batch_norm, relu and dropout as separate passes that each read and write the [N, C] activation and
// save their own copy for autograd; here one forward pass reads the pre-BN activation once and writes the
// layer output once, and the backward recomputes x_hat / ReLU sign / dropout mask from the saved pre-BN
// activation (nothing else is stored):
//   finalize : (sum, m2, n) -> mean, invstd = rsqrt(m2/n + eps); running stats updated like torch
//   fwd      : y = drop(relu((x - mean) * invstd * gamma + beta))
//   bwd      : g1 = g * drop' * relu';  dbeta = sum g1;  dgamma = sum g1 * xhat          (two-stage, fixed order)
//              dx = gamma * invstd * (g1 - dbeta/n - xhat * dgamma/n)
// Dropout uses a counter-based Philox4x32-10 stream keyed by (seed, element index / 4): the same mask is
// regenerated in the backward pass.  It is Bernoulli(1-p) with 1/(1-p) scaling like torch's, but not the
// same bit stream as torch's CUDA generator (the reference's CPU and CUDA streams differ from each other too).
#include "common.cuh"
#include "ew_common.cuh"

namespace gnnb200 {

// elementwise_v2.cu: the column-stationary kernels every call takes (measured on B200 at C5 size: apply C=512 5.1 -> 4.0 ms,
// forward with dropout 2.3 -> 1.8 ms); GNNB200_EUNSUPPORTED = more than 65535 x 256 rows, which fall through to the
// flat grid-stride kernels below
int bn_act_fwd_v2(const float* x, int64_t ldx, const float* mean, const float* invstd, const float* gamma,
                  const float* beta, int relu, bool drop, uint64_t seed, uint32_t thresh, float scale, int64_t rows,
                  int64_t cols, float* y, int64_t ldy, cudaStream_t stream);
int bn_act_bwd_apply_v2(const float* g, int64_t ldg, const float* x, int64_t ldx, const float* mean, const float* invstd,
                        const float* gamma, const float* beta, const float* dgamma, const float* dbeta, int relu,
                        int training, bool drop, uint64_t seed, uint32_t thresh, float scale, int64_t rows,
                        int64_t rows_total, int64_t cols, float* dx, int64_t lddx, cudaStream_t stream);

__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ m2, float n, float eps,
                                   float momentum, int cols, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean,
                                   float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const float mu = sum[c] / n;
  const float var = m2[c] / n;
  mean[c] = mu;
  invstd[c] = rsqrtf(var + eps);
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
  if (running_var) {
    const float unbiased = n > 1.f ? m2[c] / (n - 1.f) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// Node-partitioned BatchNorm: every rank's (n, sum, centred m2) per column, all-gathered into [parts][3][cols], merged in
// rank order with Chan's formula (same inputs in the same order on every rank => identical bits everywhere) and
// finalised like bn_finalize_kernel — one launch instead of the ~15 elementwise launches of the torch expression.
__global__ void bn_merge_finalize_kernel(const float* __restrict__ moments, int parts, float eps, float momentum, int cols,
                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                         float* __restrict__ mean, float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float n = 0.f, sum = 0.f;
  for (int r = 0; r < parts; ++r) {
    n += moments[((int64_t)r * 3 + 0) * cols + c];
    sum += moments[((int64_t)r * 3 + 1) * cols + c];
  }
  const float mu = sum / fmaxf(n, 1.f);
  float m2 = 0.f;
  for (int r = 0; r < parts; ++r) {
    const float nr = moments[((int64_t)r * 3 + 0) * cols + c];
    const float d = moments[((int64_t)r * 3 + 1) * cols + c] / fmaxf(nr, 1.f) - mu;
    m2 += moments[((int64_t)r * 3 + 2) * cols + c] + nr * d * d;
  }
  const float var = m2 / fmaxf(n, 1.f);
  mean[c] = mu;
  invstd[c] = rsqrtf(var + eps);
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
  if (running_var) {
    const float unbiased = n > 1.f ? m2 / (n - 1.f) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// One thread handles 4 consecutive columns of one row (cols % 4 == 0); grid-stride over rows*cols/4.
template <bool DROP>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ mean,
                  const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                  int relu, uint64_t seed, uint32_t thresh, float scale, int64_t rows, int cols,
                  float* __restrict__ y, int64_t ldy) {
  const int c4 = cols >> 2;
  const int64_t total = rows * c4;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = q / c4;
    const int c = (int)(q - r * c4) << 2;
    const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    float o[4] = {(v.x - mu.x) * is.x * ga.x + be.x, (v.y - mu.y) * is.y * ga.y + be.y,
                  (v.z - mu.z) * is.z * ga.z + be.z, (v.w - mu.w) * is.w * ga.w + be.w};
    if (relu) {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = fmaxf(o[i], 0.f);
    }
    if (DROP) {
      bool k[4];
      keep4(seed, (uint64_t)q, thresh, k);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = k[i] ? o[i] * scale : 0.f;
    }
    *reinterpret_cast<float4*>(y + r * ldy + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

constexpr int kBnRows = 256;

// Block = 32 column-quads (128 columns) x 8 row lanes over kBnRows rows; partial [chunks][2][cols].
template <bool DROP>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x, int64_t ldx,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ gamma, const float* __restrict__ beta, int relu, uint64_t seed,
                         uint32_t thresh, float scale, int64_t rows, int cols, float* __restrict__ part) {
  __shared__ float sm[2][8][128];
  const int c4 = cols >> 2;
  const int cq = blockIdx.x * 32 + threadIdx.x;     // column quad
  const int c = cq << 2;
  const int64_t r0 = (int64_t)blockIdx.y * kBnRows;
  const int64_t r1 = min(rows, r0 + kBnRows);
  float sb[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {0.f, 0.f, 0.f, 0.f};
  if (cq < c4) {
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float4 gv = *reinterpret_cast<const float4*>(g + r * ldg + c);
      const float4 xv = *reinterpret_cast<const float4*>(x + r * ldx + c);
      float g1[4], xh[4];
      bn_g1<DROP>(gv, xv, mu, is, ga, be, relu, seed, (uint64_t)(r * c4 + cq), thresh, scale, g1, xh);
#pragma unroll
      for (int i = 0; i < 4; ++i) { sb[i] += g1[i]; sg[i] = fmaf(g1[i], xh[i], sg[i]); }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sm[0][threadIdx.y][threadIdx.x * 4 + i] = sb[i];
    sm[1][threadIdx.y][threadIdx.x * 4 + i] = sg[i];
  }
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;       // 256 threads -> 2 x 128 outputs
  const int which = t >> 7, col = t & 127;
  if (blockIdx.x * 128 + col < cols) {
    float acc = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) acc += sm[which][l][col];
    part[((int64_t)blockIdx.y * 2 + which) * cols + blockIdx.x * 128 + col] = acc;
  }
}

// Fold the chunk partials: 32 columns x 32 lanes per block, lanes stride the chunks in ascending order.
__global__ void __launch_bounds__(1024)
bn_act_bwd_finish_kernel(const float* __restrict__ part, int chunks, int cols, float* __restrict__ dgamma,
                         float* __restrict__ dbeta) {
  __shared__ float sm[2][32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float b = 0.f, gsum = 0.f;
  if (c < cols) {
    for (int k = threadIdx.y; k < chunks; k += 32) {
      b += part[((int64_t)k * 2 + 0) * cols + c];
      gsum += part[((int64_t)k * 2 + 1) * cols + c];
    }
  }
  sm[0][threadIdx.y][threadIdx.x] = b;
  sm[1][threadIdx.y][threadIdx.x] = gsum;
  __syncthreads();
  if (threadIdx.y < 2 && c < cols) {
    float acc = 0.f;
#pragma unroll
    for (int l = 0; l < 32; ++l) acc += sm[threadIdx.y][l][threadIdx.x];
    (threadIdx.y == 0 ? dbeta : dgamma)[c] = acc;
  }
}

template <bool DROP>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x, int64_t ldx,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        const float* __restrict__ dgamma, const float* __restrict__ dbeta, int relu, int training,
                        uint64_t seed, uint32_t thresh, float scale, int64_t rows, int64_t rows_total, int cols,
                        float* __restrict__ dx, int64_t lddx) {
  const int c4 = cols >> 2;
  const int64_t total = rows * c4;
  const float inv_n = 1.f / (float)rows_total;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = q / c4;
    const int cq = (int)(q - r * c4);
    const int c = cq << 2;
    const float4 gv = *reinterpret_cast<const float4*>(g + r * ldg + c);
    const float4 xv = *reinterpret_cast<const float4*>(x + r * ldx + c);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
    float g1[4], xh[4];
    bn_g1<DROP>(gv, xv, mu, is, ga, be, relu, seed, (uint64_t)q, thresh, scale, g1, xh);
    const float isv[4] = {is.x, is.y, is.z, is.w};
    const float gav[4] = {ga.x, ga.y, ga.z, ga.w};
    float o[4];
    if (training) {
      const float4 dg = __ldg(reinterpret_cast<const float4*>(dgamma + c));
      const float4 db = __ldg(reinterpret_cast<const float4*>(dbeta + c));
      const float dgv[4] = {dg.x, dg.y, dg.z, dg.w};
      const float dbv[4] = {db.x, db.y, db.z, db.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = gav[i] * isv[i] * (g1[i] - dbv[i] * inv_n - xh[i] * dgv[i] * inv_n);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = gav[i] * isv[i] * g1[i];     // eval mode: statistics are constants
    }
    *reinterpret_cast<float4*>(dx + r * lddx + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

static inline int ew_grid(int64_t work) {
  int64_t g = (work + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

static inline bool bn_layout_ok(int64_t cols, int64_t lda, int64_t ldb, const void* a, const void* b) {
  return cols % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0);
}

}  // namespace gnnb200

using namespace gnnb200;

extern "C" int gnnb200_bn_finalize_f32(const float* sum, const float* m2, int64_t rows, int64_t cols, float eps,
                                       float momentum, float* running_mean, float* running_var, float* mean,
                                       float* invstd, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows <= 0 || cols < 0) return GNNB200_EINVAL;
  if (cols == 0) return GNNB200_OK;
  if (!sum || !m2 || !mean || !invstd) return GNNB200_EINVAL;
  bn_finalize_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, stream>>>(sum, m2, (float)rows, eps, momentum, (int)cols,
                                                                        running_mean, running_var, mean, invstd);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_bn_merge_finalize_f32(const float* moments, int64_t parts, int64_t cols, float eps, float momentum,
                                             float* running_mean, float* running_var, float* mean, float* invstd,
                                             gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (parts <= 0 || cols < 0 || parts > 4096) return GNNB200_EINVAL;
  if (cols == 0) return GNNB200_OK;
  if (!moments || !mean || !invstd) return GNNB200_EINVAL;
  bn_merge_finalize_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, stream>>>(moments, (int)parts, eps, momentum, (int)cols,
                                                                              running_mean, running_var, mean, invstd);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_bn_act_fwd_f32(const float* x, int64_t ldx, const float* mean, const float* invstd,
                                      const float* gamma, const float* beta, int relu, float drop_p, uint64_t seed,
                                      int64_t rows, int64_t cols, float* y, int64_t ldy, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows < 0 || cols < 0 || !(drop_p >= 0.f && drop_p < 1.f)) return GNNB200_EINVAL;
  if (rows == 0 || cols == 0) return GNNB200_OK;
  if (!x || !mean || !invstd || !gamma || !beta || !y) return GNNB200_EINVAL;
  if (!bn_layout_ok(cols, ldx, ldy, x, y)) return GNNB200_EUNSUPPORTED;
  const uint32_t thresh = (uint32_t)((double)(1.f - drop_p) * 4294967296.0 > 4294967295.0 ? 4294967295.0 : (double)(1.f - drop_p) * 4294967296.0);
  const float scale = 1.f / (1.f - drop_p);
  {
    const int rc = bn_act_fwd_v2(x, ldx, mean, invstd, gamma, beta, relu, drop_p > 0.f, seed, thresh, scale, rows, cols, y, ldy, stream);
    if (rc != GNNB200_EUNSUPPORTED) return rc;
  }
  const int grid = ew_grid(rows * (cols / 4));
  if (drop_p > 0.f)
    bn_act_fwd_kernel<true><<<grid, 256, 0, stream>>>(x, ldx, mean, invstd, gamma, beta, relu, seed, thresh, scale, rows, (int)cols, y, ldy);
  else
    bn_act_fwd_kernel<false><<<grid, 256, 0, stream>>>(x, ldx, mean, invstd, gamma, beta, relu, seed, thresh, scale, rows, (int)cols, y, ldy);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_bn_act_bwd_f32(const float* grad_y, int64_t ldg, const float* x, int64_t ldx, const float* mean,
                                      const float* invstd, const float* gamma, const float* beta, int relu,
                                      float drop_p, uint64_t seed, int training, int phase, int64_t rows,
                                      int64_t rows_total, int64_t cols, float* grad_x, int64_t ldgx, float* dgamma,
                                      float* dbeta, void* workspace, size_t* workspace_bytes,
                                      gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows < 0 || cols < 0 || !workspace_bytes || !(drop_p >= 0.f && drop_p < 1.f) || phase < 0 || phase > 2 ||
      rows_total < rows)
    return GNNB200_EINVAL;
  const int64_t chunks = rows > 0 ? (rows + kBnRows - 1) / kBnRows : 1;
  if (chunks > 65535) return GNNB200_ERANGE;
  Workspace ws(workspace);
  float* part = ws.take<float>((size_t)chunks * 2 * cols);
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;
  if (rows == 0 || cols == 0) return GNNB200_OK;
  if (!grad_y || !x || !mean || !invstd || !gamma || !beta || !dgamma || !dbeta || (phase != 1 && !grad_x)) return GNNB200_EINVAL;
  if (!bn_layout_ok(cols, ldg, ldx, grad_y, x) || (phase != 1 && (ldgx % 4 != 0 || ((uintptr_t)grad_x % 16) != 0))) return GNNB200_EUNSUPPORTED;
  const uint32_t thresh = (uint32_t)((double)(1.f - drop_p) * 4294967296.0 > 4294967295.0 ? 4294967295.0 : (double)(1.f - drop_p) * 4294967296.0);
  const float scale = 1.f / (1.f - drop_p);
  const bool drop = drop_p > 0.f;
  dim3 rgrid((unsigned)((cols + 127) / 128), (unsigned)chunks);
  dim3 rblock(32, 8);
  if (phase == 2) goto apply;
  if (drop)
    bn_act_bwd_reduce_kernel<true><<<rgrid, rblock, 0, stream>>>(grad_y, ldg, x, ldx, mean, invstd, gamma, beta, relu, seed, thresh, scale, rows, (int)cols, part);
  else
    bn_act_bwd_reduce_kernel<false><<<rgrid, rblock, 0, stream>>>(grad_y, ldg, x, ldx, mean, invstd, gamma, beta, relu, seed, thresh, scale, rows, (int)cols, part);
  GNNB200_LAUNCH_CHECK();
  bn_act_bwd_finish_kernel<<<(unsigned)((cols + 31) / 32), dim3(32, 32), 0, stream>>>(part, (int)chunks, (int)cols, dgamma, dbeta);
  GNNB200_LAUNCH_CHECK();
  if (phase == 1) return GNNB200_OK;
apply : {
  {
    const int rc = bn_act_bwd_apply_v2(grad_y, ldg, x, ldx, mean, invstd, gamma, beta, dgamma, dbeta, relu, training, drop, seed,
                                       thresh, scale, rows, rows_total, cols, grad_x, ldgx, stream);
    if (rc != GNNB200_EUNSUPPORTED) return rc;
  }
  const int grid = ew_grid(rows * (cols / 4));
  if (drop)
    bn_act_bwd_apply_kernel<true><<<grid, 256, 0, stream>>>(grad_y, ldg, x, ldx, mean, invstd, gamma, beta, dgamma, dbeta, relu, training, seed, thresh, scale, rows, rows_total, (int)cols, grad_x, ldgx);
  else
    bn_act_bwd_apply_kernel<false><<<grid, 256, 0, stream>>>(grad_y, ldg, x, ldx, mean, invstd, gamma, beta, dgamma, dbeta, relu, training, seed, thresh, scale, rows, rows_total, (int)cols, grad_x, ldgx);
  GNNB200_LAUNCH_CHECK();
}
  return GNNB200_OK;
}
