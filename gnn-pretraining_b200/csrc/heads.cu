// Elementwise tails and loss sums of the pre-training / fine-tuning heads (reference
// src/models/heads.py:35-50 `Linear -> ReLU -> Dropout`, src/pretrain/tasks.py:84,120,305,336 `mse_loss(sum)`,
// `binary_cross_entropy(sigmoid(.), sum)`, `cross_entropy(sum)`, src/models/heads.py:16-24 gradient reversal).
// The reference runs each of these as 2-6 eager kernels on tensors of a few thousand elements: a launch-bound regime.
// Here every loss is ONE launch forward (two beyond kLossSingleBlock elements) and one backward, deterministic (fixed
// summation order, no atomics):
//   act_dropout   y = drop(relu(x));  bwd gx = g * scale * [y > 0]     (the mask is recovered from y: no RNG replay)
//   sqdiff_sum    sum (a - b)^2;      bwd ga = 2 g (a - b)
//   sigmoid_bce   p = sigmoid(z), loss = -sum(t log p + (1 - t) log(1 - p)) with torch's clamp of the logs at -100;
//                 bwd gz = g (p - t) p (1 - p) / max(p (1 - p), 1e-12)   (torch's BCE backward times sigmoid')
//   ce_sum        sum_i (lse_i - z[i, t_i]);  bwd gz = g (softmax(z) - onehot(t))
//   scale         y = alpha x  (gradient reversal: alpha = -lambda)
#include "common.cuh"
#include "ew_common.cuh"

namespace gnnb200 {
namespace {

constexpr int kLossThreads = 1024;
constexpr int64_t kLossSingleBlock = 1 << 19;   // up to here one block walks everything: one launch, ~10 us
constexpr int kLossBlocks = kNumSMs * 2;

__device__ __forceinline__ float warp_sum_h(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the block, valid in thread 0; fixed order (lane tree, then warps in order)
__device__ __forceinline__ float block_sum(float v, float* smem) {
  v = warp_sum_h(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = (lane < (int)((blockDim.x + 31) >> 5)) ? smem[lane] : 0.f;
    r = warp_sum_h(r);
  }
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kLossThreads)
sum_finish_kernel(const float* __restrict__ partial, int count, float* __restrict__ out) {
  __shared__ float smem[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) acc += partial[i];
  const float r = block_sum(acc, smem);
  if (threadIdx.x == 0) *out = r;
}

// ---- ReLU + dropout ----------------------------------------------------------------------------------------
template <bool DROP>
__global__ void __launch_bounds__(256)
act_dropout_fwd_kernel(const float* __restrict__ x, int64_t n, int relu, uint64_t seed, uint32_t thresh, float scale,
                       float* __restrict__ y) {
  const int64_t n4 = (n + 3) >> 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
    bool k[4] = {true, true, true, true};
    if (DROP) keep4(seed, (uint64_t)q, thresh, k);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t e = (q << 2) + i;
      if (e < n) {
        float v = x[e];
        if (relu) v = fmaxf(v, 0.f);
        if (DROP) v = k[i] ? v * scale : 0.f;
        y[e] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
act_dropout_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t n, float scale,
                       float* __restrict__ gx) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    gx[e] = y[e] > 0.f ? g[e] * scale : 0.f;
}

__global__ void __launch_bounds__(256)
scale_kernel(const float* __restrict__ x, int64_t n, float alpha, float* __restrict__ y) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    y[e] = alpha * x[e];
}

// ---- squared difference ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLossThreads)
sqdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ out) {
  __shared__ float smem[32];
  float acc = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float d = a[e] - b[e];
    acc = fmaf(d, d, acc);
  }
  const float r = block_sum(acc, smem);
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}

__global__ void __launch_bounds__(256)
sqdiff_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g, int64_t n,
                  float* __restrict__ ga) {
  const float g2 = 2.f * *g;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    ga[e] = g2 * (a[e] - b[e]);
}

// ---- sigmoid + binary cross entropy on probabilities -----------------------------------------------------------
__global__ void __launch_bounds__(kLossThreads)
sigmoid_bce_fwd_kernel(const float* __restrict__ z, const float* __restrict__ t, int64_t n, float* __restrict__ p,
                       float* __restrict__ out) {
  __shared__ float smem[32];
  float acc = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float pe = 1.f / (1.f + expf(-z[e]));                 // torch.sigmoid
    p[e] = pe;
    const float lp = fmaxf(logf(pe), -100.f), lq = fmaxf(log1pf(-pe), -100.f);   // F.binary_cross_entropy's clamp
    const float te = t[e];
    acc -= te * lp + (1.f - te) * lq;
  }
  const float r = block_sum(acc, smem);
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}

__global__ void __launch_bounds__(256)
sigmoid_bce_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ g, int64_t n,
                       float* __restrict__ gz) {
  const float gg = *g;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float pe = p[e], v = pe * (1.f - pe);
    gz[e] = gg * (pe - t[e]) / fmaxf(v, 1e-12f) * v;
  }
}

// ---- cross entropy over rows -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLossThreads)
ce_sum_fwd_kernel(const float* __restrict__ z, int64_t ldz, const int64_t* __restrict__ target, int64_t rows, int cols,
                  float* __restrict__ lse, float* __restrict__ out) {
  __shared__ float smem[32];
  float acc = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float* row = z + r * ldz;
    float m = row[0];
    for (int c = 1; c < cols; ++c) m = fmaxf(m, row[c]);
    float s = 0.f;
    for (int c = 0; c < cols; ++c) s += expf(row[c] - m);
    const float l = m + logf(s);
    lse[r] = l;
    const int64_t tc = target[r];
    if (tc >= 0 && tc < cols) acc += l - row[tc];
  }
  const float r = block_sum(acc, smem);
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}

__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ z, int64_t ldz, const int64_t* __restrict__ target, const float* __restrict__ lse,
              const float* __restrict__ g, int64_t rows, int cols, float* __restrict__ gz, int64_t ldg) {
  const float gg = *g;
  const int64_t total = rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols;
    const int c = (int)(e - r * cols);
    const int64_t tc = target[r];
    const bool valid = tc >= 0 && tc < cols;
    const float sm = expf(z[r * ldz + c] - lse[r]);
    gz[r * ldg + c] = valid ? gg * (sm - (c == tc ? 1.f : 0.f)) : 0.f;
  }
}

inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// one block below kLossSingleBlock units of work (a single launch writes the result), a grid + finish beyond
inline int loss_blocks(int64_t n) { return n <= kLossSingleBlock ? 1 : kLossBlocks; }

}  // namespace
}  // namespace gnnb200

using namespace gnnb200;

#define GNNB200_LOSS_WS(n)                                       \
  Workspace ws(workspace);                                       \
  float* partial = ws.take<float>((size_t)kLossBlocks);          \
  if (!workspace) {                                              \
    *workspace_bytes = ws.bytes();                               \
    return GNNB200_OK;                                           \
  }                                                              \
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;  \
  const int blocks = loss_blocks(n);

#define GNNB200_LOSS_FINISH(out)                                                    \
  GNNB200_LAUNCH_CHECK();                                                           \
  if (blocks > 1) {                                                                 \
    sum_finish_kernel<<<1, kLossThreads, 0, stream>>>(partial, blocks, out);        \
    GNNB200_LAUNCH_CHECK();                                                         \
  }                                                                                 \
  return GNNB200_OK;

extern "C" int gnnb200_act_dropout_fwd_f32(const float* x, int64_t n, int relu, float drop_p, uint64_t seed, float* y,
                                           gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || !(drop_p >= 0.f && drop_p < 1.f)) return GNNB200_EINVAL;
  if (n == 0) return GNNB200_OK;
  if (!x || !y) return GNNB200_EINVAL;
  const double t = (double)(1.f - drop_p) * 4294967296.0;
  const uint32_t thresh = (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t);
  const float scale = 1.f / (1.f - drop_p);
  const int grid = ew_blocks((n + 3) / 4);
  if (drop_p > 0.f)
    act_dropout_fwd_kernel<true><<<grid, 256, 0, stream>>>(x, n, relu, seed, thresh, scale, y);
  else
    act_dropout_fwd_kernel<false><<<grid, 256, 0, stream>>>(x, n, relu, seed, thresh, scale, y);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_act_dropout_bwd_f32(const float* grad_y, const float* y, int64_t n, float drop_p, float* grad_x,
                                           gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || !(drop_p >= 0.f && drop_p < 1.f)) return GNNB200_EINVAL;
  if (n == 0) return GNNB200_OK;
  if (!grad_y || !y || !grad_x) return GNNB200_EINVAL;
  act_dropout_bwd_kernel<<<ew_blocks(n), 256, 0, stream>>>(grad_y, y, n, 1.f / (1.f - drop_p), grad_x);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_scale_f32(const float* x, int64_t n, float alpha, float* y, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0) return GNNB200_EINVAL;
  if (n == 0) return GNNB200_OK;
  if (!x || !y) return GNNB200_EINVAL;
  scale_kernel<<<ew_blocks(n), 256, 0, stream>>>(x, n, alpha, y);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_sqdiff_sum_f32(const float* a, const float* b, int64_t n, float* out, void* workspace,
                                      size_t* workspace_bytes, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || !workspace_bytes) return GNNB200_EINVAL;
  GNNB200_LOSS_WS(n)
  if (!out || (n > 0 && (!a || !b))) return GNNB200_EINVAL;
  sqdiff_sum_kernel<<<blocks, kLossThreads, 0, stream>>>(a, b, n, blocks > 1 ? partial : out);
  GNNB200_LOSS_FINISH(out)
}

extern "C" int gnnb200_sqdiff_bwd_f32(const float* a, const float* b, const float* grad_loss, int64_t n, float* grad_a,
                                      gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0) return GNNB200_EINVAL;
  if (n == 0) return GNNB200_OK;
  if (!a || !b || !grad_loss || !grad_a) return GNNB200_EINVAL;
  sqdiff_bwd_kernel<<<ew_blocks(n), 256, 0, stream>>>(a, b, grad_loss, n, grad_a);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_sigmoid_bce_fwd_f32(const float* logits, const float* labels, int64_t n, float* probs, float* loss,
                                           void* workspace, size_t* workspace_bytes, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || !workspace_bytes) return GNNB200_EINVAL;
  GNNB200_LOSS_WS(n)
  if (!loss || (n > 0 && (!logits || !labels || !probs))) return GNNB200_EINVAL;
  sigmoid_bce_fwd_kernel<<<blocks, kLossThreads, 0, stream>>>(logits, labels, n, probs, blocks > 1 ? partial : loss);
  GNNB200_LOSS_FINISH(loss)
}

extern "C" int gnnb200_sigmoid_bce_bwd_f32(const float* probs, const float* labels, const float* grad_loss, int64_t n,
                                           float* grad_logits, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0) return GNNB200_EINVAL;
  if (n == 0) return GNNB200_OK;
  if (!probs || !labels || !grad_loss || !grad_logits) return GNNB200_EINVAL;
  sigmoid_bce_bwd_kernel<<<ew_blocks(n), 256, 0, stream>>>(probs, labels, grad_loss, n, grad_logits);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_ce_sum_fwd_f32(const float* logits, int64_t ldz, const int64_t* target, int64_t rows, int64_t cols,
                                      float* lse, float* loss, void* workspace, size_t* workspace_bytes,
                                      gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows < 0 || cols <= 0 || !workspace_bytes) return GNNB200_EINVAL;
  if (cols > 65536) return GNNB200_ERANGE;
  GNNB200_LOSS_WS(rows * cols)
  if (!loss || (rows > 0 && (!logits || !target || !lse)) || ldz < cols) return GNNB200_EINVAL;
  ce_sum_fwd_kernel<<<blocks, kLossThreads, 0, stream>>>(logits, ldz, target, rows, (int)cols, lse, blocks > 1 ? partial : loss);
  GNNB200_LOSS_FINISH(loss)
}

extern "C" int gnnb200_ce_bwd_f32(const float* logits, int64_t ldz, const int64_t* target, const float* lse,
                                  const float* grad_loss, int64_t rows, int64_t cols, float* grad_logits, int64_t ldg,
                                  gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows < 0 || cols <= 0 || ldz < cols || ldg < cols) return GNNB200_EINVAL;
  if (rows == 0) return GNNB200_OK;
  if (!logits || !target || !lse || !grad_loss || !grad_logits) return GNNB200_EINVAL;
  ce_bwd_kernel<<<ew_blocks(rows * cols), 256, 0, stream>>>(logits, ldz, target, lse, grad_loss, rows, (int)cols, grad_logits, ldg);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}
