// Deterministic reductions: dot product (d eps of GINConv, reference src/models/gnn.py:29-37 via
// autograd) and per-column statistics (BatchNorm1d batch statistics, reference
// src/models/gnn.py:15,32,38; bias gradients of nn.Linear).  Two-stage, fixed order, no atomics.
#include "common.cuh"
#include "ew_common.cuh"

namespace gnnb200 {

constexpr int kDotBlocks = kNumSMs * 4;
constexpr int kDotThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float block_sum_256(float v, float* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

__global__ void __launch_bounds__(kDotThreads)
dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ partial) {
  __shared__ float smem[8];
  float acc = 0.f;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec = (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
  if (vec) {
    const int64_t n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 u = __ldg(a4 + i), v = __ldg(b4 + i);
      acc += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) acc += a[i] * b[i];
  } else {
    for (int64_t i = tid; i < n; i += stride) acc += a[i] * b[i];
  }
  const float r = block_sum_256(acc, smem);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

__global__ void __launch_bounds__(kDotThreads)
dot_finish_kernel(const float* __restrict__ partial, int count, float* __restrict__ out) {
  __shared__ float smem[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) acc += partial[i];
  const float r = block_sum_256(acc, smem);
  if (threadIdx.x == 0) *out = r;
}

// ---- column statistics -------------------------------------------------------------------------
// Block (128 columns x 4 row lanes) covers kStatRows rows; every thread accumulates a SHIFTED
// first/second moment (shift = first value seen) so the second moment does not cancel, then the 4
// row lanes and afterwards the row chunks are merged with Chan's parallel-variance formula.
constexpr int kStatRows = 256;

__global__ void __launch_bounds__(512)
colstats_partial_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols,
                        float* __restrict__ part /* [chunks][3][cols] */) {
  __shared__ Moments sm[4][128];
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * kStatRows;
  const int64_t r1 = min(rows, r0 + kStatRows);
  Moments m = {0.f, 0.f, 0.f};
  if (c < cols) {
    int64_t r = r0 + threadIdx.y;
    if (r < r1) {
      const float shift = __ldg(x + r * ldx + c);
      float s1 = 0.f, s2 = 0.f, n = 0.f;
      for (; r < r1; r += 4) {
        const float d = __ldg(x + r * ldx + c) - shift;
        s1 += d;
        s2 = fmaf(d, d, s2);
        n += 1.f;
      }
      m.n = n;
      m.sum = fmaf(n, shift, s1);
      m.m2 = fmaxf(s2 - s1 * s1 / n, 0.f);
    }
  }
  sm[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    Moments t = merge(merge(sm[0][threadIdx.x], sm[1][threadIdx.x]), merge(sm[2][threadIdx.x], sm[3][threadIdx.x]));
    float* p = part + (int64_t)blockIdx.y * 3 * cols;
    p[c] = t.n;
    p[cols + c] = t.sum;
    p[2 * cols + c] = t.m2;
  }
}

// Merge the per-chunk partials of 32 columns per block: 32 row lanes each fold a strided subset of the
// chunks in ascending order, then the 32 lane results are folded in lane order (fixed order -> same bits).
__global__ void __launch_bounds__(1024)
colstats_finish_kernel(const float* __restrict__ part, int chunks, int cols, float* __restrict__ sum,
                       float* __restrict__ m2) {
  __shared__ Moments sm[32][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  Moments t = {0.f, 0.f, 0.f};
  if (c < cols) {
    for (int k = threadIdx.y; k < chunks; k += 32) {
      const float* p = part + (int64_t)k * 3 * cols;
      Moments b = {p[c], p[cols + c], p[2 * cols + c]};
      t = merge(t, b);
    }
  }
  sm[threadIdx.y][threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    Moments r = sm[0][threadIdx.x];
#pragma unroll
    for (int l = 1; l < 32; ++l) r = merge(r, sm[l][threadIdx.x]);
    if (sum) sum[c] = r.sum;
    if (m2) m2[c] = r.m2;
  }
}

// Shared with the GEMM epilogue's fused statistics: fold [chunks][3][cols] partials (count, sum, centred m2).
int colstats_finish(const float* part, long long chunks, long long cols, float* sum, float* m2, cudaStream_t stream) {
  if (cols <= 0) return GNNB200_OK;
  colstats_finish_kernel<<<(unsigned)((cols + 31) / 32), dim3(32, 32), 0, stream>>>(part, (int)chunks, (int)cols, sum, m2);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

}  // namespace gnnb200

namespace gnnb200 {
// elementwise_v2.cu: 128-bit loads, 4 rows in flight per lane, 1024-row chunks; same partial layout.  Taken whenever the
// rows are 16-byte aligned (measured on B200 at C5 size: 1.24 -> 0.80 ms for C = 512); the scalar kernel covers the rest
int colstats_partial_v2(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* part, int64_t* chunks_out,
                        cudaStream_t stream);
}  // namespace gnnb200

using namespace gnnb200;

extern "C" int gnnb200_dot_f32(const float* a, const float* b, int64_t n, float* out, void* workspace,
                               size_t* workspace_bytes, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || !workspace_bytes) return GNNB200_EINVAL;
  Workspace ws(workspace);
  float* partial = ws.take<float>(kDotBlocks);
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;
  if (!out || (n > 0 && (!a || !b))) return GNNB200_EINVAL;
  int64_t want = (n / 4 + kDotThreads - 1) / kDotThreads;
  int blocks = (int)(want < 1 ? 1 : (want > kDotBlocks ? kDotBlocks : want));
  dot_partial_kernel<<<blocks, kDotThreads, 0, stream>>>(a, b, n, partial);
  GNNB200_LAUNCH_CHECK();
  dot_finish_kernel<<<1, kDotThreads, 0, stream>>>(partial, blocks, out);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_colstats_f32(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* sum,
                                    float* m2, void* workspace, size_t* workspace_bytes,
                                    gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows < 0 || cols < 0 || !workspace_bytes) return GNNB200_EINVAL;
  if (cols >= (1 << 24)) return GNNB200_ERANGE;
  const int64_t chunks = rows > 0 ? (rows + kStatRows - 1) / kStatRows : 1;
  if (chunks > 65535) return GNNB200_ERANGE;
  Workspace ws(workspace);
  float* part = ws.take<float>((size_t)chunks * 3 * cols);
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;
  if (cols == 0) return GNNB200_OK;
  if (rows > 0 && !x) return GNNB200_EINVAL;
  int rc = GNNB200_EUNSUPPORTED;
  int64_t written = chunks;                    // chunks of `part` the partial kernel fills
  if (rows > 0) {
    rc = colstats_partial_v2(x, ldx, rows, cols, part, &written, stream);
    if (rc != GNNB200_OK && rc != GNNB200_EUNSUPPORTED) return rc;
  }
  if (rc == GNNB200_EUNSUPPORTED) {
    written = chunks;
    dim3 grid((unsigned)((cols + 127) / 128), (unsigned)chunks);
    dim3 block(128, 4);
    colstats_partial_kernel<<<grid, block, 0, stream>>>(x, ldx, rows, (int)cols, part);
    GNNB200_LAUNCH_CHECK();
  }
  colstats_finish_kernel<<<(unsigned)((cols + 31) / 32), dim3(32, 32), 0, stream>>>(part, (int)written, (int)cols, sum, m2);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}
