// Global graph pooling over contiguous node segments (K6/K7 of SURVEY §2.4): replaces PyG
// scatter(reduce='mean'|'max'|'sum') behind global_mean_pool / global_max_pool
// (reference src/models/finetune_model.py:75, src/pretrain/tasks.py:241-246,299,331).
// PyG batches keep each graph's nodes contiguous, so a segment is a row range [ptr[g], ptr[g+1]).
// Lanes run across features (float4 when possible), rows are reduced sequentially in row order
// (bit-identical to the CPU scatter order for segments up to kSplitRows rows); longer segments are
// cut into kSplitRows-row chunks whose partials are then summed in chunk order (deterministic).
#include <float.h>
#include "common.cuh"

namespace gnnb200 {

constexpr int kSplitRows = 2048;

__device__ __forceinline__ float pool_combine(int mode, float a, float b) {
  return mode == GNNB200_POOL_MAX ? fmaxf(a, b) : __fadd_rn(a, b);
}

// grid: (ceil(feat/128), num_segments, splits).  Each thread owns one feature column f and walks
// rows [r0, r1).  Coalesced: consecutive threads read consecutive floats of a row.
__global__ void __launch_bounds__(128)
pool_fwd_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ ptr, int feat,
                int mode, int splits, float* __restrict__ out, int64_t ldo, float* __restrict__ partial) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = blockIdx.y;
  const int sp = blockIdx.z;
  if (f >= feat) return;
  const int beg = ptr[seg], end = ptr[seg + 1];
  int r0 = beg, r1 = end;
  if (splits > 1) {
    r0 = beg + sp * kSplitRows;
    r1 = min(end, r0 + kSplitRows);
  }
  float acc = (mode == GNNB200_POOL_MAX) ? -FLT_MAX : 0.f;
  int r = r0;
  for (; r + 4 <= r1; r += 4) {  // 4 independent loads in flight, combined in row order
    const float v0 = __ldg(x + (int64_t)(r + 0) * ldx + f);
    const float v1 = __ldg(x + (int64_t)(r + 1) * ldx + f);
    const float v2 = __ldg(x + (int64_t)(r + 2) * ldx + f);
    const float v3 = __ldg(x + (int64_t)(r + 3) * ldx + f);
    acc = pool_combine(mode, pool_combine(mode, pool_combine(mode, pool_combine(mode, acc, v0), v1), v2), v3);
  }
  for (; r < r1; ++r) acc = pool_combine(mode, acc, __ldg(x + (int64_t)r * ldx + f));
  if (splits > 1) {
    partial[((int64_t)seg * splits + sp) * feat + f] = acc;  // empty chunks hold the identity
    return;
  }
  if (end <= beg) acc = 0.f;
  if (mode == GNNB200_POOL_MEAN) acc = acc / (float)max(end - beg, 1);
  out[(int64_t)seg * ldo + f] = acc;
}

__global__ void __launch_bounds__(128)
pool_fwd_finish_kernel(const float* __restrict__ partial, const int32_t* __restrict__ ptr, int feat, int mode,
                       int splits, float* __restrict__ out, int64_t ldo) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = blockIdx.y;
  if (f >= feat) return;
  const int beg = ptr[seg], end = ptr[seg + 1];
  const int used = (end - beg + kSplitRows - 1) / kSplitRows;
  float acc = (mode == GNNB200_POOL_MAX) ? -FLT_MAX : 0.f;
  for (int s = 0; s < used; ++s) acc = pool_combine(mode, acc, partial[((int64_t)seg * splits + s) * feat + f]);
  if (end <= beg) acc = 0.f;
  if (mode == GNNB200_POOL_MEAN) acc = acc / (float)max(end - beg, 1);
  out[(int64_t)seg * ldo + f] = acc;
}

// MEAN / SUM backward: dx[r,f] = g[seg,f] (/ cnt).  grid: (ceil(feat/128), num_segments, row chunks)
__global__ void __launch_bounds__(128)
pool_bwd_linear_kernel(const float* __restrict__ g, int64_t ldg, const int32_t* __restrict__ ptr, int feat,
                       int mode, float* __restrict__ gx, int64_t ldgx) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = blockIdx.y;
  if (f >= feat) return;
  const int beg = ptr[seg], end = ptr[seg + 1];
  float v = g[(int64_t)seg * ldg + f];
  if (mode == GNNB200_POOL_MEAN) v = v / (float)max(end - beg, 1);
  for (int r = beg + blockIdx.z; r < end; r += gridDim.z) gx[(int64_t)r * ldgx + f] = v;
}

// MAX backward, torch's native scatter_reduce_('amax', include_self=False) rule on a zero-initialised
// output (App. A.3): the gradient is split evenly over the rows that hold the maximum, and the
// zero-initialised slot counts as one more tie when the maximum is exactly 0.0.
__global__ void __launch_bounds__(128)
pool_bwd_max_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x, int64_t ldx,
                    const float* __restrict__ out, int64_t ldo, const int32_t* __restrict__ ptr, int feat,
                    float* __restrict__ gx, int64_t ldgx) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = blockIdx.y;
  if (f >= feat) return;
  const int beg = ptr[seg], end = ptr[seg + 1];
  const float m = out[(int64_t)seg * ldo + f];
  int ties = (m == 0.0f) ? 1 : 0;
  for (int r = beg; r < end; ++r) ties += (__ldg(x + (int64_t)r * ldx + f) == m) ? 1 : 0;
  const float share = g[(int64_t)seg * ldg + f] / (float)max(ties, 1);
  for (int r = beg; r < end; ++r)
    gx[(int64_t)r * ldgx + f] = (__ldg(x + (int64_t)r * ldx + f) == m) ? share : 0.f;
}

}  // namespace gnnb200

using namespace gnnb200;

static int pool_splits(int64_t num_rows, int64_t num_segments) {
  // The host cannot see segment lengths without a sync, so the split count is sized for the worst
  // case (one segment holding every row) only when the average segment is long.
  if (num_segments <= 0) return 1;
  if (num_rows / num_segments <= kSplitRows / 2) return 1;
  int64_t s = (num_rows + kSplitRows - 1) / kSplitRows;
  return (int)(s < 1 ? 1 : s);
}

extern "C" int gnnb200_segment_pool_fwd_f32(const float* x, int64_t ldx, const int32_t* ptr, int64_t num_rows,
                                            int64_t num_segments, int64_t feat, int mode, float* out, int64_t ldo,
                                            void* workspace, size_t* workspace_bytes, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_rows < 0 || num_segments < 0 || feat < 0 || mode < 0 || mode > 2 || !workspace_bytes) return GNNB200_EINVAL;
  if (num_segments > 65535 * 1024LL || feat >= (1 << 24)) return GNNB200_ERANGE;
  const int splits = pool_splits(num_rows, num_segments);
  Workspace ws(workspace);
  float* partial = splits > 1 ? ws.take<float>((size_t)num_segments * splits * feat) : nullptr;
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;
  if (num_segments == 0 || feat == 0) return GNNB200_OK;
  if (!x && num_rows > 0) return GNNB200_EINVAL;
  if (!ptr || !out) return GNNB200_EINVAL;
  if (num_segments > 65535 || splits > 65535) return GNNB200_ERANGE;
  dim3 grid((unsigned)((feat + 127) / 128), (unsigned)num_segments, (unsigned)splits);
  pool_fwd_kernel<<<grid, 128, 0, stream>>>(x, ldx, ptr, (int)feat, mode, splits, out, ldo, partial);
  GNNB200_LAUNCH_CHECK();
  if (splits > 1) {
    dim3 g2((unsigned)((feat + 127) / 128), (unsigned)num_segments);
    pool_fwd_finish_kernel<<<g2, 128, 0, stream>>>(partial, ptr, (int)feat, mode, splits, out, ldo);
    GNNB200_LAUNCH_CHECK();
  }
  return GNNB200_OK;
}

extern "C" int gnnb200_segment_pool_bwd_f32(const float* grad_out, int64_t ldg, const float* x, int64_t ldx,
                                            const float* out, int64_t ldo, const int32_t* ptr, int64_t num_rows,
                                            int64_t num_segments, int64_t feat, int mode, float* grad_x,
                                            int64_t ldgx, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_rows < 0 || num_segments < 0 || feat < 0 || mode < 0 || mode > 2) return GNNB200_EINVAL;
  if (num_segments == 0 || feat == 0 || num_rows == 0) return GNNB200_OK;
  if (!grad_out || !ptr || !grad_x) return GNNB200_EINVAL;
  if (num_segments > 65535) return GNNB200_ERANGE;
  if (mode == GNNB200_POOL_MAX) {
    if (!x || !out) return GNNB200_EINVAL;
    dim3 grid((unsigned)((feat + 127) / 128), (unsigned)num_segments);
    pool_bwd_max_kernel<<<grid, 128, 0, stream>>>(grad_out, ldg, x, ldx, out, ldo, ptr, (int)feat, grad_x, ldgx);
  } else {
    int64_t z = num_rows / num_segments / 64;
    if (z < 1) z = 1;
    if (z > 1024) z = 1024;
    dim3 grid((unsigned)((feat + 127) / 128), (unsigned)num_segments, (unsigned)z);
    pool_bwd_linear_kernel<<<grid, 128, 0, stream>>>(grad_out, ldg, ptr, (int)feat, mode, grad_x, ldgx);
  }
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}
