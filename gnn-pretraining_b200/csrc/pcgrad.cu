// Gradient surgery (the reference's PCGrad variant, src/pretrain/gradient_surgery.py:41-101) on one flat
// [T, P] gradient buffer (SURVEY.md §8f "next" #1).  The reference walks every parameter tensor of every
// ordered task pair on the host with three device->host syncs each (norm == 0, norm == 0, dot < 0): ~2,700
// syncs per s4 step.  Here task i's row is owned by one CTA per (task, parameter-tensor segment):
//   for j in order[0 .. pos(i)-1]:                        (only EARLIER tasks in the shuffled order, :49-58)
//     if segment present in both, |g_i| != 0 and |g_j| != 0:   projections += 1
//       d = <g_i, g_j>;  if d < 0:  conflicts += 1;  g_i -= d / |g_j|^2 * g_j        (g_j = ORIGINAL gradient)
// then out[seg] = mean over the tasks that have the segment, for the segments present in the FIRST shuffled
// task (:60-68).  All reductions are fixed-order block reductions (deterministic); no host sync inside.
#include "common.cuh"

namespace gnnb200 {

constexpr int kPcThreads = 512;

__device__ __forceinline__ float pc_block_sum(float v, float* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                      // protect smem reuse between successive reductions
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = (lane < (kPcThreads >> 5)) ? smem[lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;                             // identical in every thread
}

// grid = (segments, T).  orig/work: [T, P]; present: [T, S] (uint8); order: [T] task ids in shuffled order;
// rank_of[t] = position of task t in `order`.  counters: [T, S, 2] (conflicts, projections).
__global__ void __launch_bounds__(kPcThreads)
pcgrad_project_kernel(const float* __restrict__ orig, float* __restrict__ work, int64_t P,
                      const int64_t* __restrict__ seg_off, const uint8_t* __restrict__ present, int S,
                      const int32_t* __restrict__ order, const int32_t* __restrict__ rank_of,
                      int32_t* __restrict__ counters) {
  __shared__ float smem[32];
  const int s = blockIdx.x, ti = blockIdx.y;
  const int64_t beg = seg_off[s], len = seg_off[s + 1] - beg;
  int conflicts = 0, projections = 0;
  if (present[(int64_t)ti * S + s]) {
    float* gi = work + (int64_t)ti * P + beg;
    const int pos = rank_of[ti];
    for (int q = 0; q < pos; ++q) {
      const int tj = order[q];
      if (!present[(int64_t)tj * S + s]) continue;
      const float* gj = orig + (int64_t)tj * P + beg;
      // "norm == 0" is tested on the elements, not on the fp32 sum of squares: the reference's CPU norm accumulates in
      // double, so a gradient of magnitude 1e-20 (squares underflow in fp32) still counts as non-zero there
      float d = 0.f, nj = 0.f, any_i = 0.f, any_j = 0.f;
      for (int64_t k = threadIdx.x; k < len; k += kPcThreads) {
        const float a = gi[k], b = gj[k];
        d = fmaf(a, b, d);
        nj = fmaf(b, b, nj);
        any_i = (a != 0.f) ? 1.f : any_i;
        any_j = (b != 0.f) ? 1.f : any_j;
      }
      d = pc_block_sum(d, smem);
      nj = pc_block_sum(nj, smem);
      any_i = pc_block_sum(any_i, smem);
      any_j = pc_block_sum(any_j, smem);
      if (any_i == 0.f || any_j == 0.f) continue;    // gradient_surgery.py:89-90
      projections += 1;
      if (d < 0.f) {                                  // :95-101
        conflicts += 1;
        const float c = d / nj;
        for (int64_t k = threadIdx.x; k < len; k += kPcThreads) gi[k] = gi[k] - c * gj[k];
        __syncthreads();
      }
    }
  }
  if (threadIdx.x == 0) {
    counters[((int64_t)ti * S + s) * 2 + 0] = conflicts;
    counters[((int64_t)ti * S + s) * 2 + 1] = projections;
  }
}

// out[k] = mean over tasks (in shuffled order) that have the segment; segments absent from the first shuffled
// task keep has_out[s] = 0 and are left untouched.
__global__ void __launch_bounds__(256)
pcgrad_mean_kernel(const float* __restrict__ work, int64_t P, const int64_t* __restrict__ seg_off,
                   const uint8_t* __restrict__ present, int S, int T, const int32_t* __restrict__ order,
                   float* __restrict__ out, uint8_t* __restrict__ has_out) {
  const int s = blockIdx.y;
  const int64_t beg = seg_off[s], len = seg_off[s + 1] - beg;
  const bool take = present[(int64_t)order[0] * S + s] != 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) has_out[s] = take ? 1 : 0;
  if (!take) return;
  int cnt = 0;
  for (int q = 0; q < T; ++q) cnt += present[(int64_t)order[q] * S + s] ? 1 : 0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < len; k += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int q = 0; q < T; ++q) {
      const int t = order[q];
      if (present[(int64_t)t * S + s]) acc += work[(int64_t)t * P + beg + k];
    }
    out[beg + k] = acc / (float)cnt;
  }
}

}  // namespace gnnb200

using namespace gnnb200;

extern "C" int gnnb200_pcgrad_f32(const float* task_grads, float* work, int64_t num_tasks, int64_t num_params,
                                  const int64_t* seg_offsets, int64_t num_segments, const uint8_t* present,
                                  const int32_t* order, const int32_t* rank_of, float* out, uint8_t* has_out,
                                  int32_t* counters, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_tasks < 0 || num_params < 0 || num_segments < 0) return GNNB200_EINVAL;
  if (num_tasks == 0 || num_segments == 0) return GNNB200_OK;
  if (!task_grads || !work || !seg_offsets || !present || !order || !rank_of || !out || !has_out || !counters) return GNNB200_EINVAL;
  if (num_segments > 2147483647LL || num_tasks > 65535) return GNNB200_ERANGE;
  GNNB200_CHECK_CUDA(cudaMemcpyAsync(work, task_grads, sizeof(float) * num_tasks * num_params, cudaMemcpyDeviceToDevice, stream));
  dim3 g1((unsigned)num_segments, (unsigned)num_tasks);
  pcgrad_project_kernel<<<g1, kPcThreads, 0, stream>>>(task_grads, work, num_params, seg_offsets, present, (int)num_segments,
                                                       order, rank_of, counters);
  GNNB200_LAUNCH_CHECK();
  if (num_segments > 65535) return GNNB200_ERANGE;
  dim3 g2(8, (unsigned)num_segments);
  pcgrad_mean_kernel<<<g2, 256, 0, stream>>>(work, num_params, seg_offsets, present, (int)num_segments, (int)num_tasks, order,
                                             out, has_out);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}
