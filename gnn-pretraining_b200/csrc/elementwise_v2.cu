// Column-stationary BatchNorm(+ReLU+dropout) apply kernels and the vectorised column-statistics kernel: the kernels
// gnnb200_bn_act_fwd/bwd_f32 and gnnb200_colstats_f32 launch (bn.cu / reduce.cu keep the first versions for the shapes these
// do not cover).
//
// Why.  Per-launch times of the C5 step at quarter scale (profiles/r01e_launches_scale0.25_tf32.csv; bytes / time against
// the measured 6.55 TB/s copy peak):
//   bn_act_fwd<no dropout, C=512> 0.99   bn_act_fwd<dropout, C=256> 0.62   colstats_partial 0.79-0.84 (+ 0.09 ms finish)
//   bn_act_bwd_reduce 0.93 / 0.77        bn_act_bwd_apply<C=512> 0.69      bn_act_bwd_apply<dropout, C=256> 0.80
// The two apply kernels walk the matrix as one flat grid-stride loop: every 4 elements pay a 64-bit division to find
// their row and re-load the per-column parameter vectors (4 forward, 6 backward) through L1 — more load instructions
// for parameters than for the streams that come from HBM; with dropout the Philox rounds on top make the forward kernel
// issue-bound.  The statistics kernel reads 4 bytes per lane with one row in flight.  The reduce kernel, which is
// column-stationary, has neither problem — so every kernel here takes its shape: a block owns 128 columns (32 lanes x
// float4) of a 256-row chunk, its 8 warps interleave the rows, the per-column parameters live in registers for the
// whole chunk, and several rows' 128-bit loads are issued before the first one is consumed.
//
// Same arithmetic expressions and the same Philox counter (row * cols/4 + column quad) as the first versions (the dropout
// mask does not depend on which kernel ran); the statistics differ from the first version only in the (still fixed) order
// in which a chunk's partial moments are merged.  Measured on B200, C5 size (2.45 M rows; fraction of the 6.55 TB/s copy
// peak, first version -> this one): forward C=512 0.95 -> 0.99, forward + dropout 0.66 -> 0.83, backward reduce + apply
// C=512 0.75 -> 0.97, column statistics 0.62 -> 0.95 (gpurun_out of round 2, profiles/r02_elementwise.md).
#include "common.cuh"
#include "ew_common.cuh"

namespace gnnb200 {

constexpr int kV2Rows = 256;     // rows per block = 8 row lanes x 32 rows; equals kBnRows / kStatRows (same partial layout)
constexpr int kV2Lanes = 8;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <bool DROP>
__device__ __forceinline__ float4 bn_fwd_one(const float4 v, const float4 mu, const float4 is, const float4 ga,
                                             const float4 be, int relu, uint64_t seed, uint64_t q, uint32_t thresh,
                                             float scale) {
  float o[4] = {(v.x - mu.x) * is.x * ga.x + be.x, (v.y - mu.y) * is.y * ga.y + be.y,
                (v.z - mu.z) * is.z * ga.z + be.z, (v.w - mu.w) * is.w * ga.w + be.w};
  if (relu) {
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = fmaxf(o[i], 0.f);
  }
  if (DROP) {
    bool k[4];
    keep4(seed, q, thresh, k);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = k[i] ? o[i] * scale : 0.f;
  }
  return make_float4(o[0], o[1], o[2], o[3]);
}

template <bool DROP>
__global__ void __launch_bounds__(256)
bn_act_fwd_v2_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ mean,
                     const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                     int relu, uint64_t seed, uint32_t thresh, float scale, int64_t rows, int cols,
                     float* __restrict__ y, int64_t ldy) {
  const int c4 = cols >> 2;
  const int cq = blockIdx.x * 32 + threadIdx.x;
  if (cq >= c4) return;
  const int c = cq << 2;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
  const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
  const int64_t r_end = min(rows, ((int64_t)blockIdx.y + 1) * kV2Rows);
  int64_t r = (int64_t)blockIdx.y * kV2Rows + threadIdx.y;
  for (; r + 3 * kV2Lanes < r_end; r += 4 * kV2Lanes) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld4(x + (r + u * kV2Lanes) * ldx + c);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t ru = r + u * kV2Lanes;
      *reinterpret_cast<float4*>(y + ru * ldy + c) =
          bn_fwd_one<DROP>(v[u], mu, is, ga, be, relu, seed, (uint64_t)(ru * c4 + cq), thresh, scale);
    }
  }
  for (; r < r_end; r += kV2Lanes)
    *reinterpret_cast<float4*>(y + r * ldy + c) =
        bn_fwd_one<DROP>(ld4(x + r * ldx + c), mu, is, ga, be, relu, seed, (uint64_t)(r * c4 + cq), thresh, scale);
}

template <bool DROP>
__device__ __forceinline__ float4 bn_bwd_one(const float4 gv, const float4 xv, const float4 mu, const float4 is,
                                             const float4 ga, const float4 be, const float4 dg, const float4 db,
                                             int relu, int training, uint64_t seed, uint64_t q, uint32_t thresh,
                                             float scale, float inv_n) {
  float g1[4], xh[4];
  bn_g1<DROP>(gv, xv, mu, is, ga, be, relu, seed, q, thresh, scale, g1, xh);
  const float isv[4] = {is.x, is.y, is.z, is.w};
  const float gav[4] = {ga.x, ga.y, ga.z, ga.w};
  float o[4];
  if (training) {
    const float dgv[4] = {dg.x, dg.y, dg.z, dg.w};
    const float dbv[4] = {db.x, db.y, db.z, db.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = gav[i] * isv[i] * (g1[i] - dbv[i] * inv_n - xh[i] * dgv[i] * inv_n);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = gav[i] * isv[i] * g1[i];
  }
  return make_float4(o[0], o[1], o[2], o[3]);
}

template <bool DROP>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_v2_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x, int64_t ldx,
                           const float* __restrict__ mean, const float* __restrict__ invstd,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           const float* __restrict__ dgamma, const float* __restrict__ dbeta, int relu, int training,
                           uint64_t seed, uint32_t thresh, float scale, int64_t rows, int64_t rows_total, int cols,
                           float* __restrict__ dx, int64_t lddx) {
  const int c4 = cols >> 2;
  const int cq = blockIdx.x * 32 + threadIdx.x;
  if (cq >= c4) return;
  const int c = cq << 2;
  const float inv_n = 1.f / (float)rows_total;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + c));
  const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + c));
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
  float4 dg = make_float4(0.f, 0.f, 0.f, 0.f), db = dg;
  if (training) {
    dg = __ldg(reinterpret_cast<const float4*>(dgamma + c));
    db = __ldg(reinterpret_cast<const float4*>(dbeta + c));
  }
  const int64_t r_end = min(rows, ((int64_t)blockIdx.y + 1) * kV2Rows);
  int64_t r = (int64_t)blockIdx.y * kV2Rows + threadIdx.y;
  for (; r + kV2Lanes < r_end; r += 2 * kV2Lanes) {           // two rows (four 128-bit loads) in flight per lane
    const int64_t ra = r, rb = r + kV2Lanes;
    const float4 ga_ = ld4(g + ra * ldg + c), xa = ld4(x + ra * ldx + c);
    const float4 gb_ = ld4(g + rb * ldg + c), xb = ld4(x + rb * ldx + c);
    *reinterpret_cast<float4*>(dx + ra * lddx + c) = bn_bwd_one<DROP>(ga_, xa, mu, is, ga, be, dg, db, relu, training, seed,
                                                                       (uint64_t)(ra * c4 + cq), thresh, scale, inv_n);
    *reinterpret_cast<float4*>(dx + rb * lddx + c) = bn_bwd_one<DROP>(gb_, xb, mu, is, ga, be, dg, db, relu, training, seed,
                                                                       (uint64_t)(rb * c4 + cq), thresh, scale, inv_n);
  }
  for (; r < r_end; r += kV2Lanes)
    *reinterpret_cast<float4*>(dx + r * lddx + c) =
        bn_bwd_one<DROP>(ld4(g + r * ldg + c), ld4(x + r * ldx + c), mu, is, ga, be, dg, db, relu, training, seed,
                         (uint64_t)(r * c4 + cq), thresh, scale, inv_n);
}

// Column statistics: every lane owns 4 columns and folds its rows into SHIFTED first/second moments (shift = the
// first value it sees, so the second moment does not cancel); the 8 row lanes of a block are merged in lane order
// with Chan's formula.  Partial layout [chunks][3][cols] = (count, sum, centred m2): the first version's, so
// colstats_finish_kernel folds the chunks unchanged.
__global__ void __launch_bounds__(256)
colstats_partial_v2_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int cols, int rows_per_block,
                           float* __restrict__ part) {
  __shared__ float sm_n[kV2Lanes][32];
  __shared__ float sm_sum[kV2Lanes][128];
  __shared__ float sm_m2[kV2Lanes][128];
  const int c4 = cols >> 2;
  const int cq = blockIdx.x * 32 + threadIdx.x;
  const int c = cq << 2;
  const int64_t r_end = min(rows, ((int64_t)blockIdx.y + 1) * rows_per_block);
  int64_t r = (int64_t)blockIdx.y * rows_per_block + threadIdx.y;
  float n = 0.f;
  float sum[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
  if (cq < c4 && r < r_end) {
    const float4 first = ld4(x + r * ldx + c);
    const float shift[4] = {first.x, first.y, first.z, first.w};
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    for (; r + 3 * kV2Lanes < r_end; r += 4 * kV2Lanes) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld4(x + (r + u * kV2Lanes) * ldx + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float d[4] = {v[u].x - shift[0], v[u].y - shift[1], v[u].z - shift[2], v[u].w - shift[3]};
#pragma unroll
        for (int i = 0; i < 4; ++i) { s1[i] += d[i]; s2[i] = fmaf(d[i], d[i], s2[i]); }
      }
      n += 4.f;
    }
    for (; r < r_end; r += kV2Lanes) {
      const float4 v = ld4(x + r * ldx + c);
      const float d[4] = {v.x - shift[0], v.y - shift[1], v.z - shift[2], v.w - shift[3]};
#pragma unroll
      for (int i = 0; i < 4; ++i) { s1[i] += d[i]; s2[i] = fmaf(d[i], d[i], s2[i]); }
      n += 1.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sum[i] = fmaf(n, shift[i], s1[i]);
      m2[i] = fmaxf(s2[i] - s1[i] * s1[i] / n, 0.f);
    }
  }
  sm_n[threadIdx.y][threadIdx.x] = n;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sm_sum[threadIdx.y][threadIdx.x * 4 + i] = sum[i];
    sm_m2[threadIdx.y][threadIdx.x * 4 + i] = m2[i];
  }
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;
  const int col = blockIdx.x * 128 + t;
  if (t < 128 && col < cols) {
    Moments acc = {sm_n[0][t >> 2], sm_sum[0][t], sm_m2[0][t]};
#pragma unroll
    for (int l = 1; l < kV2Lanes; ++l) acc = merge(acc, Moments{sm_n[l][t >> 2], sm_sum[l][t], sm_m2[l][t]});
    float* p = part + (int64_t)blockIdx.y * 3 * cols;
    p[col] = acc.n;
    p[cols + col] = acc.sum;
    p[2 * cols + col] = acc.m2;
  }
}

static inline bool v2_layout_ok(int64_t cols, int64_t rows) {
  return cols % 4 == 0 && cols > 0 && rows > 0 && (rows + kV2Rows - 1) / kV2Rows <= 65535;
}

// Host launchers.  Return GNNB200_EUNSUPPORTED when the shape is outside what these variants cover (the caller then
// takes the first version); the caller has already validated pointers, alignment and leading dimensions.
int bn_act_fwd_v2(const float* x, int64_t ldx, const float* mean, const float* invstd, const float* gamma,
                  const float* beta, int relu, bool drop, uint64_t seed, uint32_t thresh, float scale, int64_t rows,
                  int64_t cols, float* y, int64_t ldy, cudaStream_t stream) {
  if (!v2_layout_ok(cols, rows)) return GNNB200_EUNSUPPORTED;
  const dim3 grid((unsigned)((cols / 4 + 31) / 32), (unsigned)((rows + kV2Rows - 1) / kV2Rows)), block(32, kV2Lanes);
  if (drop)
    bn_act_fwd_v2_kernel<true><<<grid, block, 0, stream>>>(x, ldx, mean, invstd, gamma, beta, relu, seed, thresh, scale, rows, (int)cols, y, ldy);
  else
    bn_act_fwd_v2_kernel<false><<<grid, block, 0, stream>>>(x, ldx, mean, invstd, gamma, beta, relu, seed, thresh, scale, rows, (int)cols, y, ldy);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

int bn_act_bwd_apply_v2(const float* g, int64_t ldg, const float* x, int64_t ldx, const float* mean, const float* invstd,
                        const float* gamma, const float* beta, const float* dgamma, const float* dbeta, int relu,
                        int training, bool drop, uint64_t seed, uint32_t thresh, float scale, int64_t rows,
                        int64_t rows_total, int64_t cols, float* dx, int64_t lddx, cudaStream_t stream) {
  if (!v2_layout_ok(cols, rows)) return GNNB200_EUNSUPPORTED;
  const dim3 grid((unsigned)((cols / 4 + 31) / 32), (unsigned)((rows + kV2Rows - 1) / kV2Rows)), block(32, kV2Lanes);
  if (drop)
    bn_act_bwd_apply_v2_kernel<true><<<grid, block, 0, stream>>>(g, ldg, x, ldx, mean, invstd, gamma, beta, dgamma, dbeta, relu, training, seed, thresh, scale, rows, rows_total, (int)cols, dx, lddx);
  else
    bn_act_bwd_apply_v2_kernel<false><<<grid, block, 0, stream>>>(g, ldg, x, ldx, mean, invstd, gamma, beta, dgamma, dbeta, relu, training, seed, thresh, scale, rows, rows_total, (int)cols, dx, lddx);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

// 1024 rows per block: four times fewer partial moments than the first version's 256-row chunks, so the (latency-bound,
// 0.09 ms at quarter scale) finish kernel folds four times fewer.  `part` is the caller's [ceil(rows/256)][3][cols]
// buffer, of which the first *chunks_out chunks are written.
constexpr int kColstatsV2Rows = 1024;

int colstats_partial_v2(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* part, int64_t* chunks_out,
                        cudaStream_t stream) {
  if (!v2_layout_ok(cols, rows) || ldx % 4 != 0 || ((uintptr_t)x & 15) != 0) return GNNB200_EUNSUPPORTED;
  const int64_t chunks = (rows + kColstatsV2Rows - 1) / kColstatsV2Rows;
  const dim3 grid((unsigned)((cols / 4 + 31) / 32), (unsigned)chunks), block(32, kV2Lanes);
  colstats_partial_v2_kernel<<<grid, block, 0, stream>>>(x, ldx, rows, (int)cols, kColstatsV2Rows, part);
  GNNB200_LAUNCH_CHECK();
  *chunks_out = chunks;
  return GNNB200_OK;
}

}  // namespace gnnb200
