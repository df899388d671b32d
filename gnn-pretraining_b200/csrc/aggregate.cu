// Per-destination neighbour aggregation over a sorted CSR (K1/K2/K1^T of SURVEY §2.4).
// Replaces PyG GINConv.propagate's index_select + scatter_add_ and the (1+eps)*x self term
// (reference src/models/gnn.py:29-37,41).  HBM-bound gather: G lanes own one destination row,
// each lane keeps V float4 accumulators (row width F <= G*V*4), neighbour rows are streamed with
// 128-bit no-allocate loads U rows at a time at high occupancy (register-capped), and the adds are
// applied strictly in edge order from 0.0f -> bit-identical to the CPU scatter_add_ order, no
// float atomics.  The backward pass is the same kernel on the by-src CSR.
#include "common.cuh"

namespace gnnb200 {

// SKIP (SUM mode, opt-in through GNNB200_AGG_SKIP_LONG): rows with more than GNNB200_AGG_LONG_ROW neighbours are left to
// aggregate_long_rows_kernel below, so one hub node cannot hold the whole launch hostage to a single warp's serial walk.
template <int G, int V, int MODE, int U = 4, int MINB = 1, bool SKIP = false>
__global__ void __launch_bounds__(256, MINB)
aggregate_vec_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ col, int64_t num_rows, int feat,
                     const float* __restrict__ self_x, int64_t lds, const float* __restrict__ eps_ptr,
                     const float* __restrict__ dinv, float* __restrict__ out, int64_t ldo, int accumulate) {
  // U = neighbour rows in flight per group
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);                 // lane inside the row group
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (group >= num_rows) return;
  const int64_t row = group;
  const int beg = rowptr[row], end = rowptr[row + 1];
  if (SKIP && end - beg > GNNB200_AGG_LONG_ROW) return;
  const int nvec = feat >> 2;                    // float4 per row

  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    // accumulate: continue a sum started by an earlier pass over another column block of the same rows
    // (chunked halo exchange, gnnb200/partition.py); the order of additions stays fixed
    acc[v] = (accumulate && gl + v * G < nvec) ? *reinterpret_cast<const float4*>(out + row * ldo + 4 * (gl + v * G))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float di = 0.f;
  if (MODE == GNNB200_AGG_GCN) di = dinv[row];

  for (int base = beg; base < end; base += G) {
    // one coalesced load of up to G column ids, then broadcast inside the group
    int my_col = (base + gl < end) ? col[base + gl] : 0;
    float my_w = 0.f;
    if (MODE == GNNB200_AGG_GCN) my_w = (base + gl < end) ? __fmul_rn(dinv[my_col], di) : 0.f;
    const int cnt = min(G, end - base);
    for (int j = 0; j < cnt; j += U) {
      float4 nb[U][V];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = min(j + u, cnt - 1);      // clamp: shuffles stay convergent
        const int c = __shfl_sync(gmask, my_col, jj, G);
        if (MODE == GNNB200_AGG_GCN) w[u] = __shfl_sync(gmask, my_w, jj, G);
        const float4* src = reinterpret_cast<const float4*>(x + (int64_t)c * ldx);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int k = gl + v * G;
          if (j + u < cnt && k < nvec) nb[u][v] = ldg_stream(src + k);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j + u < cnt) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            if (gl + v * G < nvec)
              acc[v] = (MODE == GNNB200_AGG_GCN) ? f4_add(acc[v], f4_scale(w[u], nb[u][v])) : f4_add(acc[v], nb[u][v]);
          }
        }
      }
    }
  }

  float scale = 0.f;
  if (self_x != nullptr) scale = (MODE == GNNB200_AGG_GCN) ? __fmul_rn(di, di) : __fadd_rn(1.0f, eps_ptr ? *eps_ptr : 0.f);
  float inv_cnt = 1.f;
  if (MODE == GNNB200_AGG_MEAN) inv_cnt = (float)max(end - beg, 1);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int k = gl + v * G;
    if (k < nvec) {
      float4 r = acc[v];
      if (MODE == GNNB200_AGG_MEAN) r = make_float4(r.x / inv_cnt, r.y / inv_cnt, r.z / inv_cnt, r.w / inv_cnt);
      if (self_x != nullptr) {
        const float4 s = *reinterpret_cast<const float4*>(self_x + row * lds + 4 * k);
        r = f4_add(r, f4_scale(scale, s));
      }
      *reinterpret_cast<float4*>(out + row * ldo + 4 * k) = r;
    }
  }
}

// Any width / alignment: one warp per row, lanes stride the features one float at a time.
template <int MODE>
__global__ void __launch_bounds__(256)
aggregate_scalar_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                        const int32_t* __restrict__ col, int64_t num_rows, int feat,
                        const float* __restrict__ self_x, int64_t lds, const float* __restrict__ eps_ptr,
                        const float* __restrict__ dinv, float* __restrict__ out, int64_t ldo, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= num_rows) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  const float di = (MODE == GNNB200_AGG_GCN) ? dinv[row] : 0.f;
  float scale = 0.f;
  if (self_x != nullptr) scale = (MODE == GNNB200_AGG_GCN) ? __fmul_rn(di, di) : __fadd_rn(1.0f, eps_ptr ? *eps_ptr : 0.f);
  const float cnt = (float)max(end - beg, 1);
  for (int f = lane; f < feat; f += 32) {
    float acc = accumulate ? out[row * ldo + f] : 0.f;
    for (int e = beg; e < end; ++e) {
      const int c = col[e];
      const float v = __ldg(x + (int64_t)c * ldx + f);
      acc = (MODE == GNNB200_AGG_GCN) ? __fadd_rn(acc, __fmul_rn(__fmul_rn(dinv[c], di), v)) : __fadd_rn(acc, v);
    }
    if (MODE == GNNB200_AGG_MEAN) acc = acc / cnt;
    if (self_x != nullptr) acc = __fadd_rn(acc, __fmul_rn(scale, self_x[row * lds + f]));
    out[row * ldo + f] = acc;
  }
}

// One CTA per long row (a hub of a power-law graph): its 8 warps walk 8 contiguous slices of the neighbour list, each in
// edge order with the same streaming loads as above, and the 8 partial rows are added in warp order through shared
// memory: deterministic (same inputs -> same bits), but a different association than the single sequential sum, so these
// rows agree with the CPU order to fp32 rounding (1e-5 class) instead of bit for bit.
template <int V>
__global__ void __launch_bounds__(256)
aggregate_long_rows_kernel(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ rowptr,
                           const int32_t* __restrict__ col, const int64_t* __restrict__ rows, int feat,
                           const float* __restrict__ self_x, int64_t lds, const float* __restrict__ eps_ptr,
                           float* __restrict__ out, int64_t ldo, int accumulate) {
  extern __shared__ float4 part[];               // [8 warps][nvec]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = rows[blockIdx.x];
  const int beg = rowptr[row], end = rowptr[row + 1];
  const int nvec = feat >> 2;
  const int per = (((end - beg + 7) >> 3) + 31) & ~31;          // neighbours per warp, a multiple of 32
  const int w_beg = min(end, beg + warp * per), w_end = min(end, w_beg + per);
  constexpr int U = 2;

  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = w_beg; base < w_end; base += 32) {
    const int my_col = (base + lane < w_end) ? col[base + lane] : 0;
    const int cnt = min(32, w_end - base);
    for (int j = 0; j < cnt; j += U) {
      float4 nb[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = min(j + u, cnt - 1);
        const int c = __shfl_sync(0xffffffffu, my_col, jj);
        const float4* src = reinterpret_cast<const float4*>(x + (int64_t)c * ldx);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int k = lane + v * 32;
          if (j + u < cnt && k < nvec) nb[u][v] = ldg_stream(src + k);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j + u < cnt) {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (lane + v * 32 < nvec) acc[v] = f4_add(acc[v], nb[u][v]);
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int k = lane + v * 32;
    if (k < nvec) part[warp * nvec + k] = acc[v];
  }
  __syncthreads();
  const float scale = (self_x != nullptr) ? __fadd_rn(1.0f, eps_ptr ? *eps_ptr : 0.f) : 0.f;
  for (int k = threadIdx.x; k < nvec; k += 256) {
    float4 r = accumulate ? *reinterpret_cast<const float4*>(out + row * ldo + 4 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < 8; ++w) r = f4_add(r, part[w * nvec + k]);
    if (self_x != nullptr) r = f4_add(r, f4_scale(scale, *reinterpret_cast<const float4*>(self_x + row * lds + 4 * k)));
    *reinterpret_cast<float4*>(out + row * ldo + 4 * k) = r;
  }
}

template <int G, int V>
static int launch_vec(int mode, int accumulate, const float* x, int64_t ldx, const int32_t* rowptr, const int32_t* col,
                      int64_t num_rows, int feat, const float* self_x, int64_t lds, const float* eps,
                      const float* dinv, float* out, int64_t ldo, cudaStream_t stream, bool skip_long = false) {
  const int block = 256;
  const int64_t threads = num_rows * G;
  const unsigned grid = (unsigned)((threads + block - 1) / block);
  // Occupancy beats per-warp memory-level parallelism for this gather (measured on B200 at C5 scale, F = 256:
  // U=4 / 91 registers -> 16.2 ms per pass, U=2 capped at 40 registers (6 CTAs/SM) -> 10.5 ms = the measured
  // HBM copy peak), so the register cap grows only with the accumulator count V.
  constexpr int U = (V >= 8) ? 1 : 2;
  constexpr int MINB = (V <= 2) ? 6 : ((V == 4) ? 3 : 2);
  switch (mode) {
    case GNNB200_AGG_SUM:
      if (skip_long)
        aggregate_vec_kernel<G, V, GNNB200_AGG_SUM, U, MINB, true><<<grid, block, 0, stream>>>(x, ldx, rowptr, col, num_rows, feat, self_x, lds, eps, dinv, out, ldo, accumulate);
      else
        aggregate_vec_kernel<G, V, GNNB200_AGG_SUM, U, MINB><<<grid, block, 0, stream>>>(x, ldx, rowptr, col, num_rows, feat, self_x, lds, eps, dinv, out, ldo, accumulate);
      break;
    case GNNB200_AGG_MEAN:
      aggregate_vec_kernel<G, V, GNNB200_AGG_MEAN, U, MINB><<<grid, block, 0, stream>>>(x, ldx, rowptr, col, num_rows, feat, self_x, lds, eps, dinv, out, ldo, accumulate);
      break;
    default:
      aggregate_vec_kernel<G, V, GNNB200_AGG_GCN, U, MINB><<<grid, block, 0, stream>>>(x, ldx, rowptr, col, num_rows, feat, self_x, lds, eps, dinv, out, ldo, accumulate);
      break;
  }
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

}  // namespace gnnb200

using namespace gnnb200;

extern "C" int gnnb200_aggregate_f32(const float* x, int64_t ldx, const int32_t* rowptr, const int32_t* col,
                                     int64_t num_rows, int64_t feat, int mode, const float* self_x, int64_t lds,
                                     const float* eps, const float* dinv, float* out, int64_t ldo,
                                     gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int accumulate = (mode & GNNB200_AGG_ACCUMULATE) ? 1 : 0;
  const bool skip_long = (mode & GNNB200_AGG_SKIP_LONG) != 0;
  mode &= ~(GNNB200_AGG_ACCUMULATE | GNNB200_AGG_SKIP_LONG);
  if (num_rows < 0 || feat < 0 || mode < 0 || mode > 2) return GNNB200_EINVAL;
  if ((accumulate || skip_long) && mode != GNNB200_AGG_SUM) return GNNB200_EINVAL;
  if (num_rows == 0 || feat == 0) return GNNB200_OK;
  if (!x || !rowptr || !out) return GNNB200_EINVAL;
  if (mode == GNNB200_AGG_GCN && (!dinv || !self_x)) return GNNB200_EINVAL;
  if (num_rows >= (int64_t)INT32_MAX || feat >= (1 << 20)) return GNNB200_ERANGE;
  const bool vec_ok = (feat % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && (!self_x || lds % 4 == 0) &&
                      ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)self_x % 16 == 0) &&
                      feat <= 1024;
  const int f = (int)feat;
#define GNNB200_AGG_ARGS mode, accumulate, x, ldx, rowptr, col, num_rows, f, self_x, lds, eps, dinv, out, ldo, stream, skip_long
  if (skip_long && !vec_ok) return GNNB200_EUNSUPPORTED;   // the long-row kernel exists for the 128-bit layouts only
  if (vec_ok) {
    const int nvec = f / 4;
    if (nvec <= 4) return launch_vec<4, 1>(GNNB200_AGG_ARGS);
    if (nvec <= 8) return launch_vec<8, 1>(GNNB200_AGG_ARGS);
    if (nvec <= 16) return launch_vec<16, 1>(GNNB200_AGG_ARGS);
    if (nvec <= 32) return launch_vec<32, 1>(GNNB200_AGG_ARGS);
    if (nvec <= 64) return launch_vec<32, 2>(GNNB200_AGG_ARGS);
    if (nvec <= 128) return launch_vec<32, 4>(GNNB200_AGG_ARGS);
    return launch_vec<32, 8>(GNNB200_AGG_ARGS);
  }
#undef GNNB200_AGG_ARGS
  const int block = 256;
  const unsigned grid = (unsigned)((num_rows * 32 + block - 1) / block);
  switch (mode) {
    case GNNB200_AGG_SUM:
      aggregate_scalar_kernel<GNNB200_AGG_SUM><<<grid, block, 0, stream>>>(x, ldx, rowptr, col, num_rows, f, self_x, lds, eps, dinv, out, ldo, accumulate);
      break;
    case GNNB200_AGG_MEAN:
      aggregate_scalar_kernel<GNNB200_AGG_MEAN><<<grid, block, 0, stream>>>(x, ldx, rowptr, col, num_rows, f, self_x, lds, eps, dinv, out, ldo, accumulate);
      break;
    default:
      aggregate_scalar_kernel<GNNB200_AGG_GCN><<<grid, block, 0, stream>>>(x, ldx, rowptr, col, num_rows, f, self_x, lds, eps, dinv, out, ldo, accumulate);
      break;
  }
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_aggregate_long_rows_f32(const float* x, int64_t ldx, const int32_t* rowptr, const int32_t* col,
                                               const int64_t* rows, int64_t num_long_rows, int64_t feat, int mode,
                                               const float* self_x, int64_t lds, const float* eps, float* out,
                                               int64_t ldo, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int accumulate = (mode & GNNB200_AGG_ACCUMULATE) ? 1 : 0;
  mode &= ~(GNNB200_AGG_ACCUMULATE | GNNB200_AGG_SKIP_LONG);
  if (num_long_rows < 0 || feat < 0 || mode != GNNB200_AGG_SUM) return GNNB200_EINVAL;
  if (num_long_rows == 0 || feat == 0) return GNNB200_OK;
  if (!x || !rowptr || !col || !rows || !out) return GNNB200_EINVAL;
  if (num_long_rows >= (int64_t)INT32_MAX) return GNNB200_ERANGE;
  if (feat % 4 != 0 || feat > 1024 || ldx % 4 != 0 || ldo % 4 != 0 || (self_x && lds % 4 != 0) || (uintptr_t)x % 16 != 0 ||
      (uintptr_t)out % 16 != 0 || (uintptr_t)self_x % 16 != 0)
    return GNNB200_EUNSUPPORTED;
  const int f = (int)feat, nvec = f / 4;
  const size_t smem = (size_t)8 * nvec * sizeof(float4);
  const unsigned grid = (unsigned)num_long_rows;
#define GNNB200_LONG_ARGS x, ldx, rowptr, col, rows, f, self_x, lds, eps, out, ldo, accumulate
  if (nvec <= 32) aggregate_long_rows_kernel<1><<<grid, 256, smem, stream>>>(GNNB200_LONG_ARGS);
  else if (nvec <= 64) aggregate_long_rows_kernel<2><<<grid, 256, smem, stream>>>(GNNB200_LONG_ARGS);
  else if (nvec <= 128) aggregate_long_rows_kernel<4><<<grid, 256, smem, stream>>>(GNNB200_LONG_ARGS);
  else aggregate_long_rows_kernel<8><<<grid, 256, smem, stream>>>(GNNB200_LONG_ARGS);
#undef GNNB200_LONG_ARGS
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}
