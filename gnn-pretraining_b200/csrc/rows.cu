// Row gather / scatter (K12 of SURVEY §2.4): NFM mask-token write and target gather
// (reference src/models/pretrain_model.py:84-86) and h[mask_indices] (src/pretrain/tasks.py:82).
// One warp per index row, 128-bit accesses when the layout allows.
#include "common.cuh"

namespace gnnb200 {

template <bool SCATTER>
__global__ void __launch_bounds__(256)
rows_move_kernel(const float* __restrict__ src, int64_t lds, int broadcast, const int64_t* __restrict__ idx,
                 int64_t num_idx, int feat, int vec, float* __restrict__ dst, int64_t ldd) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= num_idx) return;
  const int64_t r = idx[i];
  const float* s = SCATTER ? (src + (broadcast ? 0 : i * lds)) : (src + r * lds);
  float* d = SCATTER ? (dst + r * ldd) : (dst + i * ldd);
  if (vec) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    float4* d4 = reinterpret_cast<float4*>(d);
    for (int k = lane; k < (feat >> 2); k += 32) d4[k] = __ldg(s4 + k);
  } else {
    for (int k = lane; k < feat; k += 32) d[k] = __ldg(s + k);
  }
}

// Deterministic backward of gather: the caller passes the indices grouped per destination row
// (CSR over idx: rowptr [num_rows+1], eid = positions into grad_out in ascending order).
__global__ void __launch_bounds__(256)
rows_gather_bwd_kernel(const float* __restrict__ g, int64_t ldg, const int32_t* __restrict__ rowptr,
                       const int32_t* __restrict__ eid, int64_t num_rows, int feat, float* __restrict__ gx,
                       int64_t ldgx) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= num_rows) return;
  const int beg = rowptr[r], end = rowptr[r + 1];
  for (int k = lane; k < feat; k += 32) {
    float acc = 0.f;
    for (int e = beg; e < end; ++e) acc = __fadd_rn(acc, __ldg(g + (int64_t)eid[e] * ldg + k));
    gx[r * ldgx + k] = acc;
  }
}

// ---- link-prediction decoder features (K8; reference src/models/heads.py:59-65) ------------------
// feat[e] = [h[u]+h[v], h[u]*h[v], |h[u]-h[v]|]; one warp per edge.
__global__ void __launch_bounds__(256)
lp_features_kernel(const float* __restrict__ h, int64_t ldh, const int64_t* __restrict__ edges, int64_t E,
                   int hidden, float* __restrict__ feat, int64_t ldf) {
  const int lane = threadIdx.x & 31;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= E) return;
  const float* hu = h + edges[e] * ldh;
  const float* hv = h + edges[E + e] * ldh;
  float* o = feat + e * ldf;
  for (int k = lane; k < hidden; k += 32) {
    const float a = __ldg(hu + k), b = __ldg(hv + k);
    o[k] = __fadd_rn(a, b);
    o[hidden + k] = __fmul_rn(a, b);
    o[2 * hidden + k] = fabsf(__fsub_rn(a, b));
  }
}

// Backward into h: node n collects, in edge order, first the edges where it is u, then those where
// it is v (two CSRs over the decoder's edge list, built with gnnb200_csr_build_i64).
//   d/dh[u] = g_sum + g_prod*h[v] + sign(h[u]-h[v])*g_diff ;  d/dh[v] = g_sum + g_prod*h[u] - sign(.)*g_diff
__device__ __forceinline__ float sgnf(float d) { return (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(256)
lp_features_bwd_kernel(const float* __restrict__ h, int64_t ldh, const int64_t* __restrict__ edges, int64_t E,
                       int hidden, const float* __restrict__ gfeat, int64_t ldf,
                       const int32_t* __restrict__ u_ptr, const int32_t* __restrict__ u_eid,
                       const int32_t* __restrict__ v_ptr, const int32_t* __restrict__ v_eid, int64_t num_nodes,
                       float* __restrict__ gh, int64_t ldgh) {
  const int lane = threadIdx.x & 31;
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= num_nodes) return;
  const float* hn = h + n * ldh;
  for (int k = lane; k < hidden; k += 32) {
    const float self = __ldg(hn + k);
    float acc = 0.f;
    for (int p = u_ptr[n]; p < u_ptr[n + 1]; ++p) {
      const int64_t e = u_eid[p];
      const float other = __ldg(h + edges[E + e] * ldh + k);
      const float* g = gfeat + e * ldf;
      acc += __ldg(g + k) + __ldg(g + hidden + k) * other + sgnf(self - other) * __ldg(g + 2 * hidden + k);
    }
    for (int p = v_ptr[n]; p < v_ptr[n + 1]; ++p) {
      const int64_t e = v_eid[p];
      const float other = __ldg(h + edges[e] * ldh + k);
      const float* g = gfeat + e * ldf;
      acc += __ldg(g + k) + __ldg(g + hidden + k) * other - sgnf(other - self) * __ldg(g + 2 * hidden + k);
    }
    gh[n * ldgh + k] = acc;
  }
}

}  // namespace gnnb200

using namespace gnnb200;

static inline int vec_ok(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t feat) {
  return feat % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0);
}

extern "C" int gnnb200_rows_gather_f32(const float* x, int64_t ldx, const int64_t* idx, int64_t num_idx,
                                       int64_t feat, float* out, int64_t ldo, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_idx < 0 || feat < 0) return GNNB200_EINVAL;
  if (num_idx == 0 || feat == 0) return GNNB200_OK;
  if (!x || !idx || !out) return GNNB200_EINVAL;
  const unsigned grid = (unsigned)((num_idx * 32 + 255) / 256);
  rows_move_kernel<false><<<grid, 256, 0, stream>>>(x, ldx, 0, idx, num_idx, (int)feat, vec_ok(x, ldx, out, ldo, feat), out, ldo);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_rows_scatter_f32(const float* src, int64_t lds, int broadcast, const int64_t* idx,
                                        int64_t num_idx, int64_t feat, float* out, int64_t ldo,
                                        gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_idx < 0 || feat < 0) return GNNB200_EINVAL;
  if (num_idx == 0 || feat == 0) return GNNB200_OK;
  if (!src || !idx || !out) return GNNB200_EINVAL;
  const unsigned grid = (unsigned)((num_idx * 32 + 255) / 256);
  rows_move_kernel<true><<<grid, 256, 0, stream>>>(src, lds, broadcast, idx, num_idx, (int)feat, vec_ok(src, lds, out, ldo, feat), out, ldo);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_rows_gather_bwd_f32(const float* grad_out, int64_t ldg, const int32_t* rowptr,
                                           const int32_t* eid, int64_t num_rows, int64_t feat, float* grad_x,
                                           int64_t ldgx, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_rows < 0 || feat < 0) return GNNB200_EINVAL;
  if (num_rows == 0 || feat == 0) return GNNB200_OK;
  if (!grad_out || !rowptr || !eid || !grad_x) return GNNB200_EINVAL;
  const unsigned grid = (unsigned)((num_rows * 32 + 255) / 256);
  rows_gather_bwd_kernel<<<grid, 256, 0, stream>>>(grad_out, ldg, rowptr, eid, num_rows, (int)feat, grad_x, ldgx);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_lp_features_f32(const float* h, int64_t ldh, const int64_t* edges, int64_t num_edges,
                                       int64_t hidden, float* feat, int64_t ldf, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_edges < 0 || hidden < 0) return GNNB200_EINVAL;
  if (num_edges == 0 || hidden == 0) return GNNB200_OK;
  if (!h || !edges || !feat) return GNNB200_EINVAL;
  const unsigned grid = (unsigned)((num_edges * 32 + 255) / 256);
  lp_features_kernel<<<grid, 256, 0, stream>>>(h, ldh, edges, num_edges, (int)hidden, feat, ldf);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_lp_features_bwd_f32(const float* h, int64_t ldh, const int64_t* edges, int64_t num_edges,
                                           int64_t hidden, const float* grad_feat, int64_t ldf,
                                           const int32_t* u_ptr, const int32_t* u_eid, const int32_t* v_ptr,
                                           const int32_t* v_eid, int64_t num_nodes, float* grad_h, int64_t ldgh,
                                           gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_edges < 0 || hidden < 0 || num_nodes < 0) return GNNB200_EINVAL;
  if (num_nodes == 0 || hidden == 0) return GNNB200_OK;
  if (!h || !grad_h || !u_ptr || !v_ptr || (num_edges > 0 && (!edges || !grad_feat || !u_eid || !v_eid))) return GNNB200_EINVAL;
  const unsigned grid = (unsigned)((num_nodes * 32 + 255) / 256);
  lp_features_bwd_kernel<<<grid, 256, 0, stream>>>(h, ldh, edges, num_edges, (int)hidden, grad_feat, ldf, u_ptr, u_eid,
                                                  v_ptr, v_eid, num_nodes, grad_h, ldgh);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}
