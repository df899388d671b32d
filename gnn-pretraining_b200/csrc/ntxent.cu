// NT-Xent / InfoNCE loss without materialising the [2M,2M] similarity matrix (K9 of SURVEY §2.4;
// reference src/pretrain/tasks.py:192-213 and :265-287):
//   zn = z / max(||z||, 1e-12);  S = zn zn^T / T, diagonal excluded;  loss = sum_i (lse_i - S[i,pos(i)]),
//   pos(i) = (i + M) mod 2M.
// Forward: tiled similarity + online log-sum-exp per row.  Backward: because S and pos() are
// symmetric, d/dzn_i = (g/T) * sum_{j != i} (exp(S_ij-lse_i) + exp(S_ij-lse_j) - 2[j==pos(i)]) zn_j,
// one more tiled pass, then the normalisation Jacobian.  fp32 FFMA version (1e-5 class).
#include <float.h>
#include "common.cuh"

namespace gnnb200 {

constexpr int TI = 32, TJ = 64, TK = 32;
constexpr float kNormEps = 1e-12f;  // F.normalize default

__global__ void __launch_bounds__(256)
ntx_normalize_kernel(const float* __restrict__ z, int64_t ldz, int rows, int dim, float* __restrict__ zn,
                     float* __restrict__ norm) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  float ss = 0.f;
  for (int k = lane; k < dim; k += 32) {
    const float v = z[(int64_t)r * ldz + k];
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float nrm = sqrtf(ss);
  const float den = fmaxf(nrm, kNormEps);
  for (int k = lane; k < dim; k += 32) zn[(int64_t)r * dim + k] = z[(int64_t)r * ldz + k] / den;
  if (lane == 0) norm[r] = nrm;
}

// 128 threads: ty = tid>>4 (8) x 4 rows, tx = tid&15 (16) x 4 cols.  S tile [TI x TJ].
__device__ __forceinline__ void sim_tile(const float* __restrict__ zn, int rows, int dim, int i0, int j0,
                                         float (*Zi)[TI + 1], float (*Zj)[TJ + 1], float acc[4][4]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int k0 = 0; k0 < dim; k0 += TK) {
    __syncthreads();
    for (int e = tid; e < TI * TK; e += 128) {
      const int k = e & (TK - 1), r = e >> 5;
      Zi[k][r] = (i0 + r < rows && k0 + k < dim) ? zn[(int64_t)(i0 + r) * dim + k0 + k] : 0.f;
    }
    for (int e = tid; e < TJ * TK; e += 128) {
      const int k = e & (TK - 1), r = e >> 5;
      Zj[k][r] = (j0 + r < rows && k0 + k < dim) ? zn[(int64_t)(j0 + r) * dim + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < TK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = Zi[k][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Zj[k][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
  }
}

__device__ __forceinline__ float half_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(128)
ntx_fwd_kernel(const float* __restrict__ zn, int rows, int half, int dim, float inv_t, float* __restrict__ lse,
               float* __restrict__ partial_loss) {
  __shared__ float Zi[TK][TI + 1];
  __shared__ float Zj[TK][TJ + 1];
  __shared__ float red[8];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.x * TI;
  float run_m[4], run_s[4], pos_v[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) { run_m[a] = -FLT_MAX; run_s[a] = 0.f; pos_v[a] = 0.f; }

  for (int j0 = 0; j0 < rows; j0 += TJ) {
    float acc[4][4];
    sim_tile(zn, rows, dim, i0, j0, Zi, Zj, acc);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + ty * 4 + a;
      const int pos = (i + half) % rows;
      float tile_m = -FLT_MAX;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j = j0 + tx * 4 + b;
        acc[a][b] *= inv_t;
        if (j < rows && j != i) tile_m = fmaxf(tile_m, acc[a][b]);
        if (j == pos) pos_v[a] = acc[a][b];
      }
      tile_m = half_max(tile_m);
      const float new_m = fmaxf(run_m[a], tile_m);
      float s = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j = j0 + tx * 4 + b;
        if (j < rows && j != i) s += __expf(acc[a][b] - new_m);
      }
      s = half_sum(s);
      run_s[a] = run_s[a] * __expf(run_m[a] - new_m) + s;
      run_m[a] = new_m;
    }
  }
  float local = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    const float p = half_sum(pos_v[a]);
    const float l = run_m[a] + logf(run_s[a]);
    if (i < rows && tx == 0) {
      lse[i] = l;
      local += l - p;
    }
  }
  // block sum in a fixed order: tx==0 lanes of the 8 row groups
  if (tx == 0) red[ty] = local;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int q = 0; q < 8; ++q) t += red[q];
    partial_loss[blockIdx.x] = t;
  }
}

__global__ void ntx_loss_finish_kernel(const float* __restrict__ partial, int count, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < count; ++i) t += partial[i];
    *loss = t;
  }
}

// Backward.  Dynamic smem: Zi [TK][TI+1], Zj [TK][TJ+1] (similarity), W [TI][TJ+1], ZjF [TJ][dim+1].
__global__ void __launch_bounds__(128)
ntx_bwd_kernel(const float* __restrict__ zn, const float* __restrict__ lse, const float* __restrict__ norm,
               const float* __restrict__ grad_loss, int rows, int half, int dim, float inv_t,
               float* __restrict__ grad_z, int64_t ldgz) {
  extern __shared__ float smem[];
  float (*Zi)[TI + 1] = reinterpret_cast<float (*)[TI + 1]>(smem);
  float (*Zj)[TJ + 1] = reinterpret_cast<float (*)[TJ + 1]>(smem + TK * (TI + 1));
  float (*W)[TJ + 1] = reinterpret_cast<float (*)[TJ + 1]>(smem + TK * (TI + 1) + TK * (TJ + 1));
  float* ZjF = smem + TK * (TI + 1) + TK * (TJ + 1) + TI * (TJ + 1);  // [TJ][dim+1]
  const int dpad = dim + 1;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.x * TI;
  const int ncol = dim / 16;  // columns per thread: tx + 16*c  (dim % 16 == 0, dim <= 128)

  float lse_i[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    lse_i[a] = (i < rows) ? lse[i] : 0.f;
  }
  float out[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) out[a][c] = 0.f;

  for (int j0 = 0; j0 < rows; j0 += TJ) {
    float acc[4][4];
    sim_tile(zn, rows, dim, i0, j0, Zi, Zj, acc);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + ty * 4 + a;
      const int pos = (i + half) % rows;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j = j0 + tx * 4 + b;
        float w = 0.f;
        if (i < rows && j < rows && j != i) {
          const float s = acc[a][b] * inv_t;
          w = __expf(s - lse_i[a]) + __expf(s - lse[j]) - (j == pos ? 2.f : 0.f);
        }
        W[ty * 4 + a][tx * 4 + b] = w;
      }
    }
    for (int e = tid; e < TJ * dim; e += 128) {
      const int k = e % dim, r = e / dim;
      ZjF[r * dpad + k] = (j0 + r < rows) ? zn[(int64_t)(j0 + r) * dim + k] : 0.f;
    }
    __syncthreads();
    for (int j = 0; j < TJ; ++j) {
      float wv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) wv[a] = W[ty * 4 + a][j];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < ncol) {
          const float zv = ZjF[j * dpad + tx + 16 * c];
#pragma unroll
          for (int a = 0; a < 4; ++a) out[a][c] = fmaf(wv[a], zv, out[a][c]);
        }
      }
    }
    // next sim_tile starts with __syncthreads(), which also protects W / ZjF
  }

  const float g = (*grad_loss) * inv_t;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    float dotv = 0.f;
    float zi[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      zi[c] = 0.f;
      if (c < ncol && i < rows) {
        zi[c] = zn[(int64_t)i * dim + tx + 16 * c];
        out[a][c] *= g;
        dotv = fmaf(zi[c], out[a][c], dotv);
      }
    }
    dotv = half_sum(dotv);
    if (i < rows) {
      const float nrm = norm[i];
      const bool clamped = nrm < kNormEps;
      const float den = fmaxf(nrm, kNormEps);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < ncol) {
          const float d = clamped ? out[a][c] : (out[a][c] - zi[c] * dotv);
          grad_z[(int64_t)i * ldgz + tx + 16 * c] = d / den;
        }
      }
    }
  }
}

// ---- tensor-core variant: the similarity matrix S = zn zn^T is produced by the tcgen05 GEMM (gemm_tcgen05.cu) and
// consumed by the two row kernels below (large contrastive batches: the fused FFMA kernel above is O((2M)^2 D) FFMA work) ----

// One warp per row of S [R, R]: lse_i = logsumexp_{j != i}(S_ij / T), loss_i = lse_i - S[i, pos(i)] / T.
__global__ void __launch_bounds__(256)
ntx_rows_lse_kernel(const float* __restrict__ S, int64_t lds, int rows, int half, float inv_t, float* __restrict__ lse,
                    float* __restrict__ row_loss) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= rows) return;
  const float* row = S + (int64_t)i * lds;
  float m = -FLT_MAX, s = 0.f;
  for (int j = lane; j < rows; j += 32) {
    if (j == i) continue;
    const float v = row[j] * inv_t;
    const float nm = fmaxf(m, v);
    s = s * __expf(m - nm) + __expf(v - nm);
    m = nm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(m, om);
    s = s * __expf(m - nm) + os * __expf(om - nm);
    m = nm;
  }
  if (lane == 0) {
    const float l = m + logf(s);
    lse[i] = l;
    row_loss[i] = l - row[(i + half) % rows] * inv_t;
  }
}

__global__ void __launch_bounds__(256)
ntx_sum_rows_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  __shared__ float sm[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += v[i];     // fixed strided order
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sm[0];
}

// In place: S_ij <- g/T * (exp(S_ij/T - lse_i) - [j == pos(i)]) for j != i, 0 on the diagonal  (= dL/dS_ij).
__global__ void __launch_bounds__(256)
ntx_dsim_kernel(float* __restrict__ S, int64_t lds, int rows, int half, float inv_t, const float* __restrict__ lse,
                const float* __restrict__ grad_loss) {
  const int64_t total = (int64_t)rows * rows;
  const float g = (*grad_loss) * inv_t;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(q / rows), j = (int)(q - (int64_t)i * rows);
    float* p = S + (int64_t)i * lds + j;
    float w = 0.f;
    if (j != i) w = g * (__expf(*p * inv_t - lse[i]) - (j == (i + half) % rows ? 1.f : 0.f));
    *p = w;
  }
}

// dz = (dzn - zn * <zn, dzn>) / max(|z|, eps)  (Jacobian of F.normalize); one warp per row.
__global__ void __launch_bounds__(256)
ntx_normalize_bwd_kernel(const float* __restrict__ zn, const float* __restrict__ dzn, int64_t ldd,
                         const float* __restrict__ norm, int rows, int dim, float* __restrict__ dz, int64_t ldz) {
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  float d = 0.f;
  for (int k = lane; k < dim; k += 32) d = fmaf(zn[(int64_t)r * dim + k], dzn[(int64_t)r * ldd + k], d);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  const float nrm = norm[r];
  const bool clamped = nrm < kNormEps;
  const float den = fmaxf(nrm, kNormEps);
  for (int k = lane; k < dim; k += 32) {
    const float g = dzn[(int64_t)r * ldd + k];
    dz[(int64_t)r * ldz + k] = (clamped ? g : (g - zn[(int64_t)r * dim + k] * d)) / den;
  }
}

}  // namespace gnnb200

using namespace gnnb200;

extern "C" int gnnb200_normalize_rows_f32(const float* z, int64_t ldz, int64_t rows, int64_t dim, float* zn, float* norm,
                                          gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows < 0 || dim < 0) return GNNB200_EINVAL;
  if (rows == 0 || dim == 0) return GNNB200_OK;
  if (!z || !zn || !norm) return GNNB200_EINVAL;
  ntx_normalize_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, stream>>>(z, ldz, (int)rows, (int)dim, zn, norm);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_normalize_rows_bwd_f32(const float* zn, const float* grad_zn, int64_t ldg, const float* norm,
                                              int64_t rows, int64_t dim, float* grad_z, int64_t ldz,
                                              gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows < 0 || dim < 0) return GNNB200_EINVAL;
  if (rows == 0 || dim == 0) return GNNB200_OK;
  if (!zn || !grad_zn || !norm || !grad_z) return GNNB200_EINVAL;
  ntx_normalize_bwd_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, stream>>>(zn, grad_zn, ldg, norm, (int)rows,
                                                                                   (int)dim, grad_z, ldz);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_ntxent_sim_fwd_f32(const float* sim, int64_t lds, int64_t two_m, float temperature, float* lse,
                                          float* row_loss, float* loss, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (two_m < 0 || (two_m & 1)) return GNNB200_EINVAL;
  if (two_m >= (1 << 24)) return GNNB200_ERANGE;
  if (!loss || (two_m > 0 && (!sim || !lse || !row_loss)) || !(temperature > 0.f)) return GNNB200_EINVAL;
  if (two_m > 0) {
    ntx_rows_lse_kernel<<<(unsigned)((two_m * 32 + 255) / 256), 256, 0, stream>>>(sim, lds, (int)two_m, (int)(two_m / 2),
                                                                                 1.0f / temperature, lse, row_loss);
    GNNB200_LAUNCH_CHECK();
  }
  ntx_sum_rows_kernel<<<1, 256, 0, stream>>>(row_loss, (int)two_m, loss);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_ntxent_sim_bwd_f32(float* sim, int64_t lds, int64_t two_m, float temperature, const float* lse,
                                          const float* grad_loss, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (two_m < 0 || (two_m & 1)) return GNNB200_EINVAL;
  if (two_m == 0) return GNNB200_OK;
  if (two_m >= (1 << 24)) return GNNB200_ERANGE;
  if (!sim || !lse || !grad_loss || !(temperature > 0.f)) return GNNB200_EINVAL;
  int64_t blocks = (two_m * two_m + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  ntx_dsim_kernel<<<(unsigned)blocks, 256, 0, stream>>>(sim, lds, (int)two_m, (int)(two_m / 2), 1.0f / temperature, lse,
                                                        grad_loss);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}


extern "C" int gnnb200_ntxent_fwd_f32(const float* z, int64_t ldz, int64_t two_m, int64_t dim, float temperature,
                                      float* zn, float* lse, float* norm, float* loss, void* workspace,
                                      size_t* workspace_bytes, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (two_m < 0 || dim < 0 || (two_m & 1) || !workspace_bytes) return GNNB200_EINVAL;
  if (two_m >= (1 << 24)) return GNNB200_ERANGE;
  const int blocks = (int)((two_m + TI - 1) / TI);
  Workspace ws(workspace);
  float* partial = ws.take<float>(blocks > 0 ? blocks : 1);
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;
  if (!loss) return GNNB200_EINVAL;
  if (two_m == 0) {
    ntx_loss_finish_kernel<<<1, 32, 0, stream>>>(partial, 0, loss);
    GNNB200_LAUNCH_CHECK();
    return GNNB200_OK;
  }
  if (!z || !zn || !lse || !norm || !(temperature > 0.f)) return GNNB200_EINVAL;
  ntx_normalize_kernel<<<(unsigned)((two_m * 32 + 255) / 256), 256, 0, stream>>>(z, ldz, (int)two_m, (int)dim, zn, norm);
  GNNB200_LAUNCH_CHECK();
  ntx_fwd_kernel<<<blocks, 128, 0, stream>>>(zn, (int)two_m, (int)(two_m / 2), (int)dim, 1.0f / temperature, lse, partial);
  GNNB200_LAUNCH_CHECK();
  ntx_loss_finish_kernel<<<1, 32, 0, stream>>>(partial, blocks, loss);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_ntxent_bwd_f32(const float* zn, const float* lse, const float* norm, const float* grad_loss,
                                      int64_t two_m, int64_t dim, float temperature, float* grad_z, int64_t ldgz,
                                      gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (two_m < 0 || dim < 0 || (two_m & 1)) return GNNB200_EINVAL;
  if (two_m == 0 || dim == 0) return GNNB200_OK;
  if (dim % 16 != 0 || dim > 128) return GNNB200_EUNSUPPORTED;
  if (!zn || !lse || !norm || !grad_loss || !grad_z || !(temperature > 0.f)) return GNNB200_EINVAL;
  const size_t smem = sizeof(float) * (TK * (TI + 1) + TK * (TJ + 1) + TI * (TJ + 1) + TJ * (dim + 1));
  GNNB200_CHECK_CUDA(cudaFuncSetAttribute(ntx_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = (int)((two_m + TI - 1) / TI);
  ntx_bwd_kernel<<<blocks, 128, smem, stream>>>(zn, lse, norm, grad_loss, (int)two_m, (int)(two_m / 2), (int)dim,
                                               1.0f / temperature, grad_z, ldgz);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}
