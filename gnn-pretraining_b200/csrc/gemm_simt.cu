// fp32 FFMA GEMM (the 1e-5 tolerance class of K3): C = op(A) op(B) (+bias)(ReLU).
// Stands in for nn.Linear forward/backward (reference src/models/gnn.py:13,31,34,
// src/models/heads.py:41) where fp32-class accuracy is required or the tcgen05 path's layout
// rules are not met.  64x64x16 tiles, 256 threads x (4x4) outputs, both operand layouts loaded
// coalesced; split-K through a partial buffer with a fixed-order finish pass (deterministic).
#include "common.cuh"

namespace gnnb200 {

constexpr int BM = 64, BN = 64, BK = 16;

// A(m,k) = A[m*sam + k*sak]; B(k,n) = B[k*sbk + n*sbn].
template <bool A_KCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbk,
                 int64_t sbn, float* __restrict__ C, int64_t ldc, int M, int N, int64_t K, int64_t k_per_split,
                 const float* __restrict__ bias, const float* __restrict__ residual, int64_t ldr, int relu,
                 float* __restrict__ partial) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = min(K, kbeg + k_per_split);
  const int tx = tid & 15, ty = tid >> 4;  // 16x16 threads, each 4x4 outputs

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto load_tiles = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;  // 1024 elements per tile
      int m, k;
      if (A_KCONTIG) { k = e & (BK - 1); m = e >> 4; } else { m = e & (BM - 1); k = e >> 6; }
      const int gm = m0 + m;
      const int64_t gk = k0 + k;
      ra[i] = (gm < M && gk < kend) ? __ldg(A + (int64_t)gm * sam + gk * sak) : 0.f;
      int n, kb;
      if (B_NCONTIG) { n = e & (BN - 1); kb = e >> 6; } else { kb = e & (BK - 1); n = e >> 4; }
      const int gn = n0 + n;
      const int64_t gkb = k0 + kb;
      rb[i] = (gn < N && gkb < kend) ? __ldg(B + gkb * sbk + (int64_t)gn * sbn) : 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (A_KCONTIG) { k = e & (BK - 1); m = e >> 4; } else { m = e & (BM - 1); k = e >> 6; }
      As[buf][k][m] = ra[i];
      int n, kb;
      if (B_NCONTIG) { n = e & (BN - 1); kb = e >> 6; } else { kb = e & (BK - 1); n = e >> 4; }
      Bs[buf][kb][n] = rb[i];
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    load_tiles(kbeg);
    store_tiles(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK) < kend;
    if (more) load_tiles(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) store_tiles(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      if (partial) {
        partial[((int64_t)blockIdx.z * M + gm) * N + gn] = acc[i][j];
      } else {
        float v = acc[i][j];
        if (bias) v += bias[gn];
        if (residual) v += residual[(int64_t)gm * ldr + gn];
        if (relu) v = fmaxf(v, 0.f);
        C[(int64_t)gm * ldc + gn] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
gemm_splitk_finish_kernel(const float* __restrict__ partial, int splits, int M, int N, const float* __restrict__ bias,
                          const float* __restrict__ residual, int64_t ldr, int relu, float* __restrict__ C, int64_t ldc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += partial[(int64_t)s * M * N + i];
  if (bias) v += bias[n];
  if (residual) v += residual[(int64_t)m * ldr + n];
  if (relu) v = fmaxf(v, 0.f);
  C[(int64_t)m * ldc + n] = v;
}

int gemm_simt_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  if (tiles >= kNumSMs || K < 4096) return 1;
  int64_t s = (2 * kNumSMs + tiles - 1) / tiles;
  const int64_t max_s = K / 1024;
  if (s > max_s) s = max_s;
  if (s > 256) s = 256;
  return (int)(s < 1 ? 1 : s);
}

int gemm_simt(const float* A, int64_t lda, int transa, const float* B, int64_t ldb, int transb, float* C,
              int64_t ldc, int64_t M, int64_t N, int64_t K, const float* bias, const float* residual, int64_t ldr,
              int epilogue, void* workspace, size_t* workspace_bytes, cudaStream_t stream) {
  const int splits = gemm_simt_splits(M, N, K);
  Workspace ws(workspace);
  float* partial = splits > 1 ? ws.take<float>((size_t)splits * M * N) : nullptr;
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;
  if (M == 0 || N == 0) return GNNB200_OK;
  // A is [M,K] (sam=lda, sak=1) or stored [K,M] (sam=1, sak=lda); B is [K,N] or stored [N,K].
  const int64_t sam = transa ? 1 : lda, sak = transa ? lda : 1;
  const int64_t sbk = transb ? 1 : ldb, sbn = transb ? ldb : 1;
  int64_t kps = (K + splits - 1) / splits;
  kps = (kps + BK - 1) / BK * BK;
  if (kps < BK) kps = BK;
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)splits);
  if (grid.y > 65535) return GNNB200_ERANGE;
  const int relu = (epilogue & GNNB200_EPI_RELU) ? 1 : 0;
  const bool ak = !transa, bn = !transb;
#define GNNB200_SIMT_LAUNCH(AK, BNC)                                                                             \
  gemm_simt_kernel<AK, BNC><<<grid, 256, 0, stream>>>(A, sam, sak, B, sbk, sbn, C, ldc, (int)M, (int)N, K, kps, \
                                                      bias, residual, ldr, relu, partial)
  if (ak && bn) GNNB200_SIMT_LAUNCH(true, true);
  else if (ak && !bn) GNNB200_SIMT_LAUNCH(true, false);
  else if (!ak && bn) GNNB200_SIMT_LAUNCH(false, true);
  else GNNB200_SIMT_LAUNCH(false, false);
#undef GNNB200_SIMT_LAUNCH
  GNNB200_LAUNCH_CHECK();
  if (splits > 1) {
    const int64_t total = M * N;
    gemm_splitk_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(partial, splits, (int)M, (int)N,
                                                                                  bias, residual, ldr, relu, C, ldc);
    GNNB200_LAUNCH_CHECK();
  }
  return GNNB200_OK;
}

}  // namespace gnnb200
