// Device helpers shared by the elementwise / column-reduction kernels (bn.cu, reduce.cu, elementwise_v2.cu):
// the counter-based dropout stream, the recomputed BatchNorm/ReLU/dropout gradient mask, Chan's moment merge.
#pragma once
#include "common.cuh"

namespace gnnb200 {

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// keep flags for the 4 consecutive elements starting at linear index 4*q
__device__ __forceinline__ void keep4(uint64_t seed, uint64_t q, uint32_t thresh, bool (&k)[4]) {
  const uint4 r = philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
  k[0] = r.x < thresh; k[1] = r.y < thresh; k[2] = r.z < thresh; k[3] = r.w < thresh;
}

// g1 for element (r, c..c+3) recomputed from the saved pre-BN activation
template <bool DROP>
__device__ __forceinline__ void bn_g1(const float4 g, const float4 v, const float4 mu, const float4 is, const float4 ga,
                                      const float4 be, int relu, uint64_t seed, uint64_t q, uint32_t thresh, float scale,
                                      float (&g1)[4], float (&xh)[4]) {
  const float gv[4] = {g.x, g.y, g.z, g.w};
  xh[0] = (v.x - mu.x) * is.x; xh[1] = (v.y - mu.y) * is.y; xh[2] = (v.z - mu.z) * is.z; xh[3] = (v.w - mu.w) * is.w;
  const float gav[4] = {ga.x, ga.y, ga.z, ga.w};
  const float bev[4] = {be.x, be.y, be.z, be.w};
  bool k[4] = {true, true, true, true};
  if (DROP) keep4(seed, q, thresh, k);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t = gv[i];
    if (DROP) t = k[i] ? t * scale : 0.f;
    if (relu && !(xh[i] * gav[i] + bev[i] > 0.f)) t = 0.f;
    g1[i] = t;
  }
}

struct Moments {
  float n, sum, m2;  // count, plain sum, centred second moment
};

__device__ __forceinline__ Moments merge(Moments a, Moments b) {
  if (b.n == 0.f) return a;
  if (a.n == 0.f) return b;
  const float n = a.n + b.n;
  const float d = b.sum / b.n - a.sum / a.n;
  Moments r;
  r.n = n;
  r.sum = a.sum + b.sum;
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / n);
  return r;
}

}  // namespace gnnb200
