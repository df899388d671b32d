// Node-partitioned aggregation with the halo read IN the gather kernel over NVLink peer memory
// (BASELINE config 5).  Rank r owns a contiguous row range; every rank publishes its rows in a
// peer-mapped buffer (gnnb200_peer_alloc / gnnb200_peer_open: CUDA IPC, one process per GPU) and the gather
// reads neighbour rows straight from the owner's HBM: local rows at HBM speed, remote rows as 128-bit NVLink
// loads issued by the same warps, so the transfer overlaps the sums row by row — no all-gather, no packed
// send buffer, no staging copy of the halo.  Right for graphs with locality (only referenced rows travel, once
// per referencing edge); for a uniform-random graph the all-gather path moves fewer bytes (every remote row is
// referenced ~deg/P times per rank) and stays the default (gnnb200/partition.py).
//
// Column ids are pre-encoded by the host (partition.py): bits 31..28 = slot of the owning rank in `peer_x`,
// bits 27..0 = row inside that rank's buffer.  Same accumulation rule as aggregate.cu: strictly in edge order
// from 0.0f with __fadd_rn, no float atomics => the same bits as the single-device kernel on the same rows.
#include <string.h>
#include "common.cuh"

namespace gnnb200 {

constexpr int kPeerShift = 28;
constexpr int kPeerRowMask = (1 << kPeerShift) - 1;

template <int V, int U, int MINB>
__global__ void __launch_bounds__(256, MINB)
aggregate_peer_kernel(const float* const* __restrict__ peer_x, int64_t ldx, const int32_t* __restrict__ rowptr,
                      const int32_t* __restrict__ col, int64_t num_rows, int feat,
                      const float* __restrict__ self_x, int64_t lds, const float* __restrict__ eps_ptr,
                      float* __restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // one warp per destination row
  if (row >= num_rows) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  const int nvec = feat >> 2;

  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int base = beg; base < end; base += 32) {
    // every lane resolves ITS neighbour's row address once (owner table lookup off the per-neighbour critical
    // path: one L1-resident 8-byte load per lane and 32 neighbours), the inner loop only broadcasts addresses
    const int my_col = (base + lane < end) ? col[base + lane] : 0;
    const unsigned long long my_src =
        __ldg(reinterpret_cast<const unsigned long long*>(peer_x) + ((unsigned)my_col >> kPeerShift)) +
        (unsigned long long)(my_col & kPeerRowMask) * (unsigned long long)ldx * sizeof(float);
    const int cnt = min(32, end - base);
    for (int j = 0; j < cnt; j += U) {
      float4 nb[U][V];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = min(j + u, cnt - 1);
        const float4* src = reinterpret_cast<const float4*>(__shfl_sync(0xffffffffu, my_src, jj));
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

  const float scale = (self_x != nullptr) ? __fadd_rn(1.0f, eps_ptr ? *eps_ptr : 0.f) : 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int k = lane + v * 32;
    if (k < nvec) {
      float4 r = acc[v];
      if (self_x != nullptr) {
        const float4 s = *reinterpret_cast<const float4*>(self_x + row * lds + 4 * k);
        r = f4_add(r, f4_scale(scale, s));
      }
      *reinterpret_cast<float4*>(out + row * ldo + 4 * k) = r;
    }
  }
}

template <int V>
static int launch_peer(const float* const* peer_x, int64_t ldx, const int32_t* rowptr, const int32_t* col, int64_t num_rows,
                       int feat, const float* self_x, int64_t lds, const float* eps, float* out, int64_t ldo,
                       cudaStream_t stream) {
  // same register budget as aggregate.cu (occupancy beats per-warp memory-level parallelism for this gather)
  constexpr int U = (V >= 8) ? 1 : 2;
  constexpr int MINB = (V <= 2) ? 6 : ((V == 4) ? 3 : 2);
  const unsigned grid = (unsigned)((num_rows * 32 + 255) / 256);
  aggregate_peer_kernel<V, U, MINB><<<grid, 256, 0, stream>>>(peer_x, ldx, rowptr, col, num_rows, feat, self_x, lds, eps,
                                                              out, ldo);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

}  // namespace gnnb200

using namespace gnnb200;

extern "C" int gnnb200_aggregate_peer_f32(const float* const* peer_x, int num_peers, int64_t ldx, const int32_t* rowptr,
                                          const int32_t* col, int64_t num_rows, int64_t feat, const float* self_x,
                                          int64_t lds, const float* eps, float* out, int64_t ldo,
                                          gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  // num_peers only bounds the slot field of `col` (documentation of the table's length); the table itself is
  // device memory and is not read on the host
  if (num_rows < 0 || feat < 0 || num_peers < 1 || num_peers > GNNB200_MAX_PEERS || !peer_x) return GNNB200_EINVAL;
  if (num_rows == 0 || feat == 0) return GNNB200_OK;
  if (!rowptr || !out) return GNNB200_EINVAL;
  if (num_rows >= (int64_t)INT32_MAX) return GNNB200_ERANGE;
  if (feat % 4 != 0 || feat > 1024 || ldx % 4 != 0 || ldo % 4 != 0 || (self_x && lds % 4 != 0) ||
      (uintptr_t)out % 16 != 0 || (uintptr_t)self_x % 16 != 0)
    return GNNB200_EUNSUPPORTED;
  const int f = (int)feat, nvec = f / 4;
#define GNNB200_PEER_ARGS peer_x, ldx, rowptr, col, num_rows, f, self_x, lds, eps, out, ldo, stream
  if (nvec <= 32) return launch_peer<1>(GNNB200_PEER_ARGS);
  if (nvec <= 64) return launch_peer<2>(GNNB200_PEER_ARGS);
  if (nvec <= 128) return launch_peer<4>(GNNB200_PEER_ARGS);
  return launch_peer<8>(GNNB200_PEER_ARGS);
#undef GNNB200_PEER_ARGS
}

// Publish: strided device-to-device copy of this rank's rows into its peer-mapped buffer (stream-ordered).
extern "C" int gnnb200_peer_publish_f32(const float* src, int64_t lds, int64_t rows, int64_t feat, float* dst, int64_t ldd,
                                        gnnb200_stream_t stream_) {
  if (rows < 0 || feat < 0 || lds < feat || ldd < feat) return GNNB200_EINVAL;
  if (rows == 0 || feat == 0) return GNNB200_OK;
  if (!src || !dst) return GNNB200_EINVAL;
  GNNB200_CHECK_CUDA(cudaMemcpy2DAsync(dst, (size_t)ldd * sizeof(float), src, (size_t)lds * sizeof(float),
                                       (size_t)feat * sizeof(float), (size_t)rows, cudaMemcpyDeviceToDevice,
                                       (cudaStream_t)stream_));
  return GNNB200_OK;
}

// Pull a peer's published block into local memory with the copy engines (no SM time): the 'peercopy' exchange.
extern "C" int gnnb200_peer_copy_f32(float* dst, const float* src, int64_t count, gnnb200_stream_t stream_) {
  if (count < 0) return GNNB200_EINVAL;
  if (count == 0) return GNNB200_OK;
  if (!dst || !src) return GNNB200_EINVAL;
  GNNB200_CHECK_CUDA(cudaMemcpyAsync(dst, src, (size_t)count * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream_));
  return GNNB200_OK;
}

// ---- peer-mapped buffers (CUDA IPC; the only entry points of the library that allocate) -------------------

static_assert(sizeof(cudaIpcMemHandle_t) == GNNB200_PEER_HANDLE_BYTES, "handle size");

extern "C" int gnnb200_peer_alloc(size_t bytes, void** ptr, unsigned char* handle) {
  if (!ptr || !handle || bytes == 0) return GNNB200_EINVAL;
  void* p = nullptr;
  GNNB200_CHECK_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  memcpy(handle, &h, sizeof(h));
  *ptr = p;
  return GNNB200_OK;
}

extern "C" int gnnb200_peer_open(const unsigned char* handle, void** ptr) {
  if (!ptr || !handle) return GNNB200_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  // maps the exporting GPU's allocation into the CURRENT device's address space and enables peer access to it
  GNNB200_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GNNB200_OK;
}

extern "C" int gnnb200_peer_close(void* ptr) {
  if (!ptr) return GNNB200_EINVAL;
  GNNB200_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return GNNB200_OK;
}

extern "C" int gnnb200_peer_free(void* ptr) {
  if (!ptr) return GNNB200_EINVAL;
  GNNB200_CHECK_CUDA(cudaFree(ptr));
  return GNNB200_OK;
}
