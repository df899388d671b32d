// Device-side edge_index -> CSR/CSC build, sorted-batch -> ptr, and coalesce.
// Replaces the COO indexing of PyG GINConv.propagate (reference src/models/gnn.py:41) and
// to_undirected/coalesce (reference src/pretrain/tasks.py:108).  See include/gnnb200.h.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "common.cuh"

namespace gnnb200 {

// keys[i] = key row of column i (narrowed to int32), vals[i] = i.
__global__ void csr_keys_kernel(const int64_t* __restrict__ key_row, int32_t* __restrict__ keys,
                                int32_t* __restrict__ vals, int64_t E) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < E; i += stride) {
    keys[i] = (int32_t)key_row[i];
    vals[i] = (int32_t)i;
  }
}

// col[i] = other_row[perm[i]]; rowptr from the boundaries of the sorted keys:
// position i closes every row r with sorted[i-1] < r <= sorted[i] (rowptr[r] = i).
__global__ void csr_finish_kernel(const int64_t* __restrict__ other_row, const int32_t* __restrict__ sorted_keys,
                                  const int32_t* __restrict__ perm, int32_t* __restrict__ col,
                                  int32_t* __restrict__ rowptr, int64_t E, int64_t N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i <= E; i += stride) {
    int64_t prev = (i == 0) ? -1 : (int64_t)sorted_keys[i - 1];
    int64_t cur = (i == E) ? N : (int64_t)sorted_keys[i];
    if (cur > N) cur = N;  // out-of-range ids are undefined in the reference; stay in bounds
    if (i < E) col[i] = (int32_t)other_row[perm[i]];
    for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = (int32_t)i;
  }
}

__global__ void segment_ptr_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t S,
                                   int32_t* __restrict__ ptr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i <= n; i += stride) {
    int64_t prev = (i == 0) ? -1 : ids[i - 1];
    int64_t cur = (i == n) ? S : ids[i];
    if (cur > S) cur = S;
    for (int64_t r = prev + 1; r <= cur; ++r) ptr[r] = (int32_t)i;
  }
}

__global__ void coalesce_keys_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                                     int64_t* __restrict__ keys) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < E; i += stride) keys[i] = ei[i] * N + ei[E + i];
}

// flags[i] = 1 when sorted key i differs from its predecessor (first occurrence kept).
__global__ void coalesce_flags_kernel(const int64_t* __restrict__ keys, int64_t E, int32_t* __restrict__ flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < E; i += stride) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// pos = inclusive scan of flags; kept column i lands at pos[i]-1; out has capacity E columns.
__global__ void coalesce_emit_kernel(const int64_t* __restrict__ keys, const int32_t* __restrict__ flags,
                                     const int32_t* __restrict__ pos, int64_t E, int64_t N,
                                     int64_t* __restrict__ out, int64_t* __restrict__ out_count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < E; i += stride) {
    if (flags[i]) {
      int64_t k = keys[i];
      int64_t p = pos[i] - 1;
      out[p] = k / N;
      out[E + p] = k % N;
    }
    if (i == E - 1) *out_count = pos[i];
  }
}

__global__ void zero_i64_kernel(int64_t* p) { *p = 0; }

static inline int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)kNumSMs * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static inline int bits_for(int64_t n) {
  int b = 1;
  while (b < 63 && ((int64_t)1 << b) < n) ++b;
  return b;
}

}  // namespace gnnb200

using namespace gnnb200;

extern "C" int gnnb200_csr_build_i64(const int64_t* edge_index, int64_t E, int64_t N, int by_src,
                                     int32_t* rowptr, int32_t* col, int32_t* eid, void* workspace,
                                     size_t* workspace_bytes, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (E < 0 || N < 0 || !workspace_bytes) return GNNB200_EINVAL;
  if (E >= (int64_t)INT32_MAX || N >= (int64_t)INT32_MAX) return GNNB200_ERANGE;
  Workspace ws(workspace);
  int32_t* keys_in = ws.take<int32_t>(E);
  int32_t* keys_out = ws.take<int32_t>(E);
  int32_t* vals_in = ws.take<int32_t>(E);
  int32_t* vals_out = eid ? eid : ws.take<int32_t>(E);
  size_t cub_bytes = 0;
  int end_bit = bits_for(N > 1 ? N : 2);
  if (end_bit > 32) end_bit = 32;
  GNNB200_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys_in, keys_out, vals_in, vals_out,
                                                     (int)E, 0, end_bit, stream));
  char* cub_tmp = ws.take<char>(cub_bytes);
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;
  if (!rowptr || (E > 0 && (!edge_index || !col))) return GNNB200_EINVAL;
  const int64_t* key_row = edge_index + (by_src ? 0 : E);
  const int64_t* other_row = edge_index + (by_src ? E : 0);
  if (E > 0) {
    csr_keys_kernel<<<grid_for(E, 256), 256, 0, stream>>>(key_row, keys_in, vals_in, E);
    GNNB200_LAUNCH_CHECK();
    GNNB200_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, vals_in, vals_out,
                                                       (int)E, 0, end_bit, stream));
  }
  csr_finish_kernel<<<grid_for(E + 1, 256), 256, 0, stream>>>(other_row, keys_out, vals_out, col, rowptr, E, N);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_segment_ptr_i64(const int64_t* ids, int64_t n, int64_t S, int32_t* ptr,
                                       gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || S < 0 || !ptr || (n > 0 && !ids)) return GNNB200_EINVAL;
  if (n >= (int64_t)INT32_MAX || S >= (int64_t)INT32_MAX) return GNNB200_ERANGE;
  segment_ptr_kernel<<<grid_for(n + 1, 256), 256, 0, stream>>>(ids, n, S, ptr);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_coalesce_i64(const int64_t* edge_index, int64_t E, int64_t N, int64_t* out,
                                    int64_t* out_count, void* workspace, size_t* workspace_bytes,
                                    gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (E < 0 || N < 0 || !workspace_bytes) return GNNB200_EINVAL;
  if (E >= (int64_t)INT32_MAX) return GNNB200_ERANGE;
  Workspace ws(workspace);
  int64_t* keys_in = ws.take<int64_t>(E);
  int64_t* keys_out = ws.take<int64_t>(E);
  int32_t* flags = ws.take<int32_t>(E);
  int32_t* pos = ws.take<int32_t>(E);
  size_t sort_bytes = 0, scan_bytes = 0;
  int end_bit = 2 * bits_for(N > 1 ? N : 2);
  if (end_bit > 64) end_bit = 64;
  GNNB200_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, keys_in, keys_out, (int)E, 0, end_bit, stream));
  GNNB200_CHECK_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, flags, pos, (int)E, stream));
  size_t tmp_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  char* tmp = ws.take<char>(tmp_bytes);
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;
  if (!out_count || (E > 0 && (!edge_index || !out))) return GNNB200_EINVAL;
  if (E == 0) {
    zero_i64_kernel<<<1, 1, 0, stream>>>(out_count);
    GNNB200_LAUNCH_CHECK();
    return GNNB200_OK;
  }
  coalesce_keys_kernel<<<grid_for(E, 256), 256, 0, stream>>>(edge_index, E, N, keys_in);
  GNNB200_LAUNCH_CHECK();
  GNNB200_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(tmp, sort_bytes, keys_in, keys_out, (int)E, 0, end_bit, stream));
  coalesce_flags_kernel<<<grid_for(E, 256), 256, 0, stream>>>(keys_out, E, flags);
  GNNB200_LAUNCH_CHECK();
  GNNB200_CHECK_CUDA(cub::DeviceScan::InclusiveSum(tmp, scan_bytes, flags, pos, (int)E, stream));
  coalesce_emit_kernel<<<grid_for(E, 256), 256, 0, stream>>>(keys_out, flags, pos, E, N, out, out_count);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}
