// Batched negative sampling, deterministic branch, on the device (reference src/pretrain/tasks.py:107-111 ->
// PyG batched_negative_sampling / negative_sampling, App. A.5).  At the reference's call site the per-graph quota is the
// WHOLE batch's edge count, so for TU-sized graphs the sampler's candidate draw is arange(population): the result is "every
// non-edge of the graph in ascending code order, truncated to the quota" and no random number is consumed.  That branch is a
// bitmap complement:
//   count : one block per graph marks the codes row * (n - 1) + col' of its off-diagonal edges in a shared-memory bitmap
//           over the n (n - 1) code space, evaluates upstream's branch condition in the same double arithmetic
//           (k = int(1.1 * quota / (1 - codes / population)); deterministic iff population <= k) and counts the free codes;
//           a graph that would draw from Python's random.sample, or does not fit the bitmap, raises `needs_host`
//   write : the same bitmap again; free codes are ranked by word-level prefix popcounts and decoded in ascending order
//           into out[:, offset_g + rank] (rank < quota), with the graph's node offset added back.
// Bit-exact with the oracle (integer work).  The caller scans the counts and reads (total, needs_host) back once.
#include "common.cuh"

namespace gnnb200 {
namespace {

constexpr int kNegThreads = 256;
constexpr int kNegMaxNodes = 724;                        // n (n - 1) <= 2^19 bits = 64 KB of shared memory
constexpr int kNegMaxWords = (1 << 19) / 32;

struct GraphView {
  int64_t node_lo, n, e_lo, e_hi, population;
};

__device__ __forceinline__ GraphView view_of(const int32_t* node_ptr, const int32_t* edge_ptr, int g) {
  GraphView v;
  v.node_lo = node_ptr[g];
  v.n = node_ptr[g + 1] - v.node_lo;
  v.e_lo = edge_ptr[g];
  v.e_hi = edge_ptr[g + 1];
  v.population = v.n * (v.n - 1);
  return v;
}

// marks the edge codes; returns (in every thread) the number of off-diagonal edges of the graph
__device__ int64_t mark_codes(const int64_t* __restrict__ ei, int64_t E, const GraphView& v, uint32_t* bits, int words,
                              int* sm_count) {
  for (int w = threadIdx.x; w < words; w += blockDim.x) bits[w] = 0u;
  if (threadIdx.x == 0) *sm_count = 0;
  __syncthreads();
  int mine = 0;
  for (int64_t e = v.e_lo + threadIdx.x; e < v.e_hi; e += blockDim.x) {
    const int64_t r = ei[e] - v.node_lo, c = ei[E + e] - v.node_lo;
    if (r != c) {
      ++mine;
      const int64_t code = r * (v.n - 1) + (r < c ? c - 1 : c);
      atomicOr(&bits[code >> 5], 1u << (code & 31));
    }
  }
  atomicAdd(sm_count, mine);
  __syncthreads();
  return (int64_t)*sm_count;
}

__device__ __forceinline__ bool deterministic_branch(int64_t codes, int64_t population, int64_t quota) {
  if (codes >= population) return false;                                  // no negative exists
  const double p_neg = 1.0 - (double)codes / (double)population;
  const int64_t k = (int64_t)(1.1 * (double)quota / p_neg);               // int(): truncation toward zero
  return population <= k;
}

__global__ void __launch_bounds__(kNegThreads)
negsample_count_kernel(const int64_t* __restrict__ ei, int64_t E, const int32_t* __restrict__ node_ptr,
                       const int32_t* __restrict__ edge_ptr, const int64_t* __restrict__ last_graph, int64_t quota,
                       int64_t* __restrict__ counts, int32_t* __restrict__ needs_host) {
  extern __shared__ uint32_t bits[];
  __shared__ int sm_count, sm_free;
  const int g = blockIdx.x;
  if (threadIdx.x == 0) counts[g] = 0;
  if ((int64_t)g > *last_graph) return;                                    // upstream skips the graphs behind the last edge
  const GraphView v = view_of(node_ptr, edge_ptr, g);
  if (v.n < 2) return;
  if (v.n > kNegMaxNodes) {
    if (threadIdx.x == 0) atomicExch(needs_host, 1);
    return;
  }
  const int words = (int)((v.population + 31) >> 5);
  const int64_t codes = mark_codes(ei, E, v, bits, words, &sm_count);
  if (codes >= v.population) return;                                       // empty result for this graph
  if (!deterministic_branch(codes, v.population, quota)) {
    if (threadIdx.x == 0) atomicExch(needs_host, 1);
    return;
  }
  if (threadIdx.x == 0) sm_free = 0;
  __syncthreads();
  int mine = 0;
  for (int w = threadIdx.x; w < words; w += blockDim.x) {
    const int valid = (int)min((int64_t)32, v.population - ((int64_t)w << 5));
    const uint32_t mask = valid == 32 ? 0xFFFFFFFFu : ((1u << valid) - 1u);
    mine += __popc(~bits[w] & mask);
  }
  atomicAdd(&sm_free, mine);
  __syncthreads();
  if (threadIdx.x == 0) counts[g] = min((int64_t)sm_free, quota);
}

__global__ void __launch_bounds__(kNegThreads)
negsample_write_kernel(const int64_t* __restrict__ ei, int64_t E, const int32_t* __restrict__ node_ptr,
                       const int32_t* __restrict__ edge_ptr, const int64_t* __restrict__ counts,
                       const int64_t* __restrict__ offsets, int64_t total, int64_t* __restrict__ out) {
  extern __shared__ uint32_t bits[];
  __shared__ int sm_count;
  __shared__ int chunk_base[kNegThreads + 1];
  const int g = blockIdx.x;
  const int64_t want = counts[g];
  if (want == 0) return;
  const GraphView v = view_of(node_ptr, edge_ptr, g);
  const int words = (int)((v.population + 31) >> 5);
  mark_codes(ei, E, v, bits, words, &sm_count);
  // every thread owns a contiguous run of words: count its free codes, scan over the threads, then emit in order
  const int per = (words + blockDim.x - 1) / blockDim.x;
  const int w0 = min(words, (int)threadIdx.x * per), w1 = min(words, w0 + per);
  int mine = 0;
  for (int w = w0; w < w1; ++w) {
    const int valid = (int)min((int64_t)32, v.population - ((int64_t)w << 5));
    const uint32_t mask = valid == 32 ? 0xFFFFFFFFu : ((1u << valid) - 1u);
    mine += __popc(~bits[w] & mask);
  }
  chunk_base[threadIdx.x + 1] = mine;
  if (threadIdx.x == 0) chunk_base[0] = 0;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int t = 1; t <= (int)blockDim.x; ++t) chunk_base[t] += chunk_base[t - 1];
  __syncthreads();
  int64_t rank = chunk_base[threadIdx.x];
  const int64_t base = offsets[g];
  for (int w = w0; w < w1 && rank < want; ++w) {
    const int valid = (int)min((int64_t)32, v.population - ((int64_t)w << 5));
    const uint32_t mask = valid == 32 ? 0xFFFFFFFFu : ((1u << valid) - 1u);
    uint32_t free_bits = ~bits[w] & mask;
    while (free_bits && rank < want) {
      const int b = __ffs(free_bits) - 1;
      free_bits &= free_bits - 1;
      const int64_t code = ((int64_t)w << 5) + b;
      const int64_t r = code / (v.n - 1);
      int64_t c = code - r * (v.n - 1);
      if (r <= c) ++c;
      out[base + rank] = r + v.node_lo;
      out[total + base + rank] = c + v.node_lo;
      ++rank;
    }
  }
}

}  // namespace
}  // namespace gnnb200

using namespace gnnb200;

static size_t negsample_smem(void) { return (size_t)kNegMaxWords * 4; }

extern "C" int gnnb200_negsample_count_i64(const int64_t* edge_index, int64_t num_edges, const int32_t* node_ptr,
                                           const int32_t* edge_ptr, int64_t num_graphs, const int64_t* last_graph,
                                           int64_t quota, int64_t* counts, int32_t* needs_host,
                                           gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_edges < 0 || num_graphs < 0 || quota < 0) return GNNB200_EINVAL;
  if (num_graphs == 0) return GNNB200_OK;
  if (!node_ptr || !edge_ptr || !last_graph || !counts || !needs_host || (num_edges > 0 && !edge_index)) return GNNB200_EINVAL;
  if (num_graphs >= (int64_t)INT32_MAX || num_edges >= (int64_t)INT32_MAX) return GNNB200_ERANGE;
  GNNB200_CHECK_CUDA(cudaFuncSetAttribute(negsample_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)negsample_smem()));
  negsample_count_kernel<<<(unsigned)num_graphs, kNegThreads, negsample_smem(), stream>>>(edge_index, num_edges, node_ptr, edge_ptr,
                                                                                       last_graph, quota, counts, needs_host);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

extern "C" int gnnb200_negsample_write_i64(const int64_t* edge_index, int64_t num_edges, const int32_t* node_ptr,
                                           const int32_t* edge_ptr, int64_t num_graphs, const int64_t* counts,
                                           const int64_t* offsets, int64_t total, int64_t* out, gnnb200_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (num_edges < 0 || num_graphs < 0 || total < 0) return GNNB200_EINVAL;
  if (num_graphs == 0 || total == 0) return GNNB200_OK;
  if (!node_ptr || !edge_ptr || !counts || !offsets || !out || (num_edges > 0 && !edge_index)) return GNNB200_EINVAL;
  if (num_graphs >= (int64_t)INT32_MAX || num_edges >= (int64_t)INT32_MAX) return GNNB200_ERANGE;
  GNNB200_CHECK_CUDA(cudaFuncSetAttribute(negsample_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)negsample_smem()));
  negsample_write_kernel<<<(unsigned)num_graphs, kNegThreads, negsample_smem(), stream>>>(edge_index, num_edges, node_ptr, edge_ptr,
                                                                                       counts, offsets, total, out);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}
