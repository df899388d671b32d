// TF32 tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32) with TMA-fed 128B-swizzled shared
// memory tiles and fp32 accumulators in TMEM.  Replaces the cuBLAS addmm calls behind nn.Linear
// forward/backward in the reference (src/models/gnn.py:13,31,34; src/models/heads.py:41), K3 of
// SURVEY §2.4.  fp32 operands are consumed in place (kind::tf32 reads 32-bit words from shared
// memory), so no cast pass is needed.
//
//   C[M,N] = op(A) op(B) (+bias)(ReLU),  op(A) is [M,K], op(B) is [K,N]
//   A "K-major"  : stored [M,K] row-major (transa=0)      A "MN-major": stored [K,M] (transa=1)
//   B "K-major"  : stored [N,K] row-major (transb=1)      B "MN-major": stored [K,N] (transb=0)
//
// Persistent kernel, one CTA per SM, 384 threads:
//   warp 0   TMA producer (one elected lane): kStages-deep ring of {A tile 128x32, B tile BNx32} fp32
//   warp 1   MMA issuer  (one elected lane): 4 x tcgen05.mma (M=128, N=BN, K=8) per stage,
//            tcgen05.commit releases the stage; accumulators double-buffered in TMEM (2 x BN columns)
//   warp 2   TMEM allocator (512 columns)
//   warps 4-11 epilogue (two per TMEM lane quarter, each half of the columns): tcgen05.ld 32 lanes x 32
//            columns (next chunk prefetched) -> smem transpose -> bias/residual/ReLU -> coalesced 128-bit
//            global stores, overlapped with the next tile's main loop
// Split-K (used for the weight gradients, K = number of nodes): each (tile, split) writes an fp32
// partial to the workspace, a second kernel sums the splits in order (deterministic) and applies the
// epilogue.  Out-of-range rows/columns/k are zero-filled by TMA and masked in the store.
#include <cuda.h>
#include <stdio.h>
#include "common.cuh"

namespace gnnb200 {

int colstats_finish(const float* part, long long chunks, long long cols, float* sum, float* m2, cudaStream_t stream);

namespace tc {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 32;           // fp32 elements per stage along K = one 128-byte swizzle span
constexpr int UMMA_K = 8;        // kind::tf32
constexpr int kMaxStages = 4;
constexpr int kSmemBudget = 232448 - 1024;   // 227 KB per CTA minus the 1024 B alignment slack
constexpr int kThreads = 384;     // 4 control warps + 8 epilogue warps
constexpr int kEpiWarps = 8;       // plain tf32: warps 4..11 run the epilogue; 3xTF32: warps 4..7 do, warps 2,3,8..11 split operands
constexpr int kTmemCols = 512;
constexpr int kEpiTileFloats = 32 * 32;   // per-warp epilogue transpose tile, XOR-swizzled (no padding)
constexpr uint32_t kSpinLimit = 1u << 27;   // bounded mbarrier spin: trap instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// TMA store of one shared-memory box (bulk async-group completion): the epilogue's TMA_ST variant
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING shared memory (the tile may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (SM100 UMMA): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64): 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B
// (the only legal layout for MN-major 32-bit operands: 128 B rows, 32 B swizzle atoms, 4-row groups).
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// Instruction descriptor: D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2, a_major bit15, b_major bit16,
// N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Params {
  int M, N, K;
  int m_tiles, n_tiles, splits, k_blocks_per_split;
  float* C;          // output, or the split-K partial buffer [splits][M][N]
  long long ldc;
  const float* bias;  // applied only when splits == 1
  const float* residual;  // [M,N] added in the epilogue (GINLayer's `+ h`), only when splits == 1
  long long ldr;
  int relu;
  float* col_part;    // optional [m_tiles*4][3][N] per-32-row (count, sum, centred m2) of the written C columns (BN statistics)
};

// X3 = error-compensated "3xTF32": every operand word v is split into hi = v with the 13 low mantissa bits cleared
// (exactly representable in tf32) and lo = v - hi (exact in fp32), and each K-step issues three MMAs
// hi*hi + hi*lo + lo*hi  into the same fp32 accumulator: the dropped lo*lo term and the tf32 truncation of lo are
// ~2^-21 relative, i.e. fp32-class results from the tf32 tensor pipe.  Where the split happens:
//   X3 = 1  both operands in shared memory by the splitter warps (any layout; used by precision 'tf32x3')
//   X3 = 2  B arrives pre-split from global memory (two tensor maps: hi, lo — a weight matrix, split once per optimizer
//           step); the splitter warps rewrite only the A tile (hi in place + lo twin)
//   X3 = 3  as 2, but the raw fp32 tiles serve as the hi operands (kind::tf32 reads only the upper 19 bits of each word —
//           checked bit for bit by tests/test_gpu_gemm.py::test_tf32_mma_truncates_operands), so the splitter only WRITES
//           lo(A) and the hi*hi / hi*lo(B) MMAs of a stage start as soon as TMA has landed it
template <int BN, int X3>
struct TileCfg {
  static constexpr uint32_t kABytes = BM * BK * 4;   // 16 KB
  static constexpr uint32_t kBBytes = BN * BK * 4;
  static constexpr uint32_t kHiBytes = kABytes + kBBytes;
  static constexpr uint32_t kStageBytes = kHiBytes * (X3 ? 2 : 1);
  static constexpr int kEpiBytes = kEpiWarps * kEpiTileFloats * 4;
  static constexpr int kStagesRaw = (kSmemBudget - kEpiBytes) / (int)kStageBytes;
  static constexpr int kStages = kStagesRaw > kMaxStages ? kMaxStages : kStagesRaw;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + kEpiBytes + 1024;
  static_assert(kStages >= 2, "tile does not fit shared memory");
};

// TMA_ST (plain tf32 without split-K, residual or fused statistics; measured on B200: C5 step 243.3 -> 239.1 ms): the epilogue's
// XOR-swizzled 32x32 staging tile IS the SWIZZLE_128B box layout, so instead of reading it back and issuing
// st.global the warp applies bias/ReLU on the TMEM side, and one lane hands the tile to the TMA unit
// (cp.async.bulk.tensor store, clipped at the matrix edge by the tensor map).
template <int BN, bool A_MN, bool B_MN, int X3, bool TMA_ST = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_blo, const Params p) {
  static_assert(!(TMA_ST && X3 != 0), "the TMA-store epilogue exists for the plain tf32 kernel only");
  static_assert(X3 < 2 || (!A_MN && !B_MN), "pre-split B: nn.Linear forward layout only (A [M,K], B [N,K])");
  using Cfg = TileCfg<BN, X3>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kABytes = Cfg::kABytes;
  constexpr uint32_t kBBytes = Cfg::kBBytes;
  constexpr uint32_t kHiBytes = Cfg::kHiBytes;
  constexpr uint32_t kStageBytes = Cfg::kStageBytes;
  constexpr int kEpi = X3 != 0 ? 4 : kEpiWarps;       // epilogue warps
  constexpr int kSplit = 6;                           // splitter warps (X3 only)
  constexpr uint32_t kSlabBytes = BK * 128;   // MN-major: one 32-wide slab = BK rows x 128 B
  constexpr uint32_t kIdesc = make_idesc(BN, A_MN, B_MN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SW128 needs 1024 B
  float* epi_smem = reinterpret_cast<float*>(smem + kStages * kStageBytes);                       // kEpiWarps swizzled 32x32 tiles
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], split_bar[kMaxStages], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
  const int k_blocks_total = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
    if constexpr (TMA_ST) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_c) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&split_bar[s], kSplit);  // one arrival per splitter warp (X3 only)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpi);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  auto tile_coords = [&](int t, int& m0, int& n0, int& kb0, int& kb1, int& split) {
    const int nt = t % p.n_tiles;
    const int rest = t / p.n_tiles;
    const int mt = rest % p.m_tiles;
    split = rest / p.m_tiles;
    m0 = mt * BM;
    n0 = nt * BN;
    kb0 = split * p.k_blocks_per_split;
    kb1 = min(k_blocks_total, kb0 + p.k_blocks_per_split);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m0, n0, kb0, kb1, split;
        tile_coords(t, m0, n0, kb0, kb1, split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&full_bar[stage], X3 >= 2 ? kHiBytes + kBBytes : kHiBytes);
          const int k0 = kb * BK;
          if constexpr (X3 >= 2) tma_load_2d(&map_blo, &full_bar[stage], sb + kHiBytes, k0, n0);   // lo(B) twin tile
          if (A_MN) {
#pragma unroll
            for (int s = 0; s < BM / 32; ++s) tma_load_2d(&map_a, &full_bar[stage], sa + s * kSlabBytes, m0 + 32 * s, k0);
          } else {
            tma_load_2d(&map_a, &full_bar[stage], sa, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int s = 0; s < BN / 32; ++s) tma_load_2d(&map_b, &full_bar[stage], sb + s * kSlabBytes, n0 + 32 * s, k0);
          } else {
            tma_load_2d(&map_b, &full_bar[stage], sb, k0, n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m0, n0, kb0, kb1, split;
        tile_coords(t, m0, n0, kb0, kb1, split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait((X3 == 1 || X3 == 2) ? &split_bar[stage] : &full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
          if constexpr (X3 == 3) {
            // raw tiles are the hi operands: two of the three products need nothing from the splitter warps
            const uint64_t lo_off = (uint64_t)(kHiBytes >> 4);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da = make_desc(sa + k * 32, 16, 1024, kLayoutSw128);
              const uint64_t db = make_desc(sb + k * 32, 16, 1024, kLayoutSw128);
              umma_tf32(d_tmem, da, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);             // hi(A) * hi(B)
              umma_tf32(d_tmem, da, db + lo_off, kIdesc, 1u);                               // hi(A) * lo(B)
            }
            mbar_wait(&split_bar[stage], phase);
            tcgen05_fence_after();
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da = make_desc(sa + k * 32, 16, 1024, kLayoutSw128);
              const uint64_t db = make_desc(sb + k * 32, 16, 1024, kLayoutSw128);
              umma_tf32(d_tmem, da + lo_off, db, kIdesc, 1u);                               // lo(A) * hi(B)
            }
          } else {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: rows of 128 B, 8-row groups 1024 B apart (SBO); one MMA (K=8 tf32) advances 32 B inside
            // the swizzle span.  MN-major: 32-wide slabs kSlabBytes apart (LBO), 4-row groups 512 B apart
            // (SBO); one MMA consumes 8 k-rows = 1024 B.
            const uint64_t da = A_MN ? make_desc(sa + k * 1024, kSlabBytes, 512, kLayoutSw128Base32)
                                     : make_desc(sa + k * 32, 16, 1024, kLayoutSw128);
            const uint64_t db = B_MN ? make_desc(sb + k * 1024, kSlabBytes, 512, kLayoutSw128Base32)
                                     : make_desc(sb + k * 32, 16, 1024, kLayoutSw128);
            if (X3) {
              // lo tiles live kHiBytes above their hi twins (same layout): descriptor start address += kHiBytes >> 4
              const uint64_t lo_off = (uint64_t)(kHiBytes >> 4);
              umma_tf32(d_tmem, da + lo_off, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);   // lo(A) * hi(B)
              umma_tf32(d_tmem, da, db + lo_off, kIdesc, 1u);                               // hi(A) * lo(B)
              umma_tf32(d_tmem, da, db, kIdesc, 1u);                                        // hi(A) * hi(B)
            } else {
              umma_tf32(d_tmem, da, db, kIdesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          }   // X3 != 3
          umma_commit(&empty_bar[stage]);              // frees the smem stage once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);                  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (X3 != 0 && (warp == 2 || warp == 3 || warp >= 8)) {
    // ===================== operand splitter (X3 only): hi/lo decomposition in shared memory =====================
    {
      const int tid_s = (warp < 4 ? warp - 2 : warp - 6) * 32 + lane;      // 0 .. kSplit*32-1
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m0, n0, kb0, kb1, split;
        tile_coords(t, m0, n0, kb0, kb1, split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);                 // TMA bytes have landed (hi region holds raw fp32)
          uint4* hi = reinterpret_cast<uint4*>(smem + stage * kStageBytes);
          uint4* lo = reinterpret_cast<uint4*>(smem + stage * kStageBytes + kHiBytes);
          constexpr int kSplitVecs = (int)((X3 == 1 ? kHiBytes : kABytes) / 16);   // pre-split B: only the A tile
#pragma unroll 4
          for (int i = tid_s; i < kSplitVecs; i += kSplit * 32) {
            const uint4 v = hi[i];
            uint4 h, l;
            h.x = v.x & 0xFFFFE000u; h.y = v.y & 0xFFFFE000u; h.z = v.z & 0xFFFFE000u; h.w = v.w & 0xFFFFE000u;
            l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
            l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
            l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
            l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
            if constexpr (X3 != 3) hi[i] = h;
            lo[i] = l;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive(&split_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpi) {
    // ===================== epilogue (warps 4..4+kEpi-1) =====================
    // A warp may only touch TMEM lanes 32*(warp%4)..+31; two warps share each lane quarter and split the
    // accumulator's columns, so every SM sub-partition has two epilogue warps to interleave.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int kChunksPerWarp = (BN / 32) / (kEpi / 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int m0, n0, kb0, kb1, split;
      tile_coords(t, m0, n0, kb0, kb1, split);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      float* cbase = p.C + (p.splits > 1 ? (long long)split * p.M * p.N : 0);
      const bool has_k = kb1 > kb0;                    // an empty split contributes zeros
      // TMEM gives lane = row; a direct store would touch 32 rows x 16 B per instruction.  Each warp instead
      // transposes its 32x32 chunk through a private smem tile (pitch 36 floats: conflict-free 128-bit
      // stores by row and 128-bit loads by quarter-warp) and writes 4 full 128-byte row segments per
      // instruction; bias / residual / ReLU are applied on the coalesced side.
      float* stage = epi_smem + (warp - 4) * kEpiTileFloats;
      const int r_in = lane >> 3;                      // row inside a group of 4
      const int c4 = (lane & 7) << 2;                  // column (floats) inside the 32-wide chunk
      const int c_begin = half * kChunksPerWarp, c_end = c_begin + kChunksPerWarp;
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c_begin * 32), v);
      if constexpr (TMA_ST) {
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          if (lane == 0) tma_store_wait_read();          // the previous chunk's store has drained the staging tile
          __syncwarp();
          tmem_ld_wait();
          const int col0 = n0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
            if (p.bias && col0 + j < p.N) {               // warp-uniform address: one broadcast load
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            }
            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(stage + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) = o;
          }
          if (c + 1 < c_end) {                           // next chunk's TMEM load overlaps this chunk's store
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + (c + 1) * 32), v);
          } else {                                       // accumulator fully read: hand TMEM back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA unit
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_c, stage, col0, m0 + q * 32);                // rows >= M / columns >= N are clipped
            tma_store_commit();
          }
        }
      } else {   // st.global epilogue (body below keeps its indentation)
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        const int col = n0 + c * 32 + c4;
        // residual rows for this chunk are requested first so their latency hides behind the TMEM load and the
        // transpose (8 independent 128-bit loads per lane, coalesced: 4 full 128-byte segments per instruction)
        float4 rr[8];
        if (p.splits == 1 && p.residual) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = m0 + q * 32 + it * 4 + r_in;
            rr[it] = (row < p.M && col < p.N) ? __ldg(reinterpret_cast<const float4*>(p.residual + (long long)row * p.ldr + col))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        tmem_ld_wait();
        // staging tile: element (row r, 16-byte slot s) at float offset r*32 + ((s ^ (r & 7)) << 2): both the row-wise
        // 128-bit stores and the quarter-warp row reads below are bank-conflict free without padding
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        if (c + 1 < c_end) {                           // next chunk's TMEM load overlaps this chunk's stores
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + (c + 1) * 32), v);
        } else {                                       // accumulator fully read: hand TMEM back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        __syncwarp();
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.splits == 1 && p.bias && col < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + r_in;
          const int row = m0 + q * 32 + r;
          float4 o = *reinterpret_cast<const float4*>(stage + r * 32 + ((((lane & 7)) ^ (r & 7)) << 2));
          if (row < p.M && col < p.N) {                 // N % 4 == 0 is guaranteed by the host
            if (!has_k) o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.splits == 1) {
              o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
              if (p.residual) { o.x += rr[it].x; o.y += rr[it].y; o.z += rr[it].z; o.w += rr[it].w; }
              if (p.relu) {
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
              }
            }
            *reinterpret_cast<float4*>(cbase + (long long)row * p.ldc + col) = o;
          } else {
            o = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (p.col_part) *reinterpret_cast<float4*>(stage + r * 32 + ((((lane & 7)) ^ (r & 7)) << 2)) = o;   // final values, own slot
        }
        if (p.col_part) {
          // BatchNorm statistics of the columns just written (saves the separate read pass over C): per 32-row block the
          // sum and the second moment about the block's own mean, merged later with Chan's formula (fixed order).
          const int row0 = m0 + q * 32;
          const int n_valid = max(0, min(32, p.M - row0));
          float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + r_in;
            const float4 o = *reinterpret_cast<const float4*>(stage + r * 32 + ((((lane & 7)) ^ (r & 7)) << 2));
            s4.x += o.x; s4.y += o.y; s4.z += o.z; s4.w += o.w;
          }
#pragma unroll
          for (int sh = 8; sh <= 16; sh <<= 1) {
            s4.x += __shfl_xor_sync(0xffffffffu, s4.x, sh); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, sh);
            s4.z += __shfl_xor_sync(0xffffffffu, s4.z, sh); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, sh);
          }
          const float inv_n = n_valid > 0 ? 1.f / (float)n_valid : 0.f;
          const float4 mu = make_float4(s4.x * inv_n, s4.y * inv_n, s4.z * inv_n, s4.w * inv_n);
          float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + r_in;
            if (row0 + r < p.M) {
              const float4 o = *reinterpret_cast<const float4*>(stage + r * 32 + ((((lane & 7)) ^ (r & 7)) << 2));
              const float dx = o.x - mu.x, dy = o.y - mu.y, dz = o.z - mu.z, dw = o.w - mu.w;
              q4.x = fmaf(dx, dx, q4.x); q4.y = fmaf(dy, dy, q4.y); q4.z = fmaf(dz, dz, q4.z); q4.w = fmaf(dw, dw, q4.w);
            }
          }
#pragma unroll
          for (int sh = 8; sh <= 16; sh <<= 1) {
            q4.x += __shfl_xor_sync(0xffffffffu, q4.x, sh); q4.y += __shfl_xor_sync(0xffffffffu, q4.y, sh);
            q4.z += __shfl_xor_sync(0xffffffffu, q4.z, sh); q4.w += __shfl_xor_sync(0xffffffffu, q4.w, sh);
          }
          if (r_in == 0 && col < p.N && row0 < p.M) {
            float* part = p.col_part + (long long)(row0 >> 5) * 3 * p.N;
            const float nv = (float)n_valid;
            *reinterpret_cast<float4*>(part + col) = make_float4(nv, nv, nv, nv);
            *reinterpret_cast<float4*>(part + p.N + col) = s4;
            *reinterpret_cast<float4*>(part + 2 * p.N + col) = q4;
          }
        }
        __syncwarp();                                   // the staging tile is reused by the next chunk
      }
      }   // !TMA_ST
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (TMA_ST) {
      if (lane == 0) tma_store_wait_read();              // shared memory must outlive the last store's reads
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long mn, int N, const float* __restrict__ bias,
                     const float* __restrict__ residual, long long ldr, int relu, float* __restrict__ C, long long ldc) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= mn) return;
  float4 acc = *reinterpret_cast<const float4*>(partial + i4);
  for (int s = 1; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(partial + (long long)s * mn + i4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const long long m = i4 / N;
  const int n = (int)(i4 % N);
  if (bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  if (residual) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(residual + m * ldr + n));
    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
  }
  if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
  *reinterpret_cast<float4*>(C + m * ldc + n) = acc;
}

// hi = v with the 13 low mantissa bits cleared, lo = v - hi (both exact): the pre-split form of a weight matrix
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, long long n, float* __restrict__ hi, float* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  hi[i] = h;
  lo[i] = v - h;
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;   // benign race: every thread resolves the same pointer
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D fp32 tensor map: inner (contiguous) extent `inner`, outer extent `outer`, row pitch ld elements,
// box {32 floats = 128 B, box_outer rows}, zero fill out of bounds; 128-byte swizzle with 16 B atoms for
// K-major tiles and with 32 B atoms for MN-major tiles (matches the UMMA descriptor layout types).
static int make_map(CUtensorMap* map, const float* base, long long inner, long long outer, long long ld, int box_outer,
                    bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return GNNB200_ETMA;
  // The driver entry point needs a current context on THIS host thread.  Threads that have only been
  // handed a device ordinal (e.g. PyTorch's autograd worker) may not have the primary context bound
  // yet; a no-op runtime call binds it.
  static thread_local bool context_bound = false;
  if (!context_bound) {
    cudaFree(nullptr);
    context_bound = true;
  }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "gnnb200: cuTensorMapEncodeTiled failed (CUresult %d): base=%p inner=%lld outer=%lld ld=%lld box_outer=%d mn=%d\n",
            (int)r, (const void*)base, inner, outer, ld, box_outer, (int)mn_major);
    return GNNB200_ETMA;
  }
  return GNNB200_OK;
}

static int pick_bn(long long N, int x3) {
  // x3 = 1 (both operands split in shared memory): 128-wide tiles (a 256-wide stage would leave one pipeline stage);
  // x3 >= 2 (B pre-split): 256-wide tiles with a two-stage ring — every A tile is split once, not once per 128 columns
  const bool wide = x3 != 1;
  if (wide && N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  if (wide && N >= 256) return 256;   // ragged last tile: TMA zero-fills, the store masks (N % 4 == 0 required)
  if (N >= 128) return 128;
  return 64;
}

template <int BN, bool A_MN, bool B_MN, int X3, bool TMA_ST>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mblo,
                  const Params& p, int grid, cudaStream_t stream) {
  constexpr size_t smem = TileCfg<BN, X3>::kSmemBytes;
  // the attribute is per device (a process-wide flag would leave a second GPU of the same process unconfigured);
  // racing threads at worst set it twice
  static bool configured[64] = {};
  int dev = 0;
  GNNB200_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    GNNB200_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BN, A_MN, B_MN, X3, TMA_ST>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  gemm_tf32_kernel<BN, A_MN, B_MN, X3, TMA_ST><<<grid, kThreads, smem, stream>>>(ma, mb, mc, mblo, p);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

template <int BN, int X3, bool TMA_ST = false>
static int launch_bn(bool a_mn, bool b_mn, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc,
                     const CUtensorMap& mblo, const Params& p, int grid, cudaStream_t stream) {
  if constexpr (X3 >= 2) {
    return launch<BN, false, false, X3, TMA_ST>(ma, mb, mc, mblo, p, grid, stream);   // checked by the caller
  } else {
    if (!a_mn && !b_mn) return launch<BN, false, false, X3, TMA_ST>(ma, mb, mc, mblo, p, grid, stream);
    if (!a_mn && b_mn) return launch<BN, false, true, X3, TMA_ST>(ma, mb, mc, mblo, p, grid, stream);
    if (a_mn && !b_mn) return launch<BN, true, false, X3, TMA_ST>(ma, mb, mc, mblo, p, grid, stream);
    return launch<BN, true, true, X3, TMA_ST>(ma, mb, mc, mblo, p, grid, stream);
  }
}

}  // namespace tc

int split_tf32(const float* x, int64_t n, float* hi, float* lo, cudaStream_t stream) {
  if (n <= 0) return GNNB200_OK;
  tc::split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, n, hi, lo);
  GNNB200_LAUNCH_CHECK();
  return GNNB200_OK;
}

int gemm_tf32_supported(const float* A, int64_t lda, int transa, const float* B, int64_t ldb, int transb,
                        const float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const float* residual,
                        int64_t ldr) {
  if (residual && (ldr % 4 != 0 || ((uintptr_t)residual & 15))) return 0;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (N % 4 != 0 || N < 8) return 0;
  if (lda % 4 != 0 || ldb % 4 != 0 || ldc % 4 != 0) return 0;
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return 0;
  if (M >= (1LL << 31) || K >= (1LL << 31)) return 0;
  (void)transa; (void)transb;
  return tc::encode_fn() != nullptr;
}

// x3: 0 plain tf32, 1 3xTF32 with both operands split in shared memory, 2 / 3 3xTF32 with B pre-split (B = hi, B_lo = lo;
// layout transa = 0, transb = 1 only; 3 = B holds the raw weights and the raw tiles act as hi operands)
int gemm_tf32(const float* A, int64_t lda, int transa, const float* B, const float* B_lo, int64_t ldb, int transb, float* C,
              int64_t ldc, int64_t M, int64_t N, int64_t K, const float* bias, const float* residual, int64_t ldr,
              int epilogue, int x3, float* col_sum, float* col_m2, void* workspace, size_t* workspace_bytes,
              cudaStream_t stream) {
  using namespace tc;
  if (x3 >= 2 && (transa != 0 || transb == 0 || (workspace && !B_lo))) return GNNB200_EINVAL;
  const int bn = pick_bn(N, x3);
  const int m_tiles = (int)((M + BM - 1) / BM);
  const int n_tiles = (int)((N + bn - 1) / bn);
  const int k_blocks = (int)((K + BK - 1) / BK);
  // split-K only when the output has too few tiles to fill the machine and K is long
  int splits = 1;
  const long long tiles = (long long)m_tiles * n_tiles;
  if (tiles * 2 <= kNumSMs && k_blocks >= 64) {
    splits = (int)(kNumSMs / tiles);
    const int max_s = k_blocks / 16;
    if (splits > max_s) splits = max_s;
    if (splits < 1) splits = 1;
  }
  const bool want_stats = col_sum != nullptr || col_m2 != nullptr;
  if (want_stats && (M + 31) / 32 > 2147483647LL / 4) return GNNB200_ERANGE;
  if (want_stats) splits = 1;      // statistics come from the fused epilogue
  const int kbps = (k_blocks + splits - 1) / splits;
  splits = (k_blocks + kbps - 1) / kbps;
  const long long stat_blocks = (M + 31) / 32;
  Workspace ws(workspace);
  float* partial = splits > 1 ? ws.take<float>((size_t)splits * M * N) : nullptr;
  float* col_part = want_stats ? ws.take<float>((size_t)stat_blocks * 3 * N) : nullptr;
  if (!workspace) {
    *workspace_bytes = ws.bytes();
    return GNNB200_OK;
  }
  if (*workspace_bytes < ws.bytes()) return GNNB200_EWORKSPACE;

  const bool a_mn = transa != 0;   // stored [K,M]
  const bool b_mn = transb == 0;   // stored [K,N]
  CUtensorMap ma, mb;
  int rc;
  if (a_mn) rc = make_map(&ma, A, M, K, lda, BK, true);            // inner = M, outer = K, box {32, BK}
  else rc = make_map(&ma, A, K, M, lda, BM, false);                 // inner = K, outer = M, box {32, 128}
  if (rc) return rc;
  if (b_mn) rc = make_map(&mb, B, N, K, ldb, BK, true);
  else rc = make_map(&mb, B, K, N, ldb, bn, false);
  if (rc) return rc;

  Params p;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.splits = splits; p.k_blocks_per_split = kbps;
  p.C = splits > 1 ? partial : C;
  p.ldc = splits > 1 ? N : ldc;
  p.bias = splits > 1 ? nullptr : bias;
  p.residual = splits > 1 ? nullptr : residual;
  p.ldr = ldr;
  p.relu = (splits == 1 && (epilogue & GNNB200_EPI_RELU)) ? 1 : 0;
  p.col_part = col_part;
  const long long total = tiles * splits;
  const int grid = (int)(total < kNumSMs ? total : kNumSMs);
  // TMA-store epilogue: C as a tensor map of 32x32 boxes in the staging tile's swizzle (inner = N, outer = M)
  const bool tma_st = !x3 && splits == 1 && !p.residual && !p.col_part;
  CUtensorMap mc = ma;                                   // unused unless tma_st
  if (tma_st) {
    rc = make_map(&mc, C, N, M, ldc, 32, false);
    if (rc) return rc;
  }
  CUtensorMap mblo = mb;                                 // unused unless B arrives pre-split
  if (x3 >= 2) {
    rc = make_map(&mblo, B_lo, K, N, ldb, bn, false);
    if (rc) return rc;
  }
  if (x3 == 1) {
    if (bn == 128) rc = launch_bn<128, 1>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
    else rc = launch_bn<64, 1>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
  } else if (x3 == 2) {
    if (bn == 256) rc = launch_bn<256, 2>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
    else if (bn == 128) rc = launch_bn<128, 2>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
    else rc = launch_bn<64, 2>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
  } else if (x3 == 3) {
    if (bn == 256) rc = launch_bn<256, 3>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
    else if (bn == 128) rc = launch_bn<128, 3>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
    else rc = launch_bn<64, 3>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
  } else if (tma_st) {
    if (bn == 256) rc = launch_bn<256, 0, true>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
    else if (bn == 128) rc = launch_bn<128, 0, true>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
    else rc = launch_bn<64, 0, true>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
  } else {
    if (bn == 256) rc = launch_bn<256, 0>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
    else if (bn == 128) rc = launch_bn<128, 0>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
    else rc = launch_bn<64, 0>(a_mn, b_mn, ma, mb, mc, mblo, p, grid, stream);
  }
  if (rc) return rc;
  if (splits > 1) {
    const long long mn = (long long)M * N;
    splitk_reduce_kernel<<<(unsigned)((mn / 4 + 255) / 256), 256, 0, stream>>>(partial, splits, mn, (int)N, bias, residual, ldr,
                                                                               (epilogue & GNNB200_EPI_RELU) ? 1 : 0, C, ldc);
    GNNB200_LAUNCH_CHECK();
  }
  if (want_stats) return colstats_finish(col_part, stat_blocks, N, col_sum, col_m2, stream);
  return GNNB200_OK;
}

}  // namespace gnnb200
