// tc_gemm.cuh -- the tcgen05 / TMA / TMEM GEMM of the CLIP ViT encoder (sm_100a).
//
//   C[M,N] = A[M,K] . B[N,K]^T      A, B bf16 row-major (K contiguous = "K-major"), fp32 accumulate in TMEM.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0 (1 lane)  TMA producer : cp.async.bulk.tensor.2d -> 128B-swizzled smem stages, mbarrier complete_tx
//   warp 1 (1 lane)  MMA issuer   : tcgen05.mma.cta_group::1.kind::f16, M=128 x N=BLOCK_N x K=16 per instruction,
//                                   tcgen05.commit releases smem stages / publishes the accumulator
//   warp 2           TMEM allocator (2 accumulator stages x BLOCK_N columns)
//   warps 4-7        epilogue     : tcgen05.ld 32x32b -> registers -> fused epilogue -> global
// Fused epilogues (struct GemmEpi): +bias, QuickGELU (saving the pre-activation), x gelu'(h) for the MLP
// backward, +fp32 residual, fp32 and/or bf16 outputs, and an NCHW "un-patchify" store for the patch-embed
// data gradient. Rows >= M are zero-filled by TMA on load and masked on store.
#pragma once
#include "aph_common.cuh"
#include <cuda.h>

namespace aph {

typedef __nv_bfloat16 bf16;

struct GemmEpi {
  const float* bias = nullptr;     // [N]
  const float* resid = nullptr;    // fp32 [M, N], added last
  const bf16* gelu_in = nullptr;   // bf16 [M, N]: acc *= quickgelu'(gelu_in)
  float* out_f32 = nullptr;        // fp32 [M, N] (or NCHW images when unpatch_p > 0)
  bf16* out_bf16 = nullptr;        // bf16 [M, N]
  bf16* out_pre = nullptr;         // bf16 [M, N] pre-activation (acc + bias), saved for backward
  int act = 0;                     // 1 = QuickGELU x*sigmoid(1.702x)
  int unpatch_p = 0;               // >0: out_f32 is [S,3,R,R]; row = s*g*g + gy*g + gx, col = c*p*p + py*p + px
  int unpatch_g = 0;
  int nostore = 0;                 // profiling experiment: epilogue does everything but the global stores
  // stream-K (set by launch_gemm): CTA groups own contiguous ranges of (tile, k-block) units; a tile split between two
  // groups is finished by the group holding its first k-blocks, the other parks its fp32 partial in sk_ws
  int sk = 0;
  float* sk_ws = nullptr;          // [groups * CG][128][BN] fp32 partial accumulators
  int* sk_flags = nullptr;         // [groups * CG][8] per-epilogue-warp ready flags (self-resetting)
};

// Work list of one CTA group: identical in the producer, MMA and epilogue roles.
struct WorkIter {
  int sk, G, num_tiles, kblocks, u, u1, tile_rr;
  __device__ __forceinline__ void init(int sk_, int g, int G_, int num_tiles_, int kblocks_) {
    sk = sk_; G = G_; num_tiles = num_tiles_; kblocks = kblocks_; tile_rr = g;
    const long long U = (long long)num_tiles_ * kblocks_;
    u = (int)((long long)g * U / G_); u1 = (int)((long long)(g + 1) * U / G_);
  }
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
    if (sk) {
      if (u >= u1) return false;
      tile = u / kblocks; kb0 = u - tile * kblocks; kb1 = min(kblocks, kb0 + (u1 - u)); u += kb1 - kb0;
      return true;
    }
    if (tile_rr >= num_tiles) return false;
    tile = tile_rr; kb0 = 0; kb1 = kblocks; tile_rr += G;
    return true;
  }
};

struct GemmShape { int M, N, K; };

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;       // 64 bf16 = 128 B = one swizzle-128B row
constexpr int GEMM_UK = 16;       // K per tcgen05.mma (kind::f16)
constexpr int GEMM_EPI_WARPS = 8;      // 2 per SM sub-partition (16 measured identical: the epilogue is bound by the output stores, not by warp latency)
constexpr int GEMM_THREADS = (4 + GEMM_EPI_WARPS) * 32;   // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4..: epilogue

// compile-time epilogue kinds (a runtime-flag epilogue was instruction-latency bound: 1 warp per scheduler, long predicated body)
enum : int {
  EPI_F32 = 0,            // out_f32 = acc
  EPI_BF16 = 1,           // out_bf16 = acc
  EPI_BIAS_BF16 = 2,      // out_bf16 = acc + bias
  EPI_BIAS_GELU = 3,      // out_pre = bf16(acc + bias); out_bf16 = quickgelu(acc + bias)
  EPI_BIAS_RESID = 4,     // out_f32 = acc + bias + resid
  EPI_GELUGRAD_BF16 = 5,  // out_bf16 = acc * quickgelu'(gelu_in)
  EPI_UNPATCH = 6,        // out_f32[NCHW] = acc (patch-embed data gradient)
  EPI_KINDS = 7
};

// ---- raw PTX wrappers -------------------------------------------------------------------------
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile load: coordinates (x = element index along K, y = row)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

// 3-D tile load (x = column, y = token, z = sample); out-of-bounds tokens are zero-filled
__device__ __forceinline__ void tma_load_3d(uint32_t dst_saddr, const CUtensorMap* m, uint64_t* bar, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// same with an L2 eviction-priority hint (weights are re-read by every M tile: evict_last; streamed outputs must not evict them)
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(x), "r"(y), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane = accumulator row)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) helpers: two CTAs of a cluster share one 256-row accumulator tile; each loads its own 128 rows
// of A and HALF of the B tile, the leader issues the MMAs for both, tcgen05.commit multicasts the barrier arrivals.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier that may live in the peer (leader) CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_cluster_addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (after all previously issued MMAs retire) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO). sm_100 descriptor version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                          // layout type: SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M=128, N=n
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n, int m = GEMM_BM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// sigmoid(1.702 x) through ONE special-function op: 0.5 tanh(0.851 x) + 0.5 (tanh.approx: max rel. error 2^-11, far inside
// the bf16 rounding of the value it feeds); the epilogues of the MLP GEMMs are MUFU-limited otherwise (exp + rcp per element)
__device__ __forceinline__ float sigmoid1702(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float quickgelu(float x) { return x * sigmoid1702(x); }
__device__ __forceinline__ float quickgelu_grad(float x) {
  const float s = sigmoid1702(x);
  return s * (1.f + 1.702f * x * (1.f - s));
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d));
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

template <int BN, int STAGES, int CG = 1>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / CG) * GEMM_BK * 2;      // a CTA pair splits the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;          // 4 epilogue warps x (32 rows x 128 B) transpose staging
  static constexpr int EPI_BYTES = GEMM_EPI_WARPS * 32 * 128;
  static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;   // + alignment slack
};

template <int BN, int STAGES, int EPI, int CG = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_bf16_tn(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmShape shp, GemmEpi epi) {
  using L = GemmSmem<BN, STAGES, CG>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;      // position inside the CTA pair (0 = leader)
  // a "tile" is one accumulator of (128*CG) x BN: with CG = 2 the pair shares it, each CTA owning 128 of its rows
  const int m_tiles = (shp.M + GEMM_BM * CG - 1) / (GEMM_BM * CG), n_tiles = shp.N / BN;
  const int num_tiles = m_tiles * n_tiles, k_blocks = shp.K / GEMM_BK;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], CG); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], CG * GEMM_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) { if (CG == 2) tmem_alloc_cg2(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // everything above overlapped the previous kernel's tail; operands are read from here on

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      const uint64_t pol_b = l2_policy_evict_last();      // B = weights: shared by all M tiles
      uint32_t it = 0;
      WorkIter wi; wi.init(epi.sk, tile0, tile_step, num_tiles, k_blocks);
      int tile, kb0, kb1;
      while (wi.next(tile, kb0, kb1)) {
        const int m_blk = (tile / n_tiles) * CG + (int)rank, n_blk = tile % n_tiles;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          if (CG == 2) {
            // both CTAs credit the LEADER's full barrier: the leader's MMA consumes the stage of both
            const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[s]), 0);
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES); else mbar_arrive_cluster(lead_full);
            tma_load_2d_2sm(sa, &map_a, lead_full, kb * GEMM_BK, m_blk * GEMM_BM);
            tma_load_2d_2sm(sa + L::A_BYTES, &map_b, lead_full, kb * GEMM_BK, n_blk * BN + (int)rank * (BN / 2));
          } else {
            mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
            tma_load_2d(sa, &map_a, &full_bar[s], kb * GEMM_BK, m_blk * GEMM_BM);
            tma_load_2d_hint(sa + L::A_BYTES, &map_b, &full_bar[s], kb * GEMM_BK, n_blk * BN, pol_b);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (the leader CTA issues for the pair when CG = 2)
      constexpr uint32_t idesc = make_idesc_bf16(BN, GEMM_BM * CG);
      uint32_t it = 0, tile_iter = 0;
      WorkIter wi; wi.init(epi.sk, tile0, tile_step, num_tiles, k_blocks);
      int tile, kb0, kb1;
      for (; wi.next(tile, kb0, kb1); ++tile_iter) {
        const uint32_t as = tile_iter & 1, aph_ = (tile_iter >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph_ ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sa + L::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UK; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (>>4) address field
            if (CG == 2) umma_f16_cg2(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, ((kb - kb0) | k) != 0);
            else umma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, ((kb - kb0) | k) != 0);
          }
          if (CG == 2) {
            umma_commit_mc(&empty_bar[s], 0b11);                          // both CTAs' smem stages are free
            if (kb == kb1 - 1) umma_commit_mc(&tfull_bar[as], 0b11);   // both CTAs' epilogues may read TMEM
          } else {
            umma_commit(&empty_bar[s]);                   // smem stage free once these MMAs retire
            if (kb == kb1 - 1) umma_commit(&tfull_bar[as]);   // accumulator (or this group's partial) complete
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: 8 warps. Warp w reads TMEM lanes 32*(w%4) .. +31 (one accumulator row per lane) and owns the 32-column
    // chunks c = half, half+2, ... (half = (w-4)/4). Each 32x32 fp32 chunk is transposed through a swizzled shared-memory tile
    // so that global traffic is row-contiguous (8 lanes x 16 B = one 128 B line per row); the epilogue kind is compile-time.
    const int q = warp & 3, half = (warp - 4) >> 2;      // half = column group 0..3 of this warp
    const uint32_t stage = smem_u32(smem + L::EPI_OFFSET + (warp - 4) * (32 * 128));
    const int c4 = lane & 7, rsub = lane >> 3;
    // Two loop forms. The gelu-grad kind software-pipelines its global operand (the saved pre-activation) one chunk ahead in a
    // second 16-register buffer; the other kinds keep the single-chunk loop (pipelining the 32-register residual operand, or
    // merely restructuring the operand-free kinds, measured slower in r1p).
    if constexpr (EPI == EPI_GELUGRAD_BF16) {
      uint32_t tile_iter = 0;
      WorkIter wi; wi.init(epi.sk, tile0, tile_step, num_tiles, k_blocks);
      // Global operands of the fused epilogue (the fp32 residual / the saved pre-activation) are software-pipelined one 32-column
      // chunk ahead, across tile boundaries as well: fetched right before use they cost a full HBM round trip per chunk (ncu r1l:
      // 28 % of the gelu-grad kernel's stall samples sat on the first use of these registers).
      constexpr int CSTEP = GEMM_EPI_WARPS / 4, NCH = (BN / 32) / CSTEP;          // chunks per warp per tile (even)
      constexpr int CU = 2;
      static_assert(NCH % 2 == 0, "the operand double buffer assumes an even chunk count");
      float4 res4[2][EPI == EPI_BIAS_RESID ? 8 : 1];
      uint2 gin[2][EPI == EPI_GELUGRAD_BF16 ? 8 : 1];
  #define APH_FETCH_OPS(TILE_, C_, BUF_)                                                                                          \
      do {                                                                                                                        \
        const int fm_ = ((TILE_) / n_tiles * CG + (int)rank) * GEMM_BM + q * 32;                                                  \
        const size_t fo_ = (size_t)(fm_ + rsub) * shp.N + ((TILE_) % n_tiles) * BN + (C_) * 32 + 4 * c4;                          \
        _Pragma("unroll")                                                                                                         \
        for (int it = 0; it < 8; ++it) {                                                                                          \
          const bool ok = fm_ + it * 4 + rsub < shp.M;                                                                            \
          if (EPI == EPI_BIAS_RESID) res4[BUF_][it] = ok ? __ldg(reinterpret_cast<const float4*>(epi.resid + fo_ + (size_t)it * 4 * shp.N)) : make_float4(0.f, 0.f, 0.f, 0.f); \
          if (EPI == EPI_GELUGRAD_BF16) gin[BUF_][it] = ok ? __ldg(reinterpret_cast<const uint2*>(epi.gelu_in + fo_ + (size_t)it * 4 * shp.N)) : make_uint2(0u, 0u); \
        }                                                                                                                         \
      } while (0)
      int tile, kb0, kb1;
      bool have = wi.next(tile, kb0, kb1);
      if (have && !(epi.sk && kb0 > 0)) APH_FETCH_OPS(tile, half, 0);
      while (have) {
        int ntile = 0, nkb0 = 0, nkb1 = 0;
        const bool have_n = wi.next(ntile, nkb0, nkb1);
        const int m_blk = (tile / n_tiles) * CG + (int)rank, n_blk = tile % n_tiles;
        const uint32_t as = tile_iter & 1, aph_ = (tile_iter >> 1) & 1;
        // stream-K roles of this item: park the partial (tile continues from another group's k-blocks) or fix it up
        const bool p_store = epi.sk && kb0 > 0, p_fix = epi.sk && kb1 < k_blocks;
        float* ws_st = epi.sk_ws + ((size_t)(tile0 * CG + (int)rank) * GEMM_BM) * BN;             // slot of this group
        const float* ws_fx = epi.sk_ws + ((size_t)((tile0 + 1) * CG + (int)rank) * GEMM_BM) * BN;  // slot of the next group
        if (p_fix) {
          if (lane == 0) { volatile int* f = epi.sk_flags + ((tile0 + 1) * CG + (int)rank) * GEMM_EPI_WARPS + (warp - 4); const long long t0 = clock64(); while (*f == 0 && clock64() - t0 < 4000000000LL) { } }   /* bounded (~2 s): a protocol bug must fail a test, not hang the GPU */
          __threadfence();
          __syncwarp();
        }
        mbar_wait(&tfull_bar[as], aph_);
        tc_fence_after();
        const int m_base = m_blk * GEMM_BM + q * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
  #pragma unroll 1
        for (int ci0 = 0; ci0 < NCH; ci0 += CU) {
  #pragma unroll
        for (int cu = 0; cu < CU; ++cu) {                  // pairs of chunks: the operand buffer index is a compile-time constant
          const int ci = ci0 + cu;
          const int c = half + ci * CSTEP;
          // next chunk's (or the next tile's first chunk's) operands
          if (ci + 1 < NCH) { if (!p_store) APH_FETCH_OPS(tile, c + CSTEP, (cu + 1) & 1); }
          else if (have_n && !(epi.sk && nkb0 > 0)) APH_FETCH_OPS(ntile, half, (cu + 1) & 1);
          const int col = n_blk * BN + c * 32 + 4 * c4;
          const size_t off0 = (size_t)(m_base + rsub) * shp.N + col;
          const size_t step = (size_t)4 * shp.N;
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_RESID) bias4 = __ldg(reinterpret_cast<const float4*>(epi.bias + col));
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_wait_ld();
          if (ci == NCH - 1) {                             // this warp's last read of the accumulator: hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0)); else mbar_arrive(&tempty_bar[as]); }
          }
  #pragma unroll
          for (int i = 0; i < 8; ++i)      // lane = row: write its 32 columns as 8 swizzled 16-byte chunks
            sts128(stage + lane * 128 + ((i ^ (lane & 7)) << 4), r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          __syncwarp();
          size_t off = off0;
  #pragma unroll
          for (int it = 0; it < 8; ++it, off += step) {
            const int rr = it * 4 + rsub;
            const int m = m_base + rr;
            if (m >= shp.M) break;
            float4 v = lds128(stage + rr * 128 + ((c4 ^ (rr & 7)) << 4));
            if (p_store || p_fix) {
              const size_t woff = (size_t)(q * 32 + rr) * BN + c * 32 + 4 * c4;
              if (p_store) { *reinterpret_cast<float4*>(ws_st + woff) = v; continue; }
              const float4 w = __ldcg(reinterpret_cast<const float4*>(ws_fx + woff));
              v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
            }
            if (EPI == EPI_F32) {
              if (epi.nostore) { if (v.x == 1.2345e30f) epi.out_f32[0] = v.y; continue; }
              *reinterpret_cast<float4*>(epi.out_f32 + off) = v;
            } else if (EPI == EPI_UNPATCH) {
              const int p = epi.unpatch_p, g = epi.unpatch_g, R = p * g;
              const int s = m / (g * g), pr = m - s * g * g, gy = pr / g, gx = pr - gy * g;
              const int ch = col / (p * p), rem = col - ch * p * p, py = rem / p, px = rem - py * p;
              *reinterpret_cast<float4*>(epi.out_f32 + (((size_t)s * 3 + ch) * R + gy * p + py) * R + gx * p + px) = v;
            } else if (EPI == EPI_BIAS_RESID) {
              const float4 b = res4[cu][EPI == EPI_BIAS_RESID ? it : 0];
              v.x += bias4.x + b.x; v.y += bias4.y + b.y; v.z += bias4.z + b.z; v.w += bias4.w + b.w;
              *reinterpret_cast<float4*>(epi.out_f32 + off) = v;
            } else {
              if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU) { v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w; }
              if (EPI == EPI_BIAS_GELU) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
                uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
                *reinterpret_cast<uint2*>(epi.out_pre + off) = u;
                v.x = quickgelu(v.x); v.y = quickgelu(v.y); v.z = quickgelu(v.z); v.w = quickgelu(v.w);
              }
              if (EPI == EPI_GELUGRAD_BF16) {
                const uint2 u = gin[cu][EPI == EPI_GELUGRAD_BF16 ? it : 0];
                const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
                const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
                v.x *= quickgelu_grad(h0.x); v.y *= quickgelu_grad(h0.y); v.z *= quickgelu_grad(h1.x); v.w *= quickgelu_grad(h1.y);
              }
              __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
              uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
              *reinterpret_cast<uint2*>(epi.out_bf16 + off) = u;
            }
          }
          __syncwarp();                    // staging tile is reused by this warp's next chunk
        }
        }
        if (p_store) {                     // publish this warp's part of the parked partial
          __threadfence();
          __syncwarp();
          if (lane == 0) *reinterpret_cast<volatile int*>(epi.sk_flags + (tile0 * CG + (int)rank) * GEMM_EPI_WARPS + (warp - 4)) = 1;
        } else if (p_fix) {                // consumed: re-arm the flag for the next launch
          __syncwarp();
          if (lane == 0) *reinterpret_cast<volatile int*>(epi.sk_flags + ((tile0 + 1) * CG + (int)rank) * GEMM_EPI_WARPS + (warp - 4)) = 0;
        }
        tile = ntile; kb0 = nkb0; kb1 = nkb1; have = have_n; ++tile_iter;
      }
  #undef APH_FETCH_OPS
    } else {
      uint32_t tile_iter = 0;
      WorkIter wi; wi.init(epi.sk, tile0, tile_step, num_tiles, k_blocks);
      int tile, kb0, kb1;
      for (; wi.next(tile, kb0, kb1); ++tile_iter) {
        const int m_blk = (tile / n_tiles) * CG + (int)rank, n_blk = tile % n_tiles;
        const uint32_t as = tile_iter & 1, aph_ = (tile_iter >> 1) & 1;
        // stream-K roles of this item: park the partial (tile continues from another group's k-blocks) or fix it up
        const bool p_store = epi.sk && kb0 > 0, p_fix = epi.sk && kb1 < k_blocks;
        float* ws_st = epi.sk_ws + ((size_t)(tile0 * CG + (int)rank) * GEMM_BM) * BN;             // slot of this group
        const float* ws_fx = epi.sk_ws + ((size_t)((tile0 + 1) * CG + (int)rank) * GEMM_BM) * BN;  // slot of the next group
        if (p_fix) {
          if (lane == 0) { volatile int* f = epi.sk_flags + ((tile0 + 1) * CG + (int)rank) * GEMM_EPI_WARPS + (warp - 4); const long long t0 = clock64(); while (*f == 0 && clock64() - t0 < 4000000000LL) { } }   /* bounded (~2 s): a protocol bug must fail a test, not hang the GPU */
          __threadfence();
          __syncwarp();
        }
        mbar_wait(&tfull_bar[as], aph_);
        tc_fence_after();
        const int m_base = m_blk * GEMM_BM + q * 32;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN;
  #pragma unroll 1
        for (int c = half; c < BN / 32; c += GEMM_EPI_WARPS / 4) {
          // global operands of the fused epilogue are fetched FIRST (8 independent loads in flight per lane): issued inside
          // the store loop they serialise behind the stores (possible aliasing) and the epilogue becomes load-latency bound
          const int col = n_blk * BN + c * 32 + 4 * c4;
          const size_t off0 = (size_t)(m_base + rsub) * shp.N + col;
          const size_t step = (size_t)4 * shp.N;
          float4 res4[EPI == EPI_BIAS_RESID ? 8 : 1];
          uint2 gin[EPI == EPI_GELUGRAD_BF16 ? 8 : 1];
          if ((EPI == EPI_BIAS_RESID || EPI == EPI_GELUGRAD_BF16) && !p_store) {
  #pragma unroll
            for (int it = 0; it < 8; ++it) {
              const bool ok = m_base + it * 4 + rsub < shp.M;
              if (EPI == EPI_BIAS_RESID) res4[it] = ok ? __ldg(reinterpret_cast<const float4*>(epi.resid + off0 + it * step)) : make_float4(0.f, 0.f, 0.f, 0.f);
              if (EPI == EPI_GELUGRAD_BF16) gin[it] = ok ? __ldg(reinterpret_cast<const uint2*>(epi.gelu_in + off0 + it * step)) : make_uint2(0u, 0u);
            }
          }
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU || EPI == EPI_BIAS_RESID) bias4 = __ldg(reinterpret_cast<const float4*>(epi.bias + col));
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_wait_ld();
          if (c + GEMM_EPI_WARPS / 4 >= BN / 32) {          // this warp's last read of the accumulator: hand the TMEM stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0)); else mbar_arrive(&tempty_bar[as]); }
          }
  #pragma unroll
          for (int i = 0; i < 8; ++i)      // lane = row: write its 32 columns as 8 swizzled 16-byte chunks
            sts128(stage + lane * 128 + ((i ^ (lane & 7)) << 4), r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          __syncwarp();
          size_t off = off0;
  #pragma unroll
          for (int it = 0; it < 8; ++it, off += step) {
            const int rr = it * 4 + rsub;
            const int m = m_base + rr;
            if (m >= shp.M) break;
            float4 v = lds128(stage + rr * 128 + ((c4 ^ (rr & 7)) << 4));
            if (p_store || p_fix) {
              const size_t woff = (size_t)(q * 32 + rr) * BN + c * 32 + 4 * c4;
              if (p_store) { *reinterpret_cast<float4*>(ws_st + woff) = v; continue; }
              const float4 w = __ldcg(reinterpret_cast<const float4*>(ws_fx + woff));
              v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
            }
            if (EPI == EPI_F32) {
              if (epi.nostore) { if (v.x == 1.2345e30f) epi.out_f32[0] = v.y; continue; }
              *reinterpret_cast<float4*>(epi.out_f32 + off) = v;
            } else if (EPI == EPI_UNPATCH) {
              const int p = epi.unpatch_p, g = epi.unpatch_g, R = p * g;
              const int s = m / (g * g), pr = m - s * g * g, gy = pr / g, gx = pr - gy * g;
              const int ch = col / (p * p), rem = col - ch * p * p, py = rem / p, px = rem - py * p;
              *reinterpret_cast<float4*>(epi.out_f32 + (((size_t)s * 3 + ch) * R + gy * p + py) * R + gx * p + px) = v;
            } else if (EPI == EPI_BIAS_RESID) {
              const float4 b = res4[EPI == EPI_BIAS_RESID ? it : 0];
              v.x += bias4.x + b.x; v.y += bias4.y + b.y; v.z += bias4.z + b.z; v.w += bias4.w + b.w;
              *reinterpret_cast<float4*>(epi.out_f32 + off) = v;
            } else {
              if (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU) { v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w; }
              if (EPI == EPI_BIAS_GELU) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
                uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
                *reinterpret_cast<uint2*>(epi.out_pre + off) = u;
                v.x = quickgelu(v.x); v.y = quickgelu(v.y); v.z = quickgelu(v.z); v.w = quickgelu(v.w);
              }
              if (EPI == EPI_GELUGRAD_BF16) {
                const uint2 u = gin[EPI == EPI_GELUGRAD_BF16 ? it : 0];
                const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
                const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
                v.x *= quickgelu_grad(h0.x); v.y *= quickgelu_grad(h0.y); v.z *= quickgelu_grad(h1.x); v.w *= quickgelu_grad(h1.y);
              }
              __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
              uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
              *reinterpret_cast<uint2*>(epi.out_bf16 + off) = u;
            }
          }
          __syncwarp();                    // staging tile is reused by this warp's next chunk
        }
        if (p_store) {                     // publish this warp's part of the parked partial
          __threadfence();
          __syncwarp();
          if (lane == 0) *reinterpret_cast<volatile int*>(epi.sk_flags + (tile0 * CG + (int)rank) * GEMM_EPI_WARPS + (warp - 4)) = 1;
        } else if (p_fix) {                // consumed: re-arm the flag for the next launch
          __syncwarp();
          if (lane == 0) *reinterpret_cast<volatile int*>(epi.sk_flags + ((tile0 + 1) * CG + (int)rank) * GEMM_EPI_WARPS + (warp - 4)) = 0;
        }
      }
    }
  }
  __syncwarp();               // single-lane roles rejoin their warp before the (warp-aligned) cluster barrier
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) { if (CG == 2) tmem_dealloc_cg2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---- host side ---------------------------------------------------------------------------------
// 2-D bf16 tensor map: tensor [rows, K] row-major, box [box_rows, 64] with 128B swizzle.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rows, int K, int box_rows);
int make_tmap_bf16_tokens(CUtensorMap* out, const void* base, int cols, int T, int S, int box_rows);
// Launches the GEMM on `st`. A: [M,K], B: [N,K] device bf16. Requires K % 64 == 0, N % 128 == 0.
int launch_gemm(const void* A, const void* B, GemmShape shp, const GemmEpi& epi, cudaStream_t st);

}  // namespace aph
