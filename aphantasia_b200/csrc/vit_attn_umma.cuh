// vit_attn_umma.cuh -- tcgen05 / TMEM attention core of the CLIP ViT for short sequences (T <= 64: ViT-B/32, T = 50), forward.
//
// softmax(Q K^T / 8) V per (sample, head), head dim 64. Two (sample, head) items are stacked into one 128-row UMMA tile:
//   S[128 x 128] = [Q_a; Q_b] . [K_a; K_b]^T          tcgen05.mma kind::f16, A / B from 128B-swizzled shared memory (TMA)
//   softmax on the two diagonal 64 x 64 blocks          one thread per row: tcgen05.ld -> exp2 -> bf16 P
//   P stays in TENSOR MEMORY (tcgen05.st over the columns S occupied; the off-diagonal half is written as zeros)
//   O[128 x 64]  = P[128 x 128] . [V_a; V_b]            tcgen05.mma with A FROM TMEM and V as an MN-major (transposed) B operand
//   O -> bf16 -> swizzled shared tile -> cp.async.bulk.tensor store (token rows >= T are clipped by the [S][T][D] tensor map)
// Operand tiles of the next pair are prefetched by TMA into the other half of a double buffer (rows >= T arrive zero-filled).
// Warps 0-3: softmax / epilogue (thread = TMEM lane = row); warp 4 (one lane): TMA producer + MMA issuer; warp 5: TMEM allocator.
// Two CTAs per SM (96 KB shared, 256 TMEM columns each) overlap one pair's softmax with the other's MMAs and loads.
#pragma once
#include "tc_gemm.cuh"

namespace aph {

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem desc]      (A operand read from tensor memory: lane = row, one 32-bit column = two bf16 along K)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src_saddr, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_saddr), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

constexpr int AU_TILE = 128 * 128;                 // 128 rows x 64 bf16: one stacked operand tile (two items), bytes
constexpr int AU_BUF = 3 * AU_TILE;                // Q | K | V of one pair
constexpr int AU_S_COL = 0, AU_O_COL = 128;        // TMEM columns: S (fp32, 128) -- P (bf16 pairs, 64) aliases S's first 64 -- and O (fp32, 64)
constexpr uint32_t AU_TMEM_COLS = 256;
constexpr int AU_THREADS = 192;

__global__ void __launch_bounds__(AU_THREADS, 2)
k_attn_fwd_umma(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, int T, int D, int heads, int items) {
  pdl_trigger(); pdl_wait();
  extern __shared__ uint8_t au_raw[];
  __shared__ __align__(8) uint64_t full[2], bar_s, bar_p, bar_o, bar_free;
  __shared__ uint32_t tmem_slot;
  const uint32_t sm_a = (smem_u32(au_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npairs = (items + 1) >> 1;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&bar_s, 1); mbar_init(&bar_p, 128); mbar_init(&bar_o, 1); mbar_init(&bar_free, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_in); tma_prefetch_desc(&tm_out);
  }
  if (warp == 5) tmem_alloc(&tmem_slot, AU_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      // ===== TMA producer + MMA issuer
      auto issue_loads = [&](int pair, int b) {
        const int ia = 2 * pair, ib = min(2 * pair + 1, items - 1);           // an odd tail pairs the last item with itself (stored once)
        const int sa = ia / heads, ha = ia - sa * heads, sb = ib / heads, hb = ib - sb * heads;
        const uint32_t base = sm_a + b * AU_BUF;
        mbar_expect_tx(&full[b], AU_BUF);
        tma_load_3d(base, &tm_in, &full[b], ha * 64, 0, sa);                                   // Q_a
        tma_load_3d(base + AU_TILE / 2, &tm_in, &full[b], hb * 64, 0, sb);                     // Q_b
        tma_load_3d(base + AU_TILE, &tm_in, &full[b], D + ha * 64, 0, sa);                     // K_a
        tma_load_3d(base + AU_TILE + AU_TILE / 2, &tm_in, &full[b], D + hb * 64, 0, sb);       // K_b
        tma_load_3d(base + 2 * AU_TILE, &tm_in, &full[b], 2 * D + ha * 64, 0, sa);             // V_a
        tma_load_3d(base + 2 * AU_TILE + AU_TILE / 2, &tm_in, &full[b], 2 * D + hb * 64, 0, sb);   // V_b
      };
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = make_idesc_bf16(64, 128) | (1u << 16);     // B (= V) is MN-major: rows of the tile are K (keys)
      int n = 0;
      if (blockIdx.x < npairs) issue_loads(blockIdx.x, 0);
      for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++n) {
        const int b = n & 1;
        const uint32_t ph = (uint32_t)n & 1u;
        mbar_wait(&full[b], (uint32_t)(n >> 1) & 1u);
        tc_fence_after();
        const uint32_t base = sm_a + b * AU_BUF;
        const uint64_t dq = make_smem_desc(base), dk = make_smem_desc(base + AU_TILE);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base + AU_S_COL, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0);
        umma_commit(&bar_s);
        // the other buffer (pair n-1) is free once that pair's epilogue has stored its output; prefetch pair n+1 into it
        if (n > 0) mbar_wait(&bar_free, (uint32_t)(n - 1) & 1u);
        if (pair + (int)gridDim.x < npairs) issue_loads(pair + gridDim.x, b ^ 1);
        mbar_wait(&bar_p, ph);
        tc_fence_after();
        const uint64_t dv = make_smem_desc(base + 2 * AU_TILE);
#pragma unroll
        for (int k = 0; k < 8; ++k)      // 16 keys per instruction: 8 TMEM columns of P, 16 rows (2048 B) of V
          umma_f16_ts(tmem_base + AU_O_COL, tmem_base + AU_S_COL + 8 * k, dv + (uint64_t)(128 * k), idesc_o, k != 0);
        umma_commit(&bar_o);
      }
    }
  } else if (warp < 4) {
    // ===== softmax + epilogue: thread = row of the stacked tile
    const int row = threadIdx.x, blk = row >> 6, r = row & 63;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    int n = 0;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x, ++n) {
      const int b = n & 1;
      const uint32_t ph = (uint32_t)n & 1u;
      mbar_wait(&bar_s, ph);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(lane_addr + AU_S_COL + blk * 64, s0);
      tmem_ld_32x32(lane_addr + AU_S_COL + blk * 64 + 32, s1);
      tmem_wait_ld();
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float a = (j < T) ? __uint_as_float(s0[j]) * kAttnScaleLog2 : -INFINITY;
        const float c = (j + 32 < T) ? __uint_as_float(s1[j]) * kAttnScaleLog2 : -INFINITY;
        s0[j] = __float_as_uint(a); s1[j] = __float_as_uint(c);
        mx = fmaxf(mx, fmaxf(a, c));
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float a = exp2f(__uint_as_float(s0[j]) - mx), c = exp2f(__uint_as_float(s1[j]) - mx);
        s0[j] = __float_as_uint(a); s1[j] = __float_as_uint(c);
        sum += a + c;
      }
      const float inv = 1.f / sum;
      // P row as bf16 pairs: 32 words for this item's 64 keys, 32 zero words for the other item's keys
      uint32_t p[32], z[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        p[j] = pack_bf16(__uint_as_float(s0[2 * j]) * inv, __uint_as_float(s0[2 * j + 1]) * inv);
        p[16 + j] = pack_bf16(__uint_as_float(s1[2 * j]) * inv, __uint_as_float(s1[2 * j + 1]) * inv);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0u;
      tmem_st_32x32(lane_addr + AU_S_COL + blk * 32, p);
      tmem_st_32x32(lane_addr + AU_S_COL + (blk ^ 1) * 32, z);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bar_p);
      // ---- epilogue
      mbar_wait(&bar_o, ph);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(lane_addr + AU_O_COL, o0);
      tmem_ld_32x32(lane_addr + AU_O_COL + 32, o1);
      tmem_wait_ld();
      const uint32_t stage = sm_a + b * AU_BUF + row * 128;          // the Q tile of this pair is dead: reuse it as the output staging tile
      const int sw = row & 7;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        sts128(stage + ((c ^ sw) << 4), pack_bf16(__uint_as_float(o0[8 * c]), __uint_as_float(o0[8 * c + 1])), pack_bf16(__uint_as_float(o0[8 * c + 2]), __uint_as_float(o0[8 * c + 3])),
               pack_bf16(__uint_as_float(o0[8 * c + 4]), __uint_as_float(o0[8 * c + 5])), pack_bf16(__uint_as_float(o0[8 * c + 6]), __uint_as_float(o0[8 * c + 7])));
        sts128(stage + (((4 + c) ^ sw) << 4), pack_bf16(__uint_as_float(o1[8 * c]), __uint_as_float(o1[8 * c + 1])), pack_bf16(__uint_as_float(o1[8 * c + 2]), __uint_as_float(o1[8 * c + 3])),
               pack_bf16(__uint_as_float(o1[8 * c + 4]), __uint_as_float(o1[8 * c + 5])), pack_bf16(__uint_as_float(o1[8 * c + 6]), __uint_as_float(o1[8 * c + 7])));
      }
      (void)r;
      fence_proxy_async();
      tc_fence_before();
      named_bar_sync(1, 128);
      if (threadIdx.x == 0) {
        const int ia = 2 * pair, ib = 2 * pair + 1;
        const int sa = ia / heads, ha = ia - sa * heads;
        tma_store_3d(&tm_out, sm_a + b * AU_BUF, ha * 64, 0, sa);
        if (ib < items) { const int sb = ib / heads, hb = ib - sb * heads; tma_store_3d(&tm_out, sm_a + b * AU_BUF + AU_TILE / 2, hb * 64, 0, sb); }
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(&bar_free);
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, AU_TMEM_COLS);
}

// S items of `heads` heads each; qkv [S*T, 3D] bf16, out [S*T, D] bf16
static int attn_fwd_umma_launch(const bf16* qkv, bf16* out, int S, int T, int D, int heads, cudaStream_t st) {
  constexpr size_t smem = (size_t)2 * AU_BUF + 1024;
  static bool cfg = false;
  if (!cfg) {
    APH_CUDA_OK(cudaFuncSetAttribute(k_attn_fwd_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    APH_CUDA_OK(cudaFuncSetAttribute(k_attn_fwd_umma, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    cfg = true;
  }
  CUtensorMap tm_in, tm_out;
  if (int e = make_tmap_bf16_tokens(&tm_in, qkv, 3 * D, T, S, 64)) return e;
  if (int e = make_tmap_bf16_tokens(&tm_out, out, D, T, S, 64)) return e;
  const int items = S * heads, npairs = (items + 1) / 2;
  const int grid = npairs < 2 * kNumSMs ? npairs : 2 * kNumSMs;
  APH_CUDA_OK(launch_k(k_attn_fwd_umma, dim3(grid), dim3(AU_THREADS), smem, st, 1, tm_in, tm_out, T, D, heads, items));
  APH_LAUNCH_OK();
  return 0;
}

}  // namespace aph
