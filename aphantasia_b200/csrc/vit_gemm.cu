// vit_gemm.cu -- host side of the tcgen05 GEMM (tensor-map encoding, launch) + the exported test entry.
#include "tc_gemm.cuh"
#include <atomic>
#include <mutex>
#include <vector>
#include <utility>
#include <stdlib.h>

namespace aph {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  APH_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  APH_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && K % 8 == 0, "tensor map: base/stride not 16-byte aligned");
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  APH_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: CUresult %d (rows=%d K=%d box_rows=%d)", (int)r, rows, K, box_rows);
  return 0;
}

// 3-D view [S][T][cols] of a token-major bf16 matrix [S*T, cols]: a box of `box_rows` tokens x 64 columns of ONE sample; rows past T
// are out of bounds in the T dimension and arrive zero-filled (the attention kernels rely on that for their padded tiles).
int make_tmap_bf16_tokens(CUtensorMap* out, const void* base, int cols, int T, int S, int box_rows) {
  EncodeTiledFn enc = get_encode();
  APH_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  APH_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && cols % 8 == 0, "tensor map: base/stride not 16-byte aligned");
  const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)T, (cuuint64_t)S};
  const cuuint64_t gstride[2] = {(cuuint64_t)cols * 2, (cuuint64_t)T * cols * 2};
  const cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  APH_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (tokens) failed: CUresult %d (cols=%d T=%d S=%d box_rows=%d)", (int)r, cols, T, S, box_rows);
  return 0;
}

// ---- optional per-launch event timing (bench.py's roofline: the GEMM kernel's real time inside a step)
static bool g_prof = false;
bool gemm_profiling_on() { return g_prof; }
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_ev;
static std::vector<double> g_prof_flops;

// stream-K scratch (one in-flight GEMM at a time: all launches of this library go to the caller's single stream)
static float* g_sk_ws = nullptr;
static int* g_sk_flags = nullptr;

static int sk_prepare() {
  if (g_sk_ws) return 0;
  APH_CUDA_OK(cudaMalloc(&g_sk_ws, (size_t)kNumSMs * GEMM_BM * 256 * sizeof(float)));
  APH_CUDA_OK(cudaMalloc(&g_sk_flags, (size_t)(kNumSMs + 2) * GEMM_EPI_WARPS * sizeof(int)));
  APH_CUDA_OK(cudaMemset(g_sk_flags, 0, (size_t)(kNumSMs + 2) * GEMM_EPI_WARPS * sizeof(int)));
  return 0;
}

// launches per (tile variant, epilogue kind): variant 0 = 128x128 single CTA, 1 = 128x256 single CTA, 2 = 256x256 CTA pair.
// Read by the tests to prove which kernel a shape really ran (aph_gemm_variant_launches).
static std::atomic<long long> g_variant_launches[3][EPI_KINDS];

template <int BN, int STAGES, int EPI, int CG = 1>
static int launch_cfg(const void* A, const void* B, GemmShape shp, const GemmEpi& epi_in, cudaStream_t st) {
  using L = GemmSmem<BN, STAGES, CG>;
  GemmEpi epi = epi_in;
  g_variant_launches[CG == 2 ? 2 : (BN == 256 ? 1 : 0)][EPI].fetch_add(1, std::memory_order_relaxed);
  static bool configured = false;
  if (!configured) {
    APH_CUDA_OK(cudaFuncSetAttribute(k_gemm_bf16_tn<BN, STAGES, EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  CUtensorMap ma, mb;
  if (int e = make_tmap_bf16(&ma, A, shp.M, shp.K, GEMM_BM)) return e;
  if (int e = make_tmap_bf16(&mb, B, shp.N, shp.K, BN / CG)) return e;
  const int tiles = ((shp.M + GEMM_BM * CG - 1) / (GEMM_BM * CG)) * (shp.N / BN);
  const int slots = kNumSMs / CG;
  const int groups = tiles < slots ? tiles : slots;
  const int grid = CG * groups;
  {
    // stream-K (opt-in, APH_GEMM_STREAMK=1): units = tiles x k-blocks split evenly over the groups, split tiles fixed up through
    // a parked fp32 partial. Correct (tests pass with it on) but MEASURED SLOWER at the ViT-B shapes (profiles/README.md): with
    // K = 768 a tile's epilogue costs as much as its main loop, so the extra partial-store / fix-up epilogue per group outweighs
    // the k-blocks saved by removing the ragged last round (9500x768x768: 31.4 vs 19.1 us; 9500x3072x768: 67.6 vs 47.7 us).
    static int sk_on = -1;
    if (sk_on < 0) { const char* e = getenv("APH_GEMM_STREAMK"); sk_on = (e && e[0] == '1') ? 1 : 0; }
    const int kb = shp.K / GEMM_BK;
    const long long units = (long long)tiles * kb;
    const long long cost_static = (long long)((tiles + groups - 1) / groups) * kb, cost_sk = (units + groups - 1) / groups;
    if (sk_on && tiles > groups && units / groups >= kb && cost_sk + 3 < cost_static) {
      if (int e = sk_prepare()) return e;
      epi.sk = 1; epi.sk_ws = g_sk_ws; epi.sk_flags = g_sk_flags;
    }
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_prof) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
  APH_CUDA_OK(launch_k(k_gemm_bf16_tn<BN, STAGES, EPI, CG>, dim3(grid), dim3(GEMM_THREADS), (size_t)L::TOTAL, st, CG, ma, mb, shp, epi));
  APH_LAUNCH_OK();
  if (g_prof) { cudaEventRecord(e1, st); g_prof_ev.emplace_back(e0, e1); g_prof_flops.push_back(2.0 * shp.M * shp.N * shp.K); }
  return 0;
}

int launch_gemm(const void* A, const void* B, GemmShape shp, const GemmEpi& epi, cudaStream_t st) {
  APH_REQUIRE(A && B && shp.M > 0, "gemm: null operand or empty M");
  APH_REQUIRE(shp.K % GEMM_BK == 0 && shp.K > 0, "gemm: K=%d must be a positive multiple of %d", shp.K, GEMM_BK);
  APH_REQUIRE(shp.N % 128 == 0 && shp.N > 0, "gemm: N=%d must be a positive multiple of 128", shp.N);
  // tile width: APH_GEMM_BN=128|256 forces one; default picks 256-wide tiles when N allows it and the grid still fills
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("APH_GEMM_BN"); forced = e ? atoi(e) : 0; }
  bool wide = (shp.N % 256 == 0);
  if (forced == 128) wide = false;
  else if (forced != 256 && wide) {
    const int mt = (shp.M + GEMM_BM - 1) / GEMM_BM;
    wide = mt * (shp.N / 256) >= kNumSMs;          // small problems keep the finer 128-wide tiling
  }
  // map the requested fusion onto one of the compiled epilogue kinds
  int kind = -1;
  const bool b = epi.bias, r = epi.resid, gi = epi.gelu_in, f = epi.out_f32, h = epi.out_bf16, pre = epi.out_pre, act = epi.act == 1, un = epi.unpatch_p > 0;
  if (un && f && !b && !r && !gi && !h && !pre && !act) kind = EPI_UNPATCH;
  else if (f && !h && !b && !r && !gi && !pre && !act) kind = EPI_F32;
  else if (h && !f && !b && !r && !gi && !pre && !act) kind = EPI_BF16;
  else if (h && !f && b && !r && !gi && !pre && !act) kind = EPI_BIAS_BF16;
  else if (h && !f && b && !r && !gi && pre && act) kind = EPI_BIAS_GELU;
  else if (f && !h && b && r && !gi && !pre && !act) kind = EPI_BIAS_RESID;
  else if (h && !f && !b && !r && gi && !pre && !act) kind = EPI_GELUGRAD_BF16;
  APH_REQUIRE(kind >= 0, "gemm: unsupported epilogue combination");
  // CTA-pair tiles (cta_group::2, 256 x 256 accumulator per pair, B tile split across the pair) are the default for large
  // problems: measured +3 % steps/s at C2 (profiles/README.md). APH_GEMM_2CTA=0 forces single-CTA tiles.
  static int pair = -1;
  if (pair < 0) { const char* e = getenv("APH_GEMM_2CTA"); pair = (e && e[0] == '0') ? 0 : 1; }
  const bool use_pair = pair && wide && ((shp.M + 255) / 256) * (shp.N / 256) >= kNumSMs / 2;
#define APH_GEMM_CASE(K) case K: return use_pair ? launch_cfg<256, 6, K, 2>(A, B, shp, epi, st) : (wide ? launch_cfg<256, 4, K>(A, B, shp, epi, st) : launch_cfg<128, 6, K>(A, B, shp, epi, st));
  switch (kind) {
    APH_GEMM_CASE(EPI_F32) APH_GEMM_CASE(EPI_BF16) APH_GEMM_CASE(EPI_BIAS_BF16) APH_GEMM_CASE(EPI_BIAS_GELU)
    APH_GEMM_CASE(EPI_BIAS_RESID) APH_GEMM_CASE(EPI_GELUGRAD_BF16) APH_GEMM_CASE(EPI_UNPATCH)
  }
#undef APH_GEMM_CASE
  return 2;
}

}  // namespace aph

using namespace aph;

extern "C" int aph_gemm_bf16_tn(const void* A, const void* B, float* C, int M, int N, int K, void* stream) {
  APH_REQUIRE(C != nullptr, "aph_gemm_bf16_tn: null output");
  GemmEpi epi;
  epi.out_f32 = C;
  { const char* e = getenv("APH_GEMM_NOSTORE"); epi.nostore = (e && e[0] == '1') ? 1 : 0; }
  return launch_gemm(A, B, GemmShape{M, N, K}, epi, (cudaStream_t)stream);
}

// Test entry: the encoder's GEMM with any of its fused epilogues, on caller-supplied operands (tests/test_gpu_bench_config.py).
// Unused pointers are NULL; the combination selects the epilogue kind exactly as the encoder's own calls do.
extern "C" int aph_gemm_epi_test(const void* A, const void* B, int M, int N, int K, const float* bias, const float* resid,
                                 const void* gelu_in, int act, float* out_f32, void* out_bf16, void* out_pre,
                                 int unpatch_p, int unpatch_g, void* stream) {
  GemmEpi epi;
  epi.bias = bias; epi.resid = resid; epi.gelu_in = reinterpret_cast<const bf16*>(gelu_in); epi.act = act;
  epi.out_f32 = out_f32; epi.out_bf16 = reinterpret_cast<bf16*>(out_bf16); epi.out_pre = reinterpret_cast<bf16*>(out_pre);
  epi.unpatch_p = unpatch_p; epi.unpatch_g = unpatch_g;
  return launch_gemm(A, B, GemmShape{M, N, K}, epi, (cudaStream_t)stream);
}

// launches so far of tile variant `variant` (0: 128x128 single CTA, 1: 128x256 single CTA, 2: 256x256 cta_group::2 pair) with
// epilogue kind `epi` (EPI_* order of tc_gemm.cuh: 0 f32, 1 bf16, 2 bias-bf16, 3 bias-gelu, 4 bias-resid, 5 gelugrad, 6 unpatch; -1 = all)
extern "C" int64_t aph_gemm_variant_launches(int variant, int epi) {
  if (variant < 0 || variant > 2 || epi >= EPI_KINDS) return -1;
  long long n = 0;
  for (int k = 0; k < EPI_KINDS; ++k) if (epi < 0 || epi == k) n += g_variant_launches[variant][k].load();
  return (int64_t)n;
}

// Profiling aid for bench.py: enable=1 starts recording a CUDA-event pair around every GEMM launch (on its stream);
// enable=0 stops, synchronises and returns the summed kernel time / FLOPs / launch count since it was enabled.
extern "C" int aph_prof_gemm(int enable, double* total_ms, double* total_flops, int* launches) {
  if (enable) { g_prof = true; g_prof_ev.clear(); g_prof_flops.clear(); return 0; }
  g_prof = false;
  double ms = 0., fl = 0.;
  for (size_t i = 0; i < g_prof_ev.size(); ++i) {
    APH_CUDA_OK(cudaEventSynchronize(g_prof_ev[i].second));
    float t = 0.f;
    APH_CUDA_OK(cudaEventElapsedTime(&t, g_prof_ev[i].first, g_prof_ev[i].second));
    ms += t; fl += g_prof_flops[i];
    cudaEventDestroy(g_prof_ev[i].first); cudaEventDestroy(g_prof_ev[i].second);
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = (int)g_prof_ev.size();
  g_prof_ev.clear(); g_prof_flops.clear();
  return 0;
}
