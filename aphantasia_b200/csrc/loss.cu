// loss.cu -- similarity loss (fused value + gradients) and the Adam update. fp32.
//
// aph_sim_fwd replaces sim_func(v1, v2, type) for type None/'cossim' and 'mix'
// (/root/reference/aphantasia/utils.py:276-282,295):
//     cos_s = <v1,v2_s> / (|v1| |v2_s|);   mix: f = cos - 0.25 * 2*asin(|v1^ - v2^_s| / 2)^2
//     value = mean_s f_s
// One warp per sample; gradients w.r.t. both operands are written in the same pass.
// aph_adam_step replaces torch.optim.Adam's single-tensor update (/root/reference/clip_fft.py:115,295).
#include "aph_common.cuh"
#include <algorithm>
#include <math.h>

namespace aph {

__global__ void __launch_bounds__(256) k_sim(const float* __restrict__ v1, int n1, const float* __restrict__ v2, int S, int D,
                                             int kind, float* __restrict__ value, float* __restrict__ g1, float* __restrict__ g2) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= S) return;
  const float* a = v1 + (size_t)(n1 == 1 ? 0 : warp) * D;
  const float* b = v2 + (size_t)warp * D;
  float dot = 0.f, na = 0.f, nb = 0.f;
  for (int i = lane; i < D; i += 32) { const float x = a[i], y = b[i]; dot += x * y; na += x * x; nb += y * y; }
  dot = warp_sum(dot); na = warp_sum(na); nb = warp_sum(nb);
  const float eps = 1e-8f;
  const float la = fmaxf(sqrtf(na), eps), lb = fmaxf(sqrtf(nb), eps);
  const float c = dot / (la * lb);
  float f = c, dfdc = 1.f;
  if (kind == APH_SIM_MIX) {
    // d^2 = |a^ - b^|^2 computed explicitly, as the reference does (F.normalize eps = 1e-12)
    const float ia = 1.f / fmaxf(sqrtf(na), 1e-12f), ib = 1.f / fmaxf(sqrtf(nb), 1e-12f);
    float d2 = 0.f;
    for (int i = lane; i < D; i += 32) { const float t = a[i] * ia - b[i] * ib; d2 += t * t; }
    d2 = warp_sum(d2);
    const float u = fminf(0.5f * sqrtf(d2), 1.f);
    const float as = asinf(u);
    f = c - 0.25f * fabsf(2.f * as * as);
    // f(c) = c - theta^2/8 with theta = 2 asin(u): df/dc = 1 + theta / (4 sin theta)
    const float theta = 2.f * as, st = sinf(theta);
    dfdc = 1.f + ((st > 1e-6f) ? theta / (4.f * st) : 0.25f);
  }
  const float invS = 1.f / (float)S;
  if (lane == 0) atomicAdd(value, f * invS);
  const float k = dfdc * invS;
  if (g2) {
    float* o = g2 + (size_t)warp * D;
    for (int i = lane; i < D; i += 32) o[i] = k * (a[i] / la - c * b[i] / lb) / lb;
  }
  if (g1) {
    if (n1 == 1) { for (int i = lane; i < D; i += 32) atomicAdd(&g1[i], k * (b[i] / lb - c * a[i] / la) / la); }
    else { float* o = g1 + (size_t)warp * D; for (int i = lane; i < D; i += 32) o[i] = k * (b[i] / lb - c * a[i] / la) / la; }
  }
}

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, size_t n, float step_size, float b1, float b2, float eps,
                                              float inv_sqrt_bc2) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

}  // namespace aph

using namespace aph;

extern "C" int aph_sim_fwd(const float* v1, int n1, const float* v2, int S, int D, int kind, float* value, float* grad_v1,
                           float* grad_v2, void* stream) {
  APH_REQUIRE(v1 && v2 && value && S > 0 && D > 0, "aph_sim_fwd: bad arguments");
  APH_REQUIRE(n1 == 1 || n1 == S, "aph_sim_fwd: v1 rows (%d) must be 1 or S (%d)", n1, S);
  APH_REQUIRE(kind == APH_SIM_COS || kind == APH_SIM_MIX, "aph_sim_fwd: unknown kind %d", kind);
  cudaStream_t st = (cudaStream_t)stream;
  APH_CUDA_OK(cudaMemsetAsync(value, 0, sizeof(float), st));
  if (grad_v1 && n1 == 1) APH_CUDA_OK(cudaMemsetAsync(grad_v1, 0, (size_t)D * sizeof(float), st));
  const int blocks = (S * 32 + 255) / 256;
  k_sim<<<blocks, 256, 0, st>>>(v1, n1, v2, S, D, kind, value, grad_v1, grad_v2);
  APH_LAUNCH_OK();
  return 0;
}

extern "C" int aph_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                             int step, void* stream) {
  APH_REQUIRE(p && g && m && v && n > 0 && step >= 1, "aph_adam_step: bad arguments");
  const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8);
  k_adam<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (size_t)n, (float)(lr / bc1), b1, b2, eps, (float)(1.0 / sqrt(bc2)));
  APH_LAUNCH_OK();
  return 0;
}
