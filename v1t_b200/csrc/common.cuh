// Shared helpers for the v1t_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/v1t_b200.h"

namespace v1t {

void set_error(const char* fmt, ...);
void count_launch();

// Optional device-side phase timing (CUDA events on the launching stream), used by bench.py for the roofline
// figures: v1t_prof_enable(1) ... run ... v1t_prof_read(phase).  Disabled -> zero overhead besides a branch.
struct ProfScope {
  int slot;
  cudaStream_t st;
  ProfScope(int phase, cudaStream_t stream);
  ~ProfScope();
};

#define V1T_CHECK_ARG(cond, ...)                \
  do {                                          \
    if (!(cond)) {                              \
      v1t::set_error(__VA_ARGS__);              \
      return V1T_ERR_INVALID;                   \
    }                                           \
  } while (0)

#define V1T_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      v1t::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return V1T_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define V1T_LAUNCH_CHECK()                                                                     \
  do {                                                                                         \
    v1t::count_launch();                                                                       \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      v1t::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return V1T_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define V1T_TRY(expr)        \
  do {                       \
    int _rc = (expr);        \
    if (_rc != V1T_OK) return _rc; \
  } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- counter-based RNG for dropout masks; replayable in backward -------------------------------------
// Philox4x32 with 7 rounds (the fewest that Salmon et al. 2011 report as Crush-resistant): one call yields 128 bits =
// EIGHT 16-bit uniforms, i.e. the keep decisions of 8 adjacent elements (keep iff u16 >= round(p * 65536)).  The
// masks cost 11 % of the training step with the 10-round / 4-elements-per-call version; every site indexes its
// elements with a row stride rounded up to 8 (drop_stride) so that a call never straddles two rows.
__host__ __device__ inline int64_t drop_stride(int64_t cols) { return (cols + 7) & ~(int64_t)7; }
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr_lo, uint32_t ctr_hi0, uint32_t ctr_hi1) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = ctr_hi0, c3 = ctr_hi1;
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint32_t dropout_thr16(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }
// keep-multipliers (0 or 1/(1-p)) of the aligned group `group` (elements 8*group .. 8*group+7): ONE Philox call
__device__ __forceinline__ void dropout_mult8(uint64_t seed, uint32_t site, uint64_t group, float p, float inv_keep,
                                              float (&m)[8]) {
  const uint4 r = philox4x32(seed, group, site, 0x5eedu);
  const uint32_t thr = dropout_thr16(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[2 * i] = (w[i] & 0xffffu) >= thr ? inv_keep : 0.0f;
    m[2 * i + 1] = (w[i] >> 16) >= thr ? inv_keep : 0.0f;
  }
}
// the same eight keep decisions as a bit mask (bit e = element 8*group + e is KEPT)
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint32_t site, uint64_t group, float p) {
  const uint4 r = philox4x32(seed, group, site, 0x5eedu);
  const uint32_t thr = dropout_thr16(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bits |= ((w[i] & 0xffffu) >= thr ? 1u : 0u) << (2 * i);
    bits |= ((w[i] >> 16) >= thr ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}
// keep-multiplier of element `idx` of dropout site `site`: call idx >> 3, 16-bit lane idx & 7
__device__ __forceinline__ float dropout_mult(uint64_t seed, uint32_t site, uint64_t idx, float p, float inv_keep) {
  const uint4 r = philox4x32(seed, idx >> 3, site, 0x5eedu);
  const uint32_t l = (uint32_t)idx & 7u;
  const uint32_t w = (l >> 1) == 0 ? r.x : (l >> 1) == 1 ? r.y : (l >> 1) == 2 ? r.z : r.w;
  const uint32_t u = (l & 1) ? (w >> 16) : (w & 0xffffu);
  return u >= dropout_thr16(p) ? inv_keep : 0.0f;
}


// Exact-erf GELU for the tensor-core epilogues: erf by Abramowitz-Stegun 7.1.26 (|error| <= 5e-7 in fp32, against
// the 2e-5 of the bf16x3 contractions) with ONE ex2.approx shared between erf and the Gaussian of the derivative:
// ~15 instructions per element instead of ~60 for erff + expf (the GELU-gradient epilogue was ALU-bound).
__device__ __forceinline__ void erf_gauss(float u, float& erf_v, float& gauss /* exp(-u^2/2) */) {
  const float x = u * 0.70710678118654752f, ax = fabsf(x);
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * x * 1.4426950408889634f));
  float p = 1.061405429f;
  p = fmaf(p, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  erf_v = copysignf(fmaf(-p * t, e, 1.0f), x);
  gauss = e;
}
__device__ __forceinline__ float gelu_fast_f(float u) {
  float er, g;
  erf_gauss(u, er, g);
  return 0.5f * u * (1.f + er);
}
__device__ __forceinline__ float gelu_fast_df(float u) {
  float er, g;
  erf_gauss(u, er, g);
  return fmaf(u * 0.3989422804014327f, g, 0.5f * (1.f + er));
}
__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_df(float u) {
  return 0.5f * (1.f + erff(u * 0.70710678118654752f)) + u * 0.3989422804014327f * expf(-0.5f * u * u);
}

// 4 multipliers of the aligned quad `quad` (elements 4*quad .. 4*quad+3): half of the Philox call quad >> 1
__device__ __forceinline__ void dropout_mult4(uint64_t seed, uint32_t site, uint64_t quad, float p, float inv_keep,
                                              float (&m)[4]) {
  const uint4 r = philox4x32(seed, quad >> 1, site, 0x5eedu);
  const uint32_t thr = dropout_thr16(p);
  const uint32_t w0 = (quad & 1) ? r.z : r.x, w1 = (quad & 1) ? r.w : r.y;
  m[0] = (w0 & 0xffffu) >= thr ? inv_keep : 0.0f;
  m[1] = (w0 >> 16) >= thr ? inv_keep : 0.0f;
  m[2] = (w1 & 0xffffu) >= thr ? inv_keep : 0.0f;
  m[3] = (w1 >> 16) >= thr ? inv_keep : 0.0f;
}

}  // namespace v1t
