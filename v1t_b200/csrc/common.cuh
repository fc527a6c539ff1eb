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

// ---- counter-based RNG for dropout masks (Philox4x32-10); replayable in backward -----------------------
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t ctr_lo, uint32_t ctr_hi0, uint32_t ctr_hi1) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = ctr_hi0, c3 = ctr_hi1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
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

// keep-multiplier of inverted dropout for element `idx` of dropout site `site`: 0 or 1/(1-p).
// One Philox call yields 4 lanes; element idx uses call idx>>2, lane idx&3.
__device__ __forceinline__ float dropout_mult(uint64_t seed, uint32_t site, uint64_t idx, float p, float inv_keep) {
  uint4 r = philox4x32(seed, idx >> 2, site, 0x5eedu);
  uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  // uniform in [0,1): keep iff u >= p  (matches "bernoulli(1-p)")
  float u = (float)(w >> 8) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.0f;
}


__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_df(float u) {
  return 0.5f * (1.f + erff(u * 0.70710678118654752f)) + u * 0.3989422804014327f * expf(-0.5f * u * u);
}

// 4 multipliers of the aligned group `group` (elements 4*group .. 4*group+3) with ONE Philox call
__device__ __forceinline__ void dropout_mult4(uint64_t seed, uint32_t site, uint64_t group, float p, float inv_keep,
                                              float (&m)[4]) {
  const uint4 r = philox4x32(seed, group, site, 0x5eedu);
  m[0] = (float)(r.x >> 8) * (1.0f / 16777216.0f) >= p ? inv_keep : 0.0f;
  m[1] = (float)(r.y >> 8) * (1.0f / 16777216.0f) >= p ? inv_keep : 0.0f;
  m[2] = (float)(r.z >> 8) * (1.0f / 16777216.0f) >= p ? inv_keep : 0.0f;
  m[3] = (float)(r.w >> 8) * (1.0f / 16777216.0f) >= p ? inv_keep : 0.0f;
}

}  // namespace v1t
