// Ensemble output module: combine the K members' pre-activation responses and apply ELU+1 in one
// HBM-bound pass, without materialising the [B,N,K] stack the reference builds.
//   reference: ensemble.py:131-151 (EnsembleModel.forward: per member Model(..., activate=False), rearrange to
//              "b d 1", torch.cat on the last dim) and ensemble.py:30-80 (OutputModule: mode 0 mean over members,
//              mode 1/2 nn.Linear(K -> 1) shared / per mouse, then ELU1).
//   y[i] = elu(sum_k w[k] x_k[i] + bias) + 1 ;  mean = (w[k] = 1/K, bias = 0), computed as sum / K like torch.mean.
// Backward (the output module is what fit_ensemble trains; the members are frozen, ensemble.py:106):
//   dz = dy * elu'(z);  dw[k] = sum_i dz[i] x_k[i];  db = sum_i dz[i]   (per-CTA partials, fixed-order finish).
#include <algorithm>

#include "common.cuh"

namespace v1t {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = kNumSMs * 8;

__device__ __forceinline__ float combine(const v1t_ensemble_members& m, const float* __restrict__ w,
                                         const float* __restrict__ bias, int64_t i) {
  float z = 0.f;
  if (w) {
    for (int k = 0; k < m.count; ++k) z = fmaf(__ldg(w + k), __ldcs(m.x[k] + i), z);
    if (bias) z += __ldg(bias);
  } else {
    for (int k = 0; k < m.count; ++k) z += __ldcs(m.x[k] + i);
    z /= (float)m.count;
  }
  return z;
}

__global__ void __launch_bounds__(kThreads) ensemble_forward_kernel(v1t_ensemble_members m,
                                                                    const float* __restrict__ w,
                                                                    const float* __restrict__ bias, int64_t n,
                                                                    float* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const float z = combine(m, w, bias, i);
    y[i] = (z > 0.f ? z : expm1f(z)) + 1.f;
  }
}

// partials [gridDim.x][count + 1]: dw[0..count-1], db
__global__ void __launch_bounds__(kThreads) ensemble_backward_kernel(v1t_ensemble_members m,
                                                                     const float* __restrict__ w,
                                                                     const float* __restrict__ bias,
                                                                     const float* __restrict__ dy, int64_t n,
                                                                     float* __restrict__ partials) {
  __shared__ float red[kThreads / 32][V1T_ENSEMBLE_MAX + 1];
  float acc[V1T_ENSEMBLE_MAX + 1];
#pragma unroll
  for (int k = 0; k <= V1T_ENSEMBLE_MAX; ++k) acc[k] = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const float z = combine(m, w, bias, i);
    const float dz = dy[i] * (z > 0.f ? 1.f : expf(z));
#pragma unroll
    for (int k = 0; k < V1T_ENSEMBLE_MAX; ++k)
      if (k < m.count) acc[k] = fmaf(dz, m.x[k][i], acc[k]);
    acc[V1T_ENSEMBLE_MAX] += dz;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k <= V1T_ENSEMBLE_MAX; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == 0) red[wid][k] = s;
  }
  __syncthreads();
  if (threadIdx.x <= m.count) {
    const int k = threadIdx.x == m.count ? V1T_ENSEMBLE_MAX : threadIdx.x;
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < kThreads / 32; ++wv) s += red[wv][k];
    partials[(int64_t)blockIdx.x * (m.count + 1) + threadIdx.x] = s;
  }
}

__global__ void ensemble_finish_kernel(const float* __restrict__ partials, int parts, int count,
                                       float* __restrict__ dw, float* __restrict__ db) {
  const int k = threadIdx.x;
  if (k > count) return;
  float s = 0.f;
  for (int p = 0; p < parts; ++p) s += partials[(int64_t)p * (count + 1) + k];
  if (k < count) {
    if (dw) dw[k] = s;
  } else if (db) {
    *db = s;
  }
}

int blocks_for(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + kThreads - 1) / kThreads, kMaxBlocks)); }

int check(const v1t_ensemble_members* m, int64_t n) {
  V1T_CHECK_ARG(m, "ensemble: null member table");
  V1T_CHECK_ARG(m->count >= 1 && m->count <= V1T_ENSEMBLE_MAX, "ensemble: %d members outside 1..%d", m->count,
                V1T_ENSEMBLE_MAX);
  V1T_CHECK_ARG(n >= 0, "ensemble: negative size");
  for (int k = 0; k < m->count; ++k) V1T_CHECK_ARG(m->x[k] || n == 0, "ensemble: member %d is null", k);
  return V1T_OK;
}

}  // namespace
}  // namespace v1t

using namespace v1t;

extern "C" size_t v1t_ensemble_scratch_bytes(int64_t n, int count) {
  return sizeof(float) * (size_t)blocks_for(n) * (size_t)(count + 1);
}

extern "C" int v1t_ensemble_forward(const v1t_ensemble_members* members, const float* weight, const float* bias,
                                    int64_t n, float* y, void* stream) {
  V1T_TRY(check(members, n));
  if (n == 0) return V1T_OK;
  V1T_CHECK_ARG(y, "ensemble_forward: null output");
  ensemble_forward_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(*members, weight, bias, n, y);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

extern "C" int v1t_ensemble_backward(const v1t_ensemble_members* members, const float* weight, const float* bias,
                                     const float* dy, int64_t n, float* d_weight, float* d_bias, void* scratch,
                                     void* stream) {
  V1T_TRY(check(members, n));
  V1T_CHECK_ARG(weight && dy && scratch, "ensemble_backward: needs the linear output module's weight, dy and scratch");
  V1T_CHECK_ARG(n > 0, "ensemble_backward: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  const int parts = blocks_for(n);
  ensemble_backward_kernel<<<parts, kThreads, 0, st>>>(*members, weight, bias, dy, n, (float*)scratch);
  V1T_LAUNCH_CHECK();
  ensemble_finish_kernel<<<1, 32, 0, st>>>((const float*)scratch, parts, members->count, d_weight, d_bias);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
