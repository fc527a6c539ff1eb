// Fused L1-regulariser + AdamW over ALL parameter tensors in one launch — HBM-bound.
//   reference: the regulariser terms  core  vit.py:419-421  (reg_scale * sum |p| over every core parameter),
//              readout gaussian2d.py:83-100 (reg_scale * sum |features|), shifter core_shifter.py:35-36,
//              summed in model.py:141-149, weighted and added to the loss in train.py:71-73 (autograd of |p| is
//              sign(p)), then torch.optim.AdamW(weight_decay=0) train.py:217-223 stepped at train.py:77-80.
//   AdamW arithmetic restated from PyTorch 2.x torch/optim/adamw.py (_single_tensor_adamw, amsgrad=False):
//       p *= 1 - lr*wd;  m += (g - m)(1 - b1);  v = b2 v + (1 - b2) g g;
//       p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//   with g = grad_scale * grad + l1 * sign(p): the regulariser's gradient never exists as a tensor, the ~100 small
//   abs/sum/sign kernels of the reference's autograd graph disappear, and sum |p| (what the reference logs as
//   reg_loss) falls out of the same pass over p.
//
// Data layout: a device table of v1t_opt_tensor records (one per parameter) and an int32 prefix array of chunk
// counts; CTA c binary-searches the prefix array for its tensor.  Per element the pass reads p, g, m, v and writes
// p, m, v (+ g = 0 when zero_grad): 28-32 B of algorithmic traffic, 128-bit accesses when the four pointers are
// 16-byte aligned.
#include "common.cuh"

namespace v1t {
namespace {

constexpr int kChunk = 4096;  // elements per CTA
constexpr int kThreads = 256;

struct AdamConsts {  // derived on the host in double precision, like torch derives 1 - beta from Python floats
  float one_minus_beta1, beta2, one_minus_beta2, eps, inv_bc1, inv_bc2_sqrt, grad_scale;
  int zero_grad;
};

__device__ __forceinline__ float signf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ void adam_one(float& p, float g_raw, float& m, float& v, const v1t_opt_tensor& t,
                                         const AdamConsts& k, float& abs_sum) {
  abs_sum += fabsf(p);
  const float g = fmaf(g_raw, k.grad_scale, t.l1 * signf(p));
  float pw = p * (1.f - t.lr * t.weight_decay);
  m = fmaf(g - m, k.one_minus_beta1, m);
  v = fmaf(v, k.beta2, k.one_minus_beta2 * g * g);
  const float denom = sqrtf(v) * k.inv_bc2_sqrt + k.eps;
  p = pw - (t.lr * k.inv_bc1) * (m / denom);
}

__global__ void __launch_bounds__(kThreads) adamw_l1_kernel(const v1t_opt_tensor* __restrict__ tensors,
                                                            const int32_t* __restrict__ chunk_prefix, int n_tensors,
                                                            AdamConsts k, float* __restrict__ abs_partials,
                                                            int32_t* __restrict__ chunk_group) {
  __shared__ v1t_opt_tensor t;
  __shared__ int64_t s_off;
  __shared__ float red[kThreads / 32];
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_tensors;  // last tensor with chunk_prefix[i] <= blockIdx.x
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (chunk_prefix[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    t = tensors[lo];
    s_off = (int64_t)((int)blockIdx.x - chunk_prefix[lo]) * kChunk;
  }
  __syncthreads();
  const int64_t off = s_off;
  const int64_t cnt = min((int64_t)kChunk, t.numel - off);
  float* p = t.param + off;
  float* g = t.grad + off;
  float* m = t.exp_avg + off;
  float* v = t.exp_avg_sq + off;
  float abs_sum = 0.f;
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15u) == 0;
  int64_t done = 0;
  if (vec) {
    const int64_t n4 = cnt >> 2;
    for (int64_t i = threadIdx.x; i < n4; i += kThreads) {
      float4 pp = reinterpret_cast<float4*>(p)[i], gg = reinterpret_cast<const float4*>(g)[i];
      float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, t, k, abs_sum);
      adam_one(pp.y, gg.y, mm.y, vv.y, t, k, abs_sum);
      adam_one(pp.z, gg.z, mm.z, vv.z, t, k, abs_sum);
      adam_one(pp.w, gg.w, mm.w, vv.w, t, k, abs_sum);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
      if (k.zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + threadIdx.x; i < cnt; i += kThreads) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_one(pp, g[i], mm, vv, t, k, abs_sum);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
    if (k.zero_grad) g[i] = 0.f;
  }
  if (abs_partials) {  // fixed-order block reduction -> one partial per chunk
    abs_sum = warp_sum(abs_sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = abs_sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) s += red[w];
      abs_partials[blockIdx.x] = s;
      chunk_group[blockIdx.x] = t.group;
    }
  }
}

// l1_sums[group] = sum of the chunk partials tagged with that group; fixed assignment of chunks to threads and a
// fixed-order block reduction -> deterministic
__global__ void __launch_bounds__(256) l1_finish_kernel(const float* __restrict__ abs_partials,
                                                        const int32_t* __restrict__ chunk_group, int n_chunks,
                                                        float* __restrict__ l1_sums) {
  __shared__ float red[8];
  const int group = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < n_chunks; c += blockDim.x)
    if (chunk_group[c] == group) s += abs_partials[c];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tsum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tsum += red[w];
    l1_sums[group] = tsum;
  }
}

}  // namespace
}  // namespace v1t

using namespace v1t;

extern "C" int v1t_opt_chunk_elems(void) { return kChunk; }

extern "C" size_t v1t_adamw_l1_scratch_bytes(int n_chunks) {  // per chunk: |p| partial (float) + group tag (int32)
  return (sizeof(float) + sizeof(int32_t)) * (size_t)(n_chunks > 0 ? n_chunks : 0);
}

extern "C" int v1t_adamw_l1_step(const v1t_opt_tensor* tensors_dev, const int32_t* chunk_prefix_dev, int n_tensors,
                                 int n_chunks, double beta1, double beta2, double eps, double bias_corr1,
                                 double bias_corr2_sqrt, double grad_scale, int zero_grad, float* l1_sums_dev,
                                 int n_groups, void* scratch, void* stream) {
  V1T_CHECK_ARG(n_tensors >= 0 && n_chunks >= 0, "adamw_l1_step: negative count");
  if (n_tensors == 0 || n_chunks == 0) return V1T_OK;
  V1T_CHECK_ARG(tensors_dev && chunk_prefix_dev, "adamw_l1_step: null table");
  V1T_CHECK_ARG(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0.,
                "adamw_l1_step: bad hyper-parameters (beta1 %g beta2 %g eps %g)", beta1, beta2, eps);
  V1T_CHECK_ARG(bias_corr1 > 0. && bias_corr2_sqrt > 0., "adamw_l1_step: bias corrections must be positive");
  V1T_CHECK_ARG(!l1_sums_dev || (n_groups > 0 && scratch), "adamw_l1_step: l1 sums need n_groups and scratch");
  cudaStream_t st = (cudaStream_t)stream;
  AdamConsts k{(float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps, (float)(1.0 / bias_corr1),
               (float)(1.0 / bias_corr2_sqrt), (float)grad_scale, zero_grad};
  float* partials = l1_sums_dev ? (float*)scratch : nullptr;
  int32_t* groups = l1_sums_dev ? (int32_t*)((float*)scratch + n_chunks) : nullptr;
  adamw_l1_kernel<<<n_chunks, kThreads, 0, st>>>(tensors_dev, chunk_prefix_dev, n_tensors, k, partials, groups);
  V1T_LAUNCH_CHECK();
  if (l1_sums_dev) {
    l1_finish_kernel<<<n_groups, 256, 0, st>>>(partials, groups, n_chunks, l1_sums_dev);
    V1T_LAUNCH_CHECK();
  }
  return V1T_OK;
}
