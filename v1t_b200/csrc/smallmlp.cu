// The two tiny MLPs either side of the readout, each as ONE forward and ONE backward kernel (+ a fixed-order finish):
//   grid predictor  mu = Tanh(Linear(ELU(Linear(source_grid))))  over N ~ 8000 neurons
//                   reference: gaussian2d.py:102-136 (init_grid_predictor), :188-193 (mu property)
//   core shifter    shifts = Tanh(Linear(Tanh(Linear(Tanh(Linear(pupil_center))))))  over the batch
//                   reference: core_shifter.py:24-40
// In eager PyTorch these are ~20 launches forward+backward per mouse, two of them cuBLAS GEMMs with K = N = 8000
// and 2..30 output columns (80 us each); here a thread owns a row (neuron / sample), the <= 2.3 k weights sit in
// shared memory, the backward recomputes the activations (nothing is saved) and reduces the weight gradients
// warp-shuffle -> per-warp smem slab -> per-CTA partial -> fixed-order sum, so it is deterministic and atomic-free.
#include "common.cuh"

namespace v1t {
namespace {

constexpr int kMaxL = V1T_MLP_MAX_LAYERS;
constexpr int kMaxW = V1T_MLP_MAX_WIDTH;
constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;

struct Layout {  // offsets (floats) of each layer's weight / bias in the flat parameter vector
  int w[kMaxL], b[kMaxL], total;
};

__host__ __device__ inline Layout layout_of(const v1t_mlp_spec& s) {
  Layout l;
  int off = 0;
  for (int i = 0; i < kMaxL; ++i) {
    l.w[i] = l.b[i] = off;
    if (i < s.layers) {
      l.w[i] = off;
      off += s.width[i] * s.width[i + 1];
      l.b[i] = off;
      off += s.width[i + 1];
    }
  }
  l.total = off;
  return l;
}

__device__ __forceinline__ float act_f(int kind, float x) {
  if (kind == V1T_ACT_TANH) return tanhf(x);
  if (kind == V1T_ACT_ELU) return x > 0.f ? x : expm1f(x);
  return x;
}
// derivative expressed through the activation's OUTPUT y (and the pre-activation sign for ELU)
__device__ __forceinline__ float act_df(int kind, float pre, float y) {
  if (kind == V1T_ACT_TANH) return 1.f - y * y;
  if (kind == V1T_ACT_ELU) return pre > 0.f ? 1.f : y + 1.f;
  return 1.f;
}

__device__ __forceinline__ void stage_params(const v1t_mlp_spec& s, const v1t_mlp_ptrs& p, const Layout& l,
                                             float* sp) {
  for (int i = 0; i < s.layers; ++i) {
    const int nw = s.width[i] * s.width[i + 1];
    for (int j = threadIdx.x; j < nw; j += blockDim.x) sp[l.w[i] + j] = p.w[i][j];
    for (int j = threadIdx.x; j < s.width[i + 1]; j += blockDim.x) sp[l.b[i] + j] = p.b[i] ? p.b[i][j] : 0.f;
  }
}

// a[l] = activations entering layer l (a[0] = x), pre[l] = pre-activation of layer l.  Arrays are indexed with
// runtime widths and live in local memory (L1-resident: at most 4 * 32 floats per thread); the kernels are
// launch-latency bound at these sizes.
__device__ __forceinline__ void forward_row(const v1t_mlp_spec& s, const Layout& l, const float* sp,
                                            const float* __restrict__ xrow, float (&a)[kMaxL + 1][kMaxW],
                                            float (&pre)[kMaxL][kMaxW]) {
  for (int j = 0; j < s.width[0]; ++j) a[0][j] = xrow[j];
  for (int i = 0; i < s.layers; ++i) {
    const int in = s.width[i], out = s.width[i + 1];
    for (int o = 0; o < out; ++o) {
      float t = sp[l.b[i] + o];
      for (int j = 0; j < in; ++j) t = fmaf(sp[l.w[i] + o * in + j], a[i][j], t);
      pre[i][o] = t;
      a[i + 1][o] = act_f(s.act[i], t);
    }
  }
}

__global__ void __launch_bounds__(kThreads) small_mlp_forward_kernel(v1t_mlp_spec s, v1t_mlp_ptrs p,
                                                                     const float* __restrict__ x,
                                                                     float* __restrict__ y) {
  extern __shared__ float sp[];
  const Layout l = layout_of(s);
  stage_params(s, p, l, sp);
  __syncthreads();
  const int r = blockIdx.x * kThreads + threadIdx.x;
  if (r >= s.rows) return;
  float a[kMaxL + 1][kMaxW], pre[kMaxL][kMaxW];
  forward_row(s, l, sp, x + (int64_t)r * s.x_ld, a, pre);
  const int out = s.width[s.layers];
  for (int o = 0; o < out; ++o) y[(int64_t)r * out + o] = a[s.layers][o];
}

// partials [gridDim.x][total]: this CTA's sum over its rows of every weight / bias gradient
__global__ void __launch_bounds__(kThreads) small_mlp_backward_kernel(v1t_mlp_spec s, v1t_mlp_ptrs p,
                                                                      const float* __restrict__ x,
                                                                      const float* __restrict__ dy,
                                                                      float* __restrict__ partials) {
  extern __shared__ float smem[];
  const Layout l = layout_of(s);
  float* sp = smem;                 // [total] parameters
  float* slab = smem + l.total;     // [kWarps][total] per-warp gradient sums
  stage_params(s, p, l, sp);
  __syncthreads();
  const int r = blockIdx.x * kThreads + threadIdx.x;
  const bool live = r < s.rows;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float a[kMaxL + 1][kMaxW], pre[kMaxL][kMaxW], d[kMaxW], dprev[kMaxW];
  const int out = s.width[s.layers];
  if (live) {
    forward_row(s, l, sp, x + (int64_t)r * s.x_ld, a, pre);
    for (int o = 0; o < out; ++o) d[o] = dy[(int64_t)r * out + o];
  } else {
    for (int i = 0; i <= s.layers; ++i)
      for (int j = 0; j < s.width[i]; ++j) a[i][j] = 0.f;
    for (int i = 0; i < s.layers; ++i)
      for (int j = 0; j < s.width[i + 1]; ++j) pre[i][j] = 0.f;
    for (int o = 0; o < out; ++o) d[o] = 0.f;
  }
  float* mine = slab + wid * l.total;
  for (int i = s.layers - 1; i >= 0; --i) {
    const int in = s.width[i], on = s.width[i + 1];
    for (int j = 0; j < in; ++j) dprev[j] = 0.f;
    for (int o = 0; o < on; ++o) {
      const float dp = live ? d[o] * act_df(s.act[i], pre[i][o], a[i + 1][o]) : 0.f;  // dL/d pre[i][o]
      const float sb = warp_sum(dp);
      if (lane == 0) mine[l.b[i] + o] = sb;
      for (int j = 0; j < in; ++j) {
        const float sw = warp_sum(dp * a[i][j]);
        if (lane == 0) mine[l.w[i] + o * in + j] = sw;
        dprev[j] = fmaf(sp[l.w[i] + o * in + j], dp, dprev[j]);
      }
    }
    for (int j = 0; j < in; ++j) d[j] = dprev[j];
  }
  __syncthreads();
  float* dst = partials + (int64_t)blockIdx.x * l.total;
  for (int q = threadIdx.x; q < l.total; q += kThreads) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += slab[w * l.total + q];
    dst[q] = t;
  }
}

__global__ void small_mlp_finish_kernel(v1t_mlp_spec s, v1t_mlp_ptrs g, const float* __restrict__ partials,
                                        int parts) {
  const Layout l = layout_of(s);
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= l.total) return;
  float t = 0.f;
  for (int c = 0; c < parts; ++c) t += partials[(int64_t)c * l.total + q];
  for (int i = 0; i < s.layers; ++i) {
    const int nw = s.width[i] * s.width[i + 1];
    if (q >= l.w[i] && q < l.w[i] + nw) {
      if (g.w[i]) g.w[i][q - l.w[i]] = t;
      return;
    }
    if (q >= l.b[i] && q < l.b[i] + s.width[i + 1]) {
      if (g.b[i]) g.b[i][q - l.b[i]] = t;
      return;
    }
  }
}

int check_spec(const v1t_mlp_spec* s, const v1t_mlp_ptrs* p) {
  V1T_CHECK_ARG(s && p, "small_mlp: null spec or parameter table");
  V1T_CHECK_ARG(s->rows >= 0 && s->layers >= 1 && s->layers <= kMaxL, "small_mlp: rows %d layers %d unsupported",
                s->rows, s->layers);
  for (int i = 0; i <= s->layers; ++i)
    V1T_CHECK_ARG(s->width[i] >= 1 && s->width[i] <= kMaxW, "small_mlp: width[%d] = %d outside 1..%d", i,
                  s->width[i], kMaxW);
  for (int i = 0; i < s->layers; ++i) {
    V1T_CHECK_ARG(s->act[i] >= V1T_ACT_NONE && s->act[i] <= V1T_ACT_ELU, "small_mlp: unknown activation %d",
                  s->act[i]);
    V1T_CHECK_ARG(p->w[i], "small_mlp: layer %d has no weight", i);
  }
  V1T_CHECK_ARG(s->x_ld >= s->width[0], "small_mlp: x_ld %lld < input width %d", (long long)s->x_ld, s->width[0]);
  return V1T_OK;
}

}  // namespace
}  // namespace v1t

using namespace v1t;

extern "C" size_t v1t_small_mlp_scratch_bytes(const v1t_mlp_spec* s) {
  if (!s || s->layers < 1 || s->layers > kMaxL) return 0;
  return sizeof(float) * (size_t)layout_of(*s).total * (size_t)cdiv(s->rows > 0 ? s->rows : 1, kThreads);
}

extern "C" int v1t_small_mlp_forward(const v1t_mlp_spec* s, const v1t_mlp_ptrs* params, const float* x, float* y,
                                     void* stream) {
  V1T_TRY(check_spec(s, params));
  if (s->rows == 0) return V1T_OK;
  V1T_CHECK_ARG(x && y, "small_mlp_forward: null tensor");
  const Layout l = layout_of(*s);
  small_mlp_forward_kernel<<<cdiv(s->rows, kThreads), kThreads, sizeof(float) * l.total, (cudaStream_t)stream>>>(
      *s, *params, x, y);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

extern "C" int v1t_small_mlp_backward(const v1t_mlp_spec* s, const v1t_mlp_ptrs* params, const float* x,
                                      const float* dy, const v1t_mlp_ptrs* grads, void* scratch, void* stream) {
  V1T_TRY(check_spec(s, params));
  V1T_CHECK_ARG(grads, "small_mlp_backward: null gradient table");
  V1T_CHECK_ARG(s->rows > 0, "small_mlp_backward: no rows");
  V1T_CHECK_ARG(x && dy && scratch, "small_mlp_backward: null tensor");
  const Layout l = layout_of(*s);
  const int parts = cdiv(s->rows, kThreads);
  const size_t smem = sizeof(float) * (size_t)l.total * (1 + kWarps);
  cudaStream_t st = (cudaStream_t)stream;
  if (smem > 48 * 1024)
    V1T_CUDA(cudaFuncSetAttribute(small_mlp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  small_mlp_backward_kernel<<<parts, kThreads, smem, st>>>(*s, *params, x, dy, (float*)scratch);
  V1T_LAUNCH_CHECK();
  small_mlp_finish_kernel<<<cdiv(l.total, 128), 128, 0, st>>>(*s, *grads, (const float*)scratch, parts);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
