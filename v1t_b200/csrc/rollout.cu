// Attention rollout of a batch of recorded attention maps — one HBM-bound pass per block.
//   reference: src/v1t/utils/attention_rollout.py:92-133 (attention_rollout / attention_rollouts):
//     M_n = max over heads of A[n];  Â_n = (M_n + I) / rowsum(M_n + I);  J_0 = Â_0;  J_n = Â_n J_{n-1};
//     heat = normalize(J_{L-1}[0, 1:] reshaped (gh, gw)) resized (bilinear, antialias off) to the image shape.
//   The reference runs, per SAMPLE in a Python loop, L-1 chained T x T x T matmuls and keeps only row 0 of the last
//   product.  Row 0 of a matrix product chain is a chain of vector-matrix products:
//     r_{L-1} = Â_{L-1}[0, :],   r_n = r_{n+1} Â_n  (n = L-2 .. 0),   heat = r_0[1:]
//   so every attention matrix is read ONCE (the last block's: only its first row), 2 T^2 flops per block instead of
//   2 T^3, and head-max, identity, row normalisation and the product are one fused pass:
//     r_n[j] = sum_i c_i (M_n[i,j] + delta_ij),   c_i = r_{n+1}[i] / (1 + sum_j M_n[i,j]).
//   Algorithmic bytes: B (L-1) H T^2 4  (2.1 GB at B=16, L=H=4, T=1654).
//
// Work decomposition: a warp owns a row i of one sample: 2*NQ columns per lane in registers (128-bit-friendly float2
// loads of the H head rows in flight together), warp-shuffle row sum, accumulation of c_i M[i,:] in registers over
// the warp's rows; the CTA's warps combine through shared memory, CTAs through per-CTA partials that the NEXT launch
// sums in a fixed order when it reads the row vector (deterministic, no atomics, no combine launch).
#include "common.cuh"

namespace v1t {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

template <int NQ, bool VEC>
__device__ __forceinline__ int col_of(int lane, int q, int e) {
  return VEC ? 2 * (lane + 32 * q) + e : lane + 32 * (2 * q + e);
}

// max over heads of row `row` of block n of sample b: m[q][e] (0 past T)
template <int NQ, bool VEC>
__device__ __forceinline__ void head_max_row(const float* __restrict__ blk /* [H,T,T] */, int H, int T, int row,
                                             int lane, float (&m)[NQ][2]) {
#pragma unroll
  for (int q = 0; q < NQ; ++q) m[q][0] = m[q][1] = 0.f;  // probabilities are >= 0
  for (int h = 0; h < H; ++h) {
    const float* src = blk + ((int64_t)h * T + row) * T;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if (VEC) {
        const int c = 2 * (lane + 32 * q);
        if (c < T) {  // T even: c + 1 < T as well
          const float2 v = __ldcs(reinterpret_cast<const float2*>(src + c));
          m[q][0] = fmaxf(m[q][0], v.x);
          m[q][1] = fmaxf(m[q][1], v.y);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = lane + 32 * (2 * q + e);
          if (c < T) m[q][e] = fmaxf(m[q][e], __ldcs(src + c));
        }
      }
    }
  }
}

// r_last[b][j] = Â_{L-1}[0, j]: one warp per sample
template <int NQ, bool VEC>
__global__ void __launch_bounds__(32) rollout_init_kernel(const float* __restrict__ attn, int L, int H, int T,
                                                          float* __restrict__ r_out) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const float* blk = attn + ((int64_t)b * L + (L - 1)) * H * (int64_t)T * T;
  float m[NQ][2];
  head_max_row<NQ, VEC>(blk, H, T, 0, lane, m);
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) s += m[q][0] + m[q][1];
  s = warp_sum(s) + 1.f;
  const float inv = 1.f / s;
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = col_of<NQ, VEC>(lane, q, e);
      if (c < T) r_out[(int64_t)b * T + c] = (m[q][e] + (c == 0 ? 1.f : 0.f)) * inv;
    }
}

// The row vector entering a step: either stored directly (from the init kernel) or still split into the previous
// step's per-CTA partials + identity term, summed here by the warp that needs the element (lanes stride over the
// partials, then a shuffle tree: fixed order) -- no combine launch between the blocks.
struct RowVec {
  const float* direct;   // [B][T] or null
  const float* partial;  // [B][G][T]
  const float* cvec;     // [B][T]
  int G;
};
__device__ __forceinline__ float row_value(const RowVec& r, int T, int b, int row, int lane) {
  if (r.direct) return r.direct[(int64_t)b * T + row];
  float t = 0.f;
  for (int g = lane; g < r.G; g += 32) t += r.partial[((int64_t)b * r.G + g) * T + row];
  return warp_sum(t) + r.cvec[(int64_t)b * T + row];
}

// grid (G, B).  partial [B][G][T]: this CTA's sum over its rows of c_i M[i,:];  cvec [B][T]: c_i (the identity term)
template <int NQ, bool VEC>
__global__ void __launch_bounds__(kThreads, 1) rollout_step_kernel(const float* __restrict__ attn, int L, int H,
                                                                   int T, int n, RowVec r_in,
                                                                   float* __restrict__ partial,
                                                                   float* __restrict__ cvec) {
  extern __shared__ float slab[];  // [kWarps][T]
  const int b = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float* blk = attn + ((int64_t)b * L + n) * H * (int64_t)T * T;
  float acc[NQ][2];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q][0] = acc[q][1] = 0.f;
  for (int row = blockIdx.x * kWarps + wid; row < T; row += gridDim.x * kWarps) {
    float m[NQ][2];
    head_max_row<NQ, VEC>(blk, H, T, row, lane, m);
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) s += m[q][0] + m[q][1];
    s = warp_sum(s) + 1.f;
    const float c = row_value(r_in, T, b, row, lane) / s;
    if (lane == 0) cvec[(int64_t)b * T + row] = c;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      acc[q][0] = fmaf(c, m[q][0], acc[q][0]);
      acc[q][1] = fmaf(c, m[q][1], acc[q][1]);
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = col_of<NQ, VEC>(lane, q, e);
      if (c < T) slab[wid * T + c] = acc[q][e];
    }
  __syncthreads();
  float* dst = partial + ((int64_t)b * gridDim.x + blockIdx.x) * T;
  for (int j = threadIdx.x; j < T; j += kThreads) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += slab[w * T + j];
    dst[j] = t;
  }
}

// heat = normalize(r[1:]) as (gh, gw), bilinear-resized (align_corners = false, no antialias) to (oh, ow); CTA per
// sample; the final row vector is assembled from the last step's partials into shared memory ([T - 1] floats)
__global__ void __launch_bounds__(256) rollout_heatmap_kernel(RowVec r, int T, int gh, int gw, int oh, int ow,
                                                              float* __restrict__ out) {
  extern __shared__ float h[];  // [gh * gw]
  __shared__ float rmin[8], rmax[8];
  const int b = blockIdx.x;
  const int n = gh * gw;
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v;
    if (r.direct) {
      v = r.direct[(int64_t)b * T + 1 + i];
    } else {
      v = 0.f;
      for (int g = 0; g < r.G; ++g) v += r.partial[((int64_t)b * r.G + g) * T + 1 + i];
      v += r.cvec[(int64_t)b * T + 1 + i];
    }
    h[i] = v;
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  hi = warp_max(hi);
  lo = -warp_max(-lo);
  if ((threadIdx.x & 31) == 0) {
    rmin[threadIdx.x >> 5] = lo;
    rmax[threadIdx.x >> 5] = hi;
  }
  __syncthreads();  // also publishes h[]
  lo = rmin[0];
  hi = rmax[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
    lo = fminf(lo, rmin[w]);
    hi = fmaxf(hi, rmax[w]);
  }
  const float range = hi - lo;  // the reference divides by (max - min) unguarded (attention_rollout.py:88-89)
  const float sy = (float)gh / (float)oh, sx = (float)gw / (float)ow;
  for (int i = threadIdx.x; i < oh * ow; i += blockDim.x) {
    const int oy = i / ow, ox = i % ow;
    const float fy = fmaxf(sy * ((float)oy + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)ox + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)fy, gh - 1), x0 = min((int)fx, gw - 1);
    const int y1 = min(y0 + 1, gh - 1), x1 = min(x0 + 1, gw - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float v00 = (h[y0 * gw + x0] - lo) / range, v01 = (h[y0 * gw + x1] - lo) / range;
    const float v10 = (h[y1 * gw + x0] - lo) / range, v11 = (h[y1 * gw + x1] - lo) / range;
    out[(int64_t)b * oh * ow + i] =
        (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

int rows_groups(int B) { return B >= kNumSMs ? 1 : (kNumSMs / B > 0 ? kNumSMs / B : 1); }

struct RolloutScratch {
  float *r0, *cvec[2], *partial[2];
  size_t total;
};
RolloutScratch carve(int B, int T, void* base) {
  RolloutScratch s;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* q = p ? p + off : nullptr;
    off += (size_t)round_up((int64_t)bytes, 256);
    return (float*)q;
  };
  s.r0 = take(sizeof(float) * (size_t)B * T);
  for (int i = 0; i < 2; ++i) {  // double-buffered: step n reads step n+1's partials while writing its own
    s.cvec[i] = take(sizeof(float) * (size_t)B * T);
    s.partial[i] = take(sizeof(float) * (size_t)B * rows_groups(B) * T);
  }
  s.total = off;
  return s;
}

template <int NQ, bool VEC>
int run(const float* attn, int B, int L, int H, int T, int gh, int gw, int oh, int ow, float* heat, void* scratch,
        cudaStream_t st) {
  RolloutScratch ws = carve(B, T, scratch);
  const int G = rows_groups(B);
  rollout_init_kernel<NQ, VEC><<<B, 32, 0, st>>>(attn, L, H, T, ws.r0);
  V1T_LAUNCH_CHECK();
  RowVec cur{ws.r0, nullptr, nullptr, 0};
  const size_t smem = sizeof(float) * (size_t)kWarps * T;
  if (L > 1 && smem > 48 * 1024)
    V1T_CUDA(cudaFuncSetAttribute(rollout_step_kernel<NQ, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  int buf = 0;
  for (int n = L - 2; n >= 0; --n, buf ^= 1) {
    rollout_step_kernel<NQ, VEC><<<dim3(G, B), kThreads, smem, st>>>(attn, L, H, T, n, cur, ws.partial[buf],
                                                                    ws.cvec[buf]);
    V1T_LAUNCH_CHECK();
    cur = RowVec{nullptr, ws.partial[buf], ws.cvec[buf], G};
  }
  rollout_heatmap_kernel<<<B, 256, sizeof(float) * (size_t)(T - 1), st>>>(cur, T, gh, gw, oh, ow, heat);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

}  // namespace
}  // namespace v1t

using namespace v1t;

extern "C" size_t v1t_rollout_scratch_bytes(int B, int T) {
  if (B <= 0 || T <= 0) return 0;
  return carve(B, T, nullptr).total;
}

extern "C" int v1t_attention_rollout(const float* attn, int B, int L, int H, int T, int gh, int gw, int out_h,
                                     int out_w, float* heatmaps, void* scratch, void* stream) {
  V1T_CHECK_ARG(attn && heatmaps && scratch, "attention_rollout: null tensor");
  V1T_CHECK_ARG(B > 0 && L > 0 && H > 0 && T > 1, "attention_rollout: bad shape B %d L %d H %d T %d", B, L, H, T);
  V1T_CHECK_ARG(gh > 0 && gw > 0 && gh * gw == T - 1, "attention_rollout: grid %d x %d is not the %d patches", gh, gw,
                T - 1);
  V1T_CHECK_ARG(out_h > 0 && out_w > 0, "attention_rollout: bad output shape");
  V1T_CHECK_ARG(T <= 2304, "attention_rollout: %d tokens > 2304 unsupported", T);
  V1T_CHECK_ARG(B <= 65535, "attention_rollout: batch %d > 65535", B);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (T % 2 == 0) && (((uintptr_t)attn & 7u) == 0);
#define V1T_ROLL(NQ)                                                                                   \
  return vec ? run<NQ, true>(attn, B, L, H, T, gh, gw, out_h, out_w, heatmaps, scratch, st)             \
             : run<NQ, false>(attn, B, L, H, T, gh, gw, out_h, out_w, heatmaps, scratch, st)
  if (T <= 512) { V1T_ROLL(8); }
  if (T <= 1024) { V1T_ROLL(16); }
  if (T <= 1664) { V1T_ROLL(26); }
  V1T_ROLL(36);
#undef V1T_ROLL
}
