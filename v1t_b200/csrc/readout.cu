// Gaussian2d readout (+ ELU1 + Poisson) forward / backward — bandwidth-bound gather/reduce kernels.
//   reference: gaussian2d.py:195-278 (sample_grid, grid_sample bilinear/zeros/align_corners=True, feature dot,
//              bias), models/utils.py:109-118 (ELU1), losses.py:114-119,153-166 (PoissonLoss + scale_ds).
//
// Layout: the core map is channel-last (each pixel's C channels contiguous, fs_x stride between pixels), so a
// warp reads a pixel's channel vector with coalesced 128 B requests.  Work decomposition is NEURON-MAJOR:
// a CTA owns 32 consecutive neurons (their feature columns staged once through shared memory, transposed with
// a +1 pad so both the coalesced global read and the per-neuron column read are conflict-free), one warp walks
// one neuron over the batch.  Everything reduced over the batch (d_features, d_bias, d_mu, d_sigma) is
// accumulated in registers by the warp that owns the neuron: no atomics, deterministic.  Reductions over
// neurons (loss, d_shifts) go through per-CTA partials summed in a fixed order.  Only the scatter into d_fmap
// uses fp32 reductions (red.global.add), 128 B coalesced per request.
#include "common.cuh"
#include "kernels.cuh"
#include <stdlib.h>

#include <algorithm>

namespace v1t {
namespace {

constexpr int kNeuronsPerCta = 32;
constexpr int kWarps = 32;     // backward: 1 neuron per warp, one 1024-thread CTA per SM (64 registers per thread)
constexpr int kFwdWarps = 32;  // forward: 1 neuron per warp, 54 warps/SM at N=8000 (58 registers per thread)
constexpr int kBatchTile = 32;  // samples staged per output tile
constexpr float kEpsF32 = 1.1920928955078125e-07f;  // torch.finfo(float32).eps (losses.py:22)

struct Corner {
  int64_t off[4];  // element offset of the pixel's channel vector (valid corners only)
  float w[4];      // bilinear weight (0 when out of bounds)
  float wx[4], wy[4];
  float valid[4];
};

struct GridPos {
  float pre_x, pre_y;  // before clamp
  float ix, iy;
};

__device__ __forceinline__ GridPos grid_position(const float* __restrict__ mu, const float* __restrict__ sigma,
                                                 const float* __restrict__ noise, const float* __restrict__ shifts,
                                                 int b, int n, int N, int gh, int gw) {
  GridPos g;
  float px = mu[2 * n], py = mu[2 * n + 1];
  if (noise) {
    const float n0 = noise[((int64_t)b * N + n) * 2], n1 = noise[((int64_t)b * N + n) * 2 + 1];
    px += sigma[4 * n + 0] * n0 + sigma[4 * n + 1] * n1;  // einsum("ancd,bnid->bnic")
    py += sigma[4 * n + 2] * n0 + sigma[4 * n + 3] * n1;
  }
  g.pre_x = px;
  g.pre_y = py;
  float gx = fminf(fmaxf(px, -1.f), 1.f), gy = fminf(fmaxf(py, -1.f), 1.f);
  if (shifts) {
    gx += shifts[2 * b];
    gy += shifts[2 * b + 1];
  }
  g.ix = (gx + 1.f) * 0.5f * (float)(gw - 1);
  g.iy = (gy + 1.f) * 0.5f * (float)(gh - 1);
  return g;
}

__device__ __forceinline__ void make_corners(const GridPos& g, int gh, int gw, int64_t fs_y, int64_t fs_x,
                                             Corner& c) {
  const float ixc = fminf(fmaxf(g.ix, -2.f), (float)gw + 1.f);
  const float iyc = fminf(fmaxf(g.iy, -2.f), (float)gh + 1.f);
  const float fx0 = floorf(ixc), fy0 = floorf(iyc);
  const int x0 = (int)fx0, y0 = (int)fy0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dx = k & 1, dy = k >> 1;
    const int xc = x0 + dx, yc = y0 + dy;
    const bool ok = (xc >= 0) && (xc <= gw - 1) && (yc >= 0) && (yc <= gh - 1);
    const float wx = 1.f - fabsf(g.ix - (float)xc), wy = 1.f - fabsf(g.iy - (float)yc);
    c.valid[k] = ok ? 1.f : 0.f;
    c.wx[k] = wx;
    c.wy[k] = wy;
    c.w[k] = ok ? wx * wy : 0.f;
    c.off[k] = ok ? (int64_t)yc * fs_y + (int64_t)xc * fs_x : 0;
  }
}

__device__ __forceinline__ float elu1(float z) { return (z > 0.f ? z : expm1f(z)) + 1.f; }

// stage features[:, n0:n0+32] -> fs[c][33]
__device__ __forceinline__ void stage_features(const float* __restrict__ features, float* fs, int C, int N, int n0) {
  for (int i = threadIdx.x; i < C * kNeuronsPerCta; i += blockDim.x) {
    const int c = i / kNeuronsPerCta, j = i % kNeuronsPerCta;
    fs[c * 33 + j] = (n0 + j < N) ? __ldg(features + (int64_t)c * N + n0 + j) : 0.f;
  }
}

// grid (ceil(N/32), batch tiles); block 1024: one warp per neuron (the walk over the batch is a chain of dependent
// L2 latencies, so the forward wants as many warps in flight as the SM holds: 2 CTAs x 32 warps).  smem: fs [C][33] + zt [kBatchTile][33]
template <int NV>
__global__ void __launch_bounds__(kFwdWarps * 32) readout_forward_kernel(
    v1t_readout_shape s, const float* __restrict__ fmap, const float* __restrict__ mu,
    const float* __restrict__ sigma, const float* __restrict__ noise, const float* __restrict__ shifts,
    const float* __restrict__ features, const float* __restrict__ bias, const float* __restrict__ y_true,
    float* __restrict__ z_out, float* __restrict__ y_out, float* __restrict__ loss_partials) {
  extern __shared__ float smem[];
  float* fs = smem;                                  // [C][33]
  float* zt = smem + (size_t)s.channels * 33;        // [kBatchTile][33]
  __shared__ float red[kFwdWarps];
  const int C = s.channels, N = s.neurons, B = s.batch;
  const int n0 = blockIdx.x * kNeuronsPerCta;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b_begin = blockIdx.y * kBatchTile, b_end = min(B, b_begin + kBatchTile);

  stage_features(features, fs, C, N, n0);
  __syncthreads();

  for (int j = wid; j < kNeuronsPerCta; j += kFwdWarps) {
    const int n = n0 + j;
    if (n >= N) break;  // warp-uniform
    float f[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      f[i] = c < C ? fs[c * 33 + j] : 0.f;
    }
    const float bn = bias ? __ldg(bias + n) : 0.f;
    for (int b = b_begin; b < b_end; ++b) {
      const GridPos g = grid_position(mu, sigma, noise, shifts, b, n, N, s.gh, s.gw);
      Corner cr;
      make_corners(g, s.gh, s.gw, s.fs_y, s.fs_x, cr);
      const float* base = fmap + (int64_t)b * s.fs_b;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (cr.w[k] != 0.f) {  // warp-uniform
          const float* px = base + cr.off[k];
          float t = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            if (c < C) t = fmaf(__ldg(px + c), f[i], t);
          }
          acc = fmaf(cr.w[k], t, acc);
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) zt[(b - b_begin) * 33 + j] = acc + bn;
    }
  }
  __syncthreads();

  // coalesced epilogue over the [batch tile][32 neurons] tile: z, y = elu(z)+1, Poisson partial
  float lsum = 0.f;
  const int n = n0 + lane;
  for (int bl = wid; bl < b_end - b_begin; bl += kFwdWarps) {
    if (n < N) {
      const int64_t o = (int64_t)(b_begin + bl) * N + n;
      const float z = zt[bl * 33 + lane];
      z_out[o] = z;
      const float y = elu1(z);
      if (y_out) y_out[o] = y;
      if (loss_partials) {
        const float yp = y + kEpsF32, yt = __ldg(y_true + o) + kEpsF32;
        lsum += yp - yt * logf(yp);
      }
    }
  }
  if (loss_partials) {
    lsum = warp_sum(lsum);
    if (lane == 0) red[wid] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kFwdWarps; ++w) t += red[w];
      loss_partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
  }
}

// ---- 128-bit backward ---------------------------------------------------------------------------------
// When the map's strides and base are 16-byte aligned (the core emits rows of emb_ld = 160 floats) a lane of the
// backward owns the 4-channel chunks lane, lane + 32, ...: one LDG.128 per chunk and one red.global.add.v4.f32
// instead of four scalar reductions (79 M scalar ones per pass at B=16, N=8000).  A chunk that straddles C
// (155 = 38 chunks + 3) is handled element-wise.  NC = ceil(C / 128) chunk rounds per lane.  Measured on the bench
// workload: backward 1.89 -> 1.69 ms per step; the same layout in the FORWARD was slower (0.48 -> 0.78 ms: 25 of 32
// lanes idle in the second round against 155/160 in the scalar layout) and is not used.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int NC>
__device__ __forceinline__ void load_feature_chunks(const float* fs, int C, int j, int lane, float (&f)[NC][4]) {
#pragma unroll
  for (int i = 0; i < NC; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = 4 * (lane + 32 * i) + e;
      f[i][e] = c < C ? fs[c * 33 + j] : 0.f;
    }
}

__global__ void sum_scale_kernel(const float* __restrict__ partials, int n, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];  // fixed order per thread
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) *out = t * scale;
  }
}

// backward.  grid (ceil(N/32), batch tiles).  Per-(CTA) outputs that still need a reduction over batch tiles
// (d_features, d_bias, d_mu, d_sigma) are written to partial slabs indexed by blockIdx.y when gridDim.y > 1.
template <int NV>
__global__ void __launch_bounds__(kWarps * 32) readout_backward_kernel(
    v1t_readout_shape s, const float* __restrict__ fmap, const float* __restrict__ mu,
    const float* __restrict__ sigma, const float* __restrict__ noise, const float* __restrict__ shifts,
    const float* __restrict__ features, const float* __restrict__ z_in, const float* __restrict__ dz_in,
    const float* __restrict__ y_true, float gscale, float* __restrict__ d_fmap, float* __restrict__ d_feat_part,
    float* __restrict__ d_small_part /* [tiles][N][7]: bias, mu(2), sigma(4) */,
    float* __restrict__ d_shift_part /* [gridDim.x][B][2] */) {
  extern __shared__ float smem[];
  float* fs = smem;                                         // [C][33] features, later d_features
  float* sh = smem + (size_t)s.channels * 33;               // [kBatchTile][32][2] d_grid per (b, neuron)
  const int C = s.channels, N = s.neurons, B = s.batch;
  const int n0 = blockIdx.x * kNeuronsPerCta;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b_begin = blockIdx.y * kBatchTile, b_end = min(B, b_begin + kBatchTile);
  const float half_w = 0.5f * (float)(s.gw - 1), half_h = 0.5f * (float)(s.gh - 1);

  stage_features(features, fs, C, N, n0);
  for (int i = threadIdx.x; i < kBatchTile * 64; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();

  float df_keep[(kNeuronsPerCta / kWarps)][NV];  // d_features of this warp's neurons
#pragma unroll
  for (int q = 0; q < kNeuronsPerCta / kWarps; ++q) {
    const int j = wid + q * kWarps;
    const int n = n0 + j;
    float f[NV], df[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      f[i] = c < C ? fs[c * 33 + j] : 0.f;
      df[i] = 0.f;
    }
    float dbias = 0.f, dmx = 0.f, dmy = 0.f, ds0 = 0.f, ds1 = 0.f, ds2 = 0.f, ds3 = 0.f;
    if (n < N) {
      for (int b = b_begin; b < b_end; ++b) {
        const int64_t o = (int64_t)b * N + n;
        float g;
        if (dz_in) {
          g = __ldg(dz_in + o);
        } else {  // fused ELU1 + Poisson gradient
          const float z = __ldg(z_in + o);
          const float y = elu1(z);
          g = gscale * (1.f - (__ldg(y_true + o) + kEpsF32) / (y + kEpsF32)) * (z > 0.f ? 1.f : expf(z));
        }
        const GridPos gp = grid_position(mu, sigma, noise, shifts, b, n, N, s.gh, s.gw);
        Corner cr;
        make_corners(gp, s.gh, s.gw, s.fs_y, s.fs_x, cr);
        const float* base = fmap + (int64_t)b * s.fs_b;
        float* dbase = d_fmap ? d_fmap + (int64_t)b * s.fs_b : nullptr;
        float dotk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float t = 0.f;
          if (cr.valid[k] != 0.f) {  // warp-uniform
            const float* px = base + cr.off[k];
            const float gw_k = g * cr.w[k];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
              const int c = lane + 32 * i;
              if (c < C) {
                const float v = __ldg(px + c);
                t = fmaf(v, f[i], t);
                df[i] = fmaf(gw_k, v, df[i]);
                if (dbase) atomicAdd(dbase + cr.off[k] + c, gw_k * f[i]);
              }
            }
          }
          dotk[k] = t;
        }
        // reduce the 4 corner dots across the warp
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) dotk[k] += __shfl_xor_sync(0xffffffffu, dotk[k], o2);
        }
        float gx = 0.f, gy = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float sx = (k & 1) ? 1.f : -1.f, sy = (k >> 1) ? 1.f : -1.f;
          gx += cr.valid[k] * sx * cr.wy[k] * dotk[k];
          gy += cr.valid[k] * sy * cr.wx[k] * dotk[k];
        }
        gx *= g * half_w;
        gy *= g * half_h;
        if (lane == 0) {
          sh[((b - b_begin) * 32 + j) * 2] = gx;
          sh[((b - b_begin) * 32 + j) * 2 + 1] = gy;
        }
        const float mx = (gp.pre_x >= -1.f && gp.pre_x <= 1.f) ? gx : 0.f;  // clamp backward
        const float my = (gp.pre_y >= -1.f && gp.pre_y <= 1.f) ? gy : 0.f;
        dbias += g;
        dmx += mx;
        dmy += my;
        if (noise) {
          const float q0 = noise[o * 2], q1 = noise[o * 2 + 1];
          ds0 = fmaf(mx, q0, ds0);
          ds1 = fmaf(mx, q1, ds1);
          ds2 = fmaf(my, q0, ds2);
          ds3 = fmaf(my, q1, ds3);
        }
      }
      if (lane == 0 && d_small_part) {
        float* dst = d_small_part + ((int64_t)blockIdx.y * N + n) * 7;
        dst[0] = dbias; dst[1] = dmx; dst[2] = dmy; dst[3] = ds0; dst[4] = ds1; dst[5] = ds2; dst[6] = ds3;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) df_keep[q][i] = df[i];
  }
  __syncthreads();  // all warps done reading fs as features
  if (d_feat_part) {
#pragma unroll
    for (int q = 0; q < kNeuronsPerCta / kWarps; ++q) {
      const int j = wid + q * kWarps;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < C) fs[c * 33 + j] = df_keep[q][i];
      }
    }
    __syncthreads();
    float* dst = d_feat_part + (int64_t)blockIdx.y * C * N;
    for (int i = threadIdx.x; i < C * kNeuronsPerCta; i += blockDim.x) {
      const int c = i / kNeuronsPerCta, j = i % kNeuronsPerCta;
      if (n0 + j < N) dst[(int64_t)c * N + n0 + j] = fs[c * 33 + j];
    }
  }
  // d_shifts partial of this CTA: sum over its 32 neurons, fixed order
  if (d_shift_part) {
    for (int i = threadIdx.x; i < (b_end - b_begin) * 2; i += blockDim.x) {
      const int bl = i >> 1, xy = i & 1;
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) t += sh[(bl * 32 + j) * 2 + xy];
      d_shift_part[((int64_t)blockIdx.x * B + b_begin + bl) * 2 + xy] = t;
    }
  }
}

// 128-bit variant of readout_backward_kernel (same grid, smem and outputs)
template <int NC>
__global__ void __launch_bounds__(kWarps * 32) readout_backward_v4_kernel(
    v1t_readout_shape s, const float* __restrict__ fmap, const float* __restrict__ mu,
    const float* __restrict__ sigma, const float* __restrict__ noise, const float* __restrict__ shifts,
    const float* __restrict__ features, const float* __restrict__ z_in, const float* __restrict__ dz_in,
    const float* __restrict__ y_true, float gscale, float* __restrict__ d_fmap, float* __restrict__ d_feat_part,
    float* __restrict__ d_small_part, float* __restrict__ d_shift_part) {
  extern __shared__ float smem[];
  float* fs = smem;                                         // [C][33] features, later d_features
  float* sh = smem + (size_t)s.channels * 33;               // [kBatchTile][32][2] d_grid per (b, neuron)
  const int C = s.channels, N = s.neurons, B = s.batch;
  const int n0 = blockIdx.x * kNeuronsPerCta;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b_begin = blockIdx.y * kBatchTile, b_end = min(B, b_begin + kBatchTile);
  const float half_w = 0.5f * (float)(s.gw - 1), half_h = 0.5f * (float)(s.gh - 1);

  stage_features(features, fs, C, N, n0);
  for (int i = threadIdx.x; i < kBatchTile * 64; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();

  float df_keep[(kNeuronsPerCta / kWarps)][NC][4];  // d_features of this warp's neurons
#pragma unroll
  for (int q = 0; q < kNeuronsPerCta / kWarps; ++q) {
    const int j = wid + q * kWarps;
    const int n = n0 + j;
    float f[NC][4], df[NC][4];
    load_feature_chunks<NC>(fs, C, j, lane, f);
    // PACKED TAIL (128 < C <= 160, e.g. 155): the second pass over the channels holds at most 8 four-channel chunks, so the
    // four corners share ONE pass -- lane group g = lane >> 3 takes corner g, chunk 32 + (lane & 7) -- instead of four
    // passes with 7 active lanes each (19 % fewer warp instructions in this kernel)
    const bool packed = NC == 2 && C > 128 && C <= 160;
    const int pg = lane >> 3, pc = 4 * (32 + (lane & 7));
    if (packed) {
#pragma unroll
      for (int e = 0; e < 4; ++e) f[NC - 1][e] = (pc + e < C) ? fs[(pc + e) * 33 + j] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NC; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) df[i][e] = 0.f;
    float dbias = 0.f, dmx = 0.f, dmy = 0.f, ds0 = 0.f, ds1 = 0.f, ds2 = 0.f, ds3 = 0.f;
    if (n < N) {
      for (int b = b_begin; b < b_end; ++b) {
        const int64_t o = (int64_t)b * N + n;
        float g;
        if (dz_in) {
          g = __ldg(dz_in + o);
        } else {  // fused ELU1 + Poisson gradient
          const float z = __ldg(z_in + o);
          const float y = elu1(z);
          g = gscale * (1.f - (__ldg(y_true + o) + kEpsF32) / (y + kEpsF32)) * (z > 0.f ? 1.f : expf(z));
        }
        const GridPos gp = grid_position(mu, sigma, noise, shifts, b, n, N, s.gh, s.gw);
        Corner cr;
        make_corners(gp, s.gh, s.gw, s.fs_y, s.fs_x, cr);
        const float* base = fmap + (int64_t)b * s.fs_b;
        float* dbase = d_fmap ? d_fmap + (int64_t)b * s.fs_b : nullptr;
        float dotk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float t = 0.f;
          if (cr.valid[k] != 0.f) {  // warp-uniform
            const float* px = base + cr.off[k];
            const float gw_k = g * cr.w[k];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
              if (packed && i == NC - 1) break;  // handled for all four corners at once below
              const int c = 4 * (lane + 32 * i);
              if (c + 3 < C) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(px + c));
                t = fmaf(v.x, f[i][0], t);
                t = fmaf(v.y, f[i][1], t);
                t = fmaf(v.z, f[i][2], t);
                t = fmaf(v.w, f[i][3], t);
                df[i][0] = fmaf(gw_k, v.x, df[i][0]);
                df[i][1] = fmaf(gw_k, v.y, df[i][1]);
                df[i][2] = fmaf(gw_k, v.z, df[i][2]);
                df[i][3] = fmaf(gw_k, v.w, df[i][3]);
                if (dbase)
                  red_add_v4(dbase + cr.off[k] + c, gw_k * f[i][0], gw_k * f[i][1], gw_k * f[i][2], gw_k * f[i][3]);
              } else if (c < C) {
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                  if (c + e < C) {
                    const float v = __ldg(px + c + e);
                    t = fmaf(v, f[i][e], t);
                    df[i][e] = fmaf(gw_k, v, df[i][e]);
                    if (dbase) atomicAdd(dbase + cr.off[k] + c + e, gw_k * f[i][e]);
                  }
                }
              }
            }
          }
          dotk[k] = t;
        }
        if (packed) {
          constexpr int i = NC - 1;
          const float vg = pg == 0 ? cr.valid[0] : pg == 1 ? cr.valid[1] : pg == 2 ? cr.valid[2] : cr.valid[3];
          float t = 0.f;
          if (vg != 0.f && pc < C) {
            const int64_t og = pg == 0 ? cr.off[0] : pg == 1 ? cr.off[1] : pg == 2 ? cr.off[2] : cr.off[3];
            const float gw_k = g * (pg == 0 ? cr.w[0] : pg == 1 ? cr.w[1] : pg == 2 ? cr.w[2] : cr.w[3]);
            const float* px = base + og;
            if (pc + 3 < C) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(px + pc));
              t = fmaf(v.x, f[i][0], t);
              t = fmaf(v.y, f[i][1], t);
              t = fmaf(v.z, f[i][2], t);
              t = fmaf(v.w, f[i][3], t);
              df[i][0] = fmaf(gw_k, v.x, df[i][0]);
              df[i][1] = fmaf(gw_k, v.y, df[i][1]);
              df[i][2] = fmaf(gw_k, v.z, df[i][2]);
              df[i][3] = fmaf(gw_k, v.w, df[i][3]);
              if (dbase) red_add_v4(dbase + og + pc, gw_k * f[i][0], gw_k * f[i][1], gw_k * f[i][2], gw_k * f[i][3]);
            } else {
#pragma unroll
              for (int e = 0; e < 3; ++e) {
                if (pc + e < C) {
                  const float v = __ldg(px + pc + e);
                  t = fmaf(v, f[i][e], t);
                  df[i][e] = fmaf(gw_k, v, df[i][e]);
                  if (dbase) atomicAdd(dbase + og + pc + e, gw_k * f[i][e]);
                }
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) dotk[k] += (k == pg) ? t : 0.f;
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) dotk[k] += __shfl_xor_sync(0xffffffffu, dotk[k], o2);
        }
        float gx = 0.f, gy = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float sx = (k & 1) ? 1.f : -1.f, sy = (k >> 1) ? 1.f : -1.f;
          gx += cr.valid[k] * sx * cr.wy[k] * dotk[k];
          gy += cr.valid[k] * sy * cr.wx[k] * dotk[k];
        }
        gx *= g * half_w;
        gy *= g * half_h;
        if (lane == 0) {
          sh[((b - b_begin) * 32 + j) * 2] = gx;
          sh[((b - b_begin) * 32 + j) * 2 + 1] = gy;
        }
        const float mx = (gp.pre_x >= -1.f && gp.pre_x <= 1.f) ? gx : 0.f;  // clamp backward
        const float my = (gp.pre_y >= -1.f && gp.pre_y <= 1.f) ? gy : 0.f;
        dbias += g;
        dmx += mx;
        dmy += my;
        if (noise) {
          const float q0 = noise[o * 2], q1 = noise[o * 2 + 1];
          ds0 = fmaf(mx, q0, ds0);
          ds1 = fmaf(mx, q1, ds1);
          ds2 = fmaf(my, q0, ds2);
          ds3 = fmaf(my, q1, ds3);
        }
      }
      if (lane == 0 && d_small_part) {
        float* dst = d_small_part + ((int64_t)blockIdx.y * N + n) * 7;
        dst[0] = dbias; dst[1] = dmx; dst[2] = dmy; dst[3] = ds0; dst[4] = ds1; dst[5] = ds2; dst[6] = ds3;
      }
    }
    if (packed) {  // chunk 32 + (lane & 7): sum the four corner groups (fixed order), lanes 0-7 then hold the totals
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = df[NC - 1][e];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        df[NC - 1][e] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < NC; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) df_keep[q][i][e] = df[i][e];
  }
  __syncthreads();  // all warps done reading fs as features
  if (d_feat_part) {
#pragma unroll
    for (int q = 0; q < kNeuronsPerCta / kWarps; ++q) {
      const int j = wid + q * kWarps;
#pragma unroll
      for (int i = 0; i < NC; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = 4 * (lane + 32 * i) + e;
          if (c < C) fs[c * 33 + j] = df_keep[q][i][e];
        }
    }
    __syncthreads();
    float* dst = d_feat_part + (int64_t)blockIdx.y * C * N;
    for (int i = threadIdx.x; i < C * kNeuronsPerCta; i += blockDim.x) {
      const int c = i / kNeuronsPerCta, j = i % kNeuronsPerCta;
      if (n0 + j < N) dst[(int64_t)c * N + n0 + j] = fs[c * 33 + j];
    }
  }
  if (d_shift_part) {
    for (int i = threadIdx.x; i < (b_end - b_begin) * 2; i += blockDim.x) {
      const int bl = i >> 1, xy = i & 1;
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) t += sh[(bl * 32 + j) * 2 + xy];
      d_shift_part[((int64_t)blockIdx.x * B + b_begin + bl) * 2 + xy] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// d_fmap without atomics, bitwise reproducible (pixel-major): d_fmap[b, pix, :] += sum over the (neuron, corner) pairs that
// touch pixel `pix` of sample b of  dz[b,n] * w_corner * features[:, n].
//   1. readout_sort_kernel (one CTA per sample): a counting sort of the sample's <= 4 N (neuron, corner) entries by pixel
//      with a FIXED placement order -- warp w owns a contiguous range of neurons and its own row of per-pixel counters, so
//      the slot of an entry is  offset[pix] + (entries of warps < w on pix) + (earlier entries of warp w on pix), all of
//      which are independent of scheduling;
//   2. transpose_features_kernel: features [C, N] -> [N, Cq] so that a neuron's feature vector is one contiguous row;
//   3. readout_gather_kernel: one warp per (sample, pixel) walks its segment in order and adds coef * featT[n, :].
// The red.global.add scatter of the neuron-major kernel (B N 4 C reductions, serialised on the few pixels all neurons sit on
// at the start of training) is not issued then (its d_fmap argument is NULL).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSortWarps = 32;

// gradient of the loss w.r.t. the pre-activation z[b, n]: given, or the fused ELU1 + Poisson gradient
__device__ __forceinline__ float readout_dz(const float* __restrict__ dz_in, const float* __restrict__ z_in,
                                            const float* __restrict__ y_true, float gscale, int64_t o) {
  if (dz_in) return __ldg(dz_in + o);
  const float z = __ldg(z_in + o);
  const float y = elu1(z);
  return gscale * (1.f - (__ldg(y_true + o) + kEpsF32) / (y + kEpsF32)) * (z > 0.f ? 1.f : expf(z));
}

// pixel index (y * gw + x) and bilinear weight of the 4 corners of (b, n); pix = -1 for corners outside the map
__device__ __forceinline__ void corner_pixels(const v1t_readout_shape& s, const float* __restrict__ mu,
                                              const float* __restrict__ sigma, const float* __restrict__ noise,
                                              const float* __restrict__ shifts, int b, int n, int (&pix)[4], float (&w)[4]) {
  const GridPos g = grid_position(mu, sigma, noise, shifts, b, n, s.neurons, s.gh, s.gw);
  const float ixc = fminf(fmaxf(g.ix, -2.f), (float)s.gw + 1.f);
  const float iyc = fminf(fmaxf(g.iy, -2.f), (float)s.gh + 1.f);
  const int x0 = (int)floorf(ixc), y0 = (int)floorf(iyc);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xc = x0 + (k & 1), yc = y0 + (k >> 1);
    const bool ok = (xc >= 0) && (xc <= s.gw - 1) && (yc >= 0) && (yc <= s.gh - 1);
    pix[k] = ok ? yc * s.gw + xc : -1;
    w[k] = (1.f - fabsf(g.ix - (float)xc)) * (1.f - fabsf(g.iy - (float)yc));
  }
}

// smem: cnt [kSortWarps][ceil(L/2)] packed 16-bit counters per (warp, pixel) | off [L + 1]
__global__ void __launch_bounds__(kSortWarps * 32) readout_sort_kernel(
    v1t_readout_shape s, const float* __restrict__ mu, const float* __restrict__ sigma, const float* __restrict__ noise,
    const float* __restrict__ shifts, const float* __restrict__ z_in, const float* __restrict__ dz_in,
    const float* __restrict__ y_true, float gscale, int* __restrict__ seg_off, int* __restrict__ sorted_n,
    float* __restrict__ sorted_coef) {
  extern __shared__ uint32_t sm_sort[];
  const int L = s.gh * s.gw, Lw = (L + 1) / 2, N = s.neurons, b = blockIdx.x;
  uint32_t* cnt = sm_sort;                                   // [kSortWarps][Lw]
  int* off = reinterpret_cast<int*>(sm_sort + kSortWarps * Lw);  // [L + 1]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int npw = ((N + kSortWarps - 1) / kSortWarps + 31) & ~31;  // neurons per warp: whole 32-neuron steps
  const int n_begin = wid * npw, n_end = min(N, n_begin + npw);
  for (int i = threadIdx.x; i < kSortWarps * Lw; i += blockDim.x) cnt[i] = 0u;
  __syncthreads();
  uint32_t* mine = cnt + wid * Lw;
  // ---- pass A: entries of this warp's neurons per pixel
  for (int n = n_begin + lane; n < n_end; n += 32) {
    int pix[4];
    float w[4];
    corner_pixels(s, mu, sigma, noise, shifts, b, n, pix, w);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (pix[k] >= 0) atomicAdd(&mine[pix[k] >> 1], (pix[k] & 1) ? 65536u : 1u);
  }
  __syncthreads();
  // ---- per pixel: exclusive prefix over the warps (in place), total -> off, then an exclusive scan over the pixels
  for (int wd = threadIdx.x; wd < Lw; wd += blockDim.x) {  // one thread per counter word = pixels 2 wd, 2 wd + 1
    uint32_t run0 = 0, run1 = 0;
    for (int w2 = 0; w2 < kSortWarps; ++w2) {
      uint32_t* word = cnt + w2 * Lw + wd;
      const uint32_t c = *word;
      *word = run0 | (run1 << 16);
      run0 += c & 0xffffu;
      run1 += c >> 16;
    }
    off[2 * wd] = (int)run0;
    if (2 * wd + 1 < L) off[2 * wd + 1] = (int)run1;
  }
  __syncthreads();
  {  // block-wide exclusive scan of off[0..L) (L is a few thousand at most: serial per thread, then over the threads)
    __shared__ int part[kSortWarps * 32];
    const int per = (L + blockDim.x - 1) / blockDim.x;
    const int lo = threadIdx.x * per, hi = min(L, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += off[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      int run = 0;
      for (int i = 0; i < (int)blockDim.x; ++i) {
        const int v = part[i];
        part[i] = run;
        run += v;
      }
      off[L] = run;
    }
    __syncthreads();
    int run = part[threadIdx.x];
    for (int i = lo; i < hi; ++i) {
      const int v = off[i];
      off[i] = run;
      run += v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= L; i += blockDim.x) seg_off[(int64_t)b * (L + 1) + i] = off[i];
  // ---- pass B: placement.  Within the warp: neuron steps in order, corner 0..3, lanes in order.
  int* out_n = sorted_n + (int64_t)b * 4 * N;
  float* out_c = sorted_coef + (int64_t)b * 4 * N;
  for (int n0 = n_begin; n0 < n_end; n0 += 32) {
    const int n = n0 + lane;
    int pix[4] = {-1, -1, -1, -1};
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    float g = 0.f;
    if (n < n_end) {
      corner_pixels(s, mu, sigma, noise, shifts, b, n, pix, w);
      g = readout_dz(dz_in, z_in, y_true, gscale, (int64_t)b * N + n);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool valid = pix[k] >= 0;
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      if (valid) {
        const uint32_t same = __match_any_sync(vmask, pix[k]);
        const int rank = __popc(same & ((1u << lane) - 1u));
        const int leader = __ffs(same) - 1;
        const int sh = (pix[k] & 1) * 16;
        uint32_t cur = 0;
        if (lane == leader) cur = atomicAdd(&mine[pix[k] >> 1], (uint32_t)__popc(same) << sh);  // returns the old word
        cur = (__shfl_sync(same, cur, leader) >> sh) & 0xffffu;
        const int slot = off[pix[k]] + (int)cur + rank;
        out_n[slot] = n;
        out_c[slot] = g * w[k];
      }
      __syncwarp();
    }
  }
}

// features [C, N] -> featT [N, Cq]  (32 x 32 tiles through shared memory); pad columns are zero
__global__ void transpose_features_kernel(const float* __restrict__ f, float* __restrict__ ft, int C, int N, int Cq) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, n = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && n < N) ? __ldg(f + (int64_t)c * N + n) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, c = c0 + threadIdx.x;
    if (n < N && c < Cq) ft[(int64_t)n * Cq + c] = tile[threadIdx.x][r];
  }
}

// one warp per (sample, pixel): d_fmap[b, pix, c] += sum_e coef[e] * featT[n[e], c], entries in segment order
template <int NV>
__global__ void __launch_bounds__(256) readout_gather_kernel(v1t_readout_shape s, const int* __restrict__ seg_off,
                                                             const int* __restrict__ sorted_n,
                                                             const float* __restrict__ sorted_coef,
                                                             const float* __restrict__ ft, int Cq, float* __restrict__ d_fmap) {
  const int L = s.gh * s.gw, C = s.channels, N = s.neurons;
  const int lane = threadIdx.x & 31;
  const int64_t wg = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wg >= (int64_t)s.batch * L) return;
  const int b = (int)(wg / L), pix = (int)(wg % L);
  const int beg = seg_off[(int64_t)b * (L + 1) + pix], end = seg_off[(int64_t)b * (L + 1) + pix + 1];
  const int* en = sorted_n + (int64_t)b * 4 * N;
  const float* ec = sorted_coef + (int64_t)b * 4 * N;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  int e = beg;
  for (; e + 4 <= end; e += 4) {  // four entries in flight
    int n4[4];
    float c4[4], v[4][NV];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      n4[u] = __ldg(en + e + u);
      c4[u] = __ldg(ec + e + u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < NV; ++i) v[u][i] = (lane + 32 * i < C) ? __ldg(ft + (int64_t)n4[u] * Cq + lane + 32 * i) : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] = fmaf(c4[u], v[u][i], acc[i]);
  }
  for (; e < end; ++e) {
    const int n = __ldg(en + e);
    const float c = __ldg(ec + e);
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < C) acc[i] = fmaf(c, __ldg(ft + (int64_t)n * Cq + lane + 32 * i), acc[i]);
  }
  if (beg == end) return;  // nothing lands here: the caller's buffer keeps its value
  float* dst = d_fmap + (int64_t)b * s.fs_b + (int64_t)(pix / s.gw) * s.fs_y + (int64_t)(pix % s.gw) * s.fs_x;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (lane + 32 * i < C) dst[lane + 32 * i] += acc[i];
}

// out[i] = sum_p part[p*n + i]
__global__ void sum_parts_kernel(const float* __restrict__ part, int parts, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < parts; ++p) s += part[(int64_t)p * n + i];
  out[i] = s;
}

// split [tiles][N][7] -> d_bias[N], d_mu[N,2], d_sigma[N,4]
__global__ void small_finish_kernel(const float* __restrict__ part, int tiles, int N, float* __restrict__ d_bias,
                                    float* __restrict__ d_mu, float* __restrict__ d_sigma) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float a[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int t = 0; t < tiles; ++t)
#pragma unroll
    for (int k = 0; k < 7; ++k) a[k] += part[((int64_t)t * N + n) * 7 + k];
  if (d_bias) d_bias[n] = a[0];
  if (d_mu) { d_mu[2 * n] = a[1]; d_mu[2 * n + 1] = a[2]; }
  if (d_sigma) { d_sigma[4 * n] = a[3]; d_sigma[4 * n + 1] = a[4]; d_sigma[4 * n + 2] = a[5]; d_sigma[4 * n + 3] = a[6]; }
}

__global__ void elu1_fwd_kernel(const float* __restrict__ z, float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = elu1(z[i]);
}
__global__ void elu1_bwd_kernel(const float* __restrict__ z, const float* __restrict__ dy, float* __restrict__ dz,
                                int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = z[i];
    dz[i] = dy[i] * (v > 0.f ? 1.f : expf(v));
  }
}
__global__ void poisson_fwd_kernel(const float* __restrict__ yp, const float* __restrict__ yt, int64_t n, float eps,
                                   float* __restrict__ partials) {
  __shared__ float red[8];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = yp[i] + eps, t = yt[i] + eps;
    s += p - t * logf(p);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}
__global__ void poisson_bwd_kernel(const float* __restrict__ yp, const float* __restrict__ yt, int64_t n, float eps,
                                   float scale, const float* __restrict__ dloss, float* __restrict__ dy) {
  const float g = scale * (dloss ? *dloss : 1.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dy[i] = g * (1.f - (yt[i] + eps) / (yp[i] + eps));
}

struct ReadoutScratch {
  float* loss_partials;  // [tiles * ctas_x]
  float* feat_part;      // [tiles][C][N]
  float* small_part;     // [tiles][N][7]
  float* shift_part;     // [ctas_x][B][2]
  int* seg_off;          // [B][L + 1] segment starts of the pixel-major d_fmap pass
  int* sorted_n;         // [B][4 N] neuron of every (neuron, corner) entry, sorted by pixel
  float* sorted_coef;    // [B][4 N] dz * bilinear weight of the entry
  float* feat_t;         // [N][Cq] transposed features
  size_t total;
};

ReadoutScratch carve(const v1t_readout_shape& s, void* base) {
  const int64_t ctas_x = cdiv(s.neurons, kNeuronsPerCta), tiles = cdiv(s.batch, kBatchTile);
  char* p = (char*)base;
  ReadoutScratch r;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* q = p ? p + off : nullptr;
    off += (size_t)round_up((int64_t)bytes, 256);
    return (float*)q;
  };
  r.loss_partials = take(sizeof(float) * ctas_x * tiles);
  r.feat_part = take(sizeof(float) * tiles * (size_t)s.channels * s.neurons);
  r.small_part = take(sizeof(float) * tiles * (size_t)s.neurons * 7);
  r.shift_part = take(sizeof(float) * ctas_x * (size_t)s.batch * 2);
  r.seg_off = (int*)take(sizeof(int) * (size_t)s.batch * ((size_t)s.gh * s.gw + 1));
  r.sorted_n = (int*)take(sizeof(int) * (size_t)s.batch * 4 * s.neurons);
  r.sorted_coef = take(sizeof(float) * (size_t)s.batch * 4 * s.neurons);
  r.feat_t = take(sizeof(float) * (size_t)s.neurons * round_up(s.channels, 4));
  r.total = off;
  return r;
}

int check_shape(const v1t_readout_shape* s) {
  V1T_CHECK_ARG(s, "readout: null shape");
  V1T_CHECK_ARG(s->batch > 0 && s->neurons > 0 && s->channels > 0 && s->gh > 0 && s->gw > 0, "readout: bad shape");
  V1T_CHECK_ARG(s->channels <= 512, "readout: channels %d > 512 unsupported", s->channels);
  return V1T_OK;
}

// 128-bit path: base pointers 16-byte aligned and every stride a multiple of 4 floats
bool vec4_ok(const v1t_readout_shape& s, const float* fmap, const float* d_fmap) {
  static const bool enabled = [] {  // measurement knob: V1T_READOUT_V4=0 forces the scalar kernels
    const char* e = getenv("V1T_READOUT_V4");
    return !(e && e[0] == '0');
  }();
  return enabled && (((uintptr_t)fmap | (uintptr_t)d_fmap) & 15u) == 0 && s.fs_b % 4 == 0 && s.fs_y % 4 == 0 && s.fs_x % 4 == 0 &&
         s.channels >= 4;
}

size_t sort_smem(const v1t_readout_shape& s) {
  const size_t L = (size_t)s.gh * s.gw;
  return sizeof(uint32_t) * kSortWarps * ((L + 1) / 2) + sizeof(int) * (L + 1);
}
// V1T_READOUT_DFMAP=sorted selects the pixel-major, atomic-free (bitwise reproducible) d_fmap pass when the per-(warp, pixel)
// counters of one sample fit shared memory; the default is the red.global.add scatter of the neuron-major kernel, which is
// faster: measured at B = 16, N = 8000 the scatter adds 30-45 us to that kernel, the three extra launches cost 190 us
// (positions spread over the map) to 670 us (all neurons on a few pixels, one warp walking ~1000 entries per pixel)
bool dfmap_sorted(const v1t_readout_shape& s) {
  const char* e = getenv("V1T_READOUT_DFMAP");  // read on every call: tests switch it between launches
  const bool atomic = !(e && e[0] == 's');
  // 16-bit counters: a pixel holds at most 4 N entries of one sample
  return !atomic && sort_smem(s) <= 200 * 1024 && (int64_t)4 * s.neurons < 65536 && (int64_t)s.batch * 4 * s.neurons < (1ll << 31);
}
size_t fwd_smem(const v1t_readout_shape& s) { return sizeof(float) * ((size_t)s.channels * 33 + kBatchTile * 33); }
size_t bwd_smem(const v1t_readout_shape& s) { return sizeof(float) * ((size_t)s.channels * 33 + kBatchTile * 64); }

}  // namespace
}  // namespace v1t

using namespace v1t;

extern "C" size_t v1t_readout_scratch_bytes(const v1t_readout_shape* s) {
  if (!s) return 0;
  return carve(*s, nullptr).total;
}

#define V1T_NC_DISPATCH(channels, CALL) \
  if (channels <= 128) { CALL(1); }     \
  else if (channels <= 256) { CALL(2); } \
  else { CALL(4); }

#define V1T_NV_DISPATCH(nv, CALL)             \
  if (nv <= 1) { CALL(1); }                   \
  else if (nv <= 2) { CALL(2); }              \
  else if (nv <= 5) { CALL(5); }              \
  else if (nv <= 8) { CALL(8); }              \
  else { CALL(16); }

extern "C" int v1t_readout_forward(const v1t_readout_shape* s, const float* fmap, const float* mu,
                                   const float* sigma, const float* noise, const float* shifts,
                                   const float* features, const float* bias, const float* y_true, float loss_scale,
                                   float* z, float* y_out, float* loss_out, void* scratch, void* stream) {
  V1T_TRY(check_shape(s));
  V1T_CHECK_ARG(fmap && mu && features && z, "readout_forward: null tensor");
  V1T_CHECK_ARG(!noise || sigma, "readout_forward: noise given without sigma");
  V1T_CHECK_ARG(!loss_out || (y_true && scratch), "readout_forward: loss needs y_true and scratch");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(V1T_PHASE_READOUT_FWD, st);
  ReadoutScratch ws = carve(*s, scratch);
  dim3 grid(cdiv(s->neurons, kNeuronsPerCta), cdiv(s->batch, kBatchTile));
  const size_t smem = fwd_smem(*s);
  const int nv = cdiv(s->channels, 32);
#define CALL(NVV)                                                                                              \
  do {                                                                                                         \
    if (smem > 48 * 1024)                                                                                      \
      V1T_CUDA(cudaFuncSetAttribute(readout_forward_kernel<NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                    (int)smem));                                                               \
    readout_forward_kernel<NVV><<<grid, kFwdWarps * 32, smem, st>>>(*s, fmap, mu, sigma, noise, shifts, features, \
                                                                 bias, y_true, z, y_out,                       \
                                                                 loss_out ? ws.loss_partials : nullptr);       \
  } while (0)
  V1T_NV_DISPATCH(nv, CALL)
#undef CALL
  V1T_LAUNCH_CHECK();
  if (loss_out) {
    sum_scale_kernel<<<1, 256, 0, st>>>(ws.loss_partials, (int)(grid.x * grid.y), loss_scale, loss_out);
    V1T_LAUNCH_CHECK();
  }
  return V1T_OK;
}

extern "C" int v1t_readout_backward(const v1t_readout_shape* s, const float* fmap, const float* mu,
                                    const float* sigma, const float* noise, const float* shifts,
                                    const float* features, const float* z, const float* dz, const float* y_true,
                                    float loss_scale, float dloss, float* d_fmap, float* d_mu, float* d_sigma,
                                    float* d_shifts, float* d_features, float* d_bias, void* scratch,
                                    void* stream) {
  V1T_TRY(check_shape(s));
  V1T_CHECK_ARG(fmap && mu && features && scratch, "readout_backward: null tensor");
  V1T_CHECK_ARG(dz || (z && y_true), "readout_backward: need dz, or z and y_true for the fused Poisson gradient");
  V1T_CHECK_ARG(!noise || sigma, "readout_backward: noise given without sigma");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(V1T_PHASE_READOUT_BWD, st);
  ReadoutScratch ws = carve(*s, scratch);
  dim3 grid(cdiv(s->neurons, kNeuronsPerCta), cdiv(s->batch, kBatchTile));
  const int tiles = grid.y;
  const size_t smem = bwd_smem(*s);
  const int nv = cdiv(s->channels, 32);
  const bool want_small = d_bias || d_mu || d_sigma;
  float* const d_fmap_out = d_fmap;
  const bool sorted = d_fmap && dfmap_sorted(*s);
  if (sorted) d_fmap = nullptr;  // the neuron-major kernel skips its scatter; the pixel-major pass below writes d_fmap
  // with a single batch tile the per-tile slabs ARE the outputs: write d_features straight to its destination
  float* feat_dst = d_features ? (tiles == 1 ? d_features : ws.feat_part) : nullptr;
#define CALL(NVV)                                                                                               \
  do {                                                                                                          \
    if (smem > 48 * 1024)                                                                                       \
      V1T_CUDA(cudaFuncSetAttribute(readout_backward_kernel<NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                    (int)smem));                                                                \
    readout_backward_kernel<NVV><<<grid, kWarps * 32, smem, st>>>(                                              \
        *s, fmap, mu, sigma, noise, shifts, features, z, dz, y_true, loss_scale * dloss, d_fmap, feat_dst,      \
        want_small ? ws.small_part : nullptr, d_shifts ? ws.shift_part : nullptr);                              \
  } while (0)
#define CALL4(NCC)                                                                                                \
  do {                                                                                                            \
    if (smem > 48 * 1024)                                                                                         \
      V1T_CUDA(cudaFuncSetAttribute(readout_backward_v4_kernel<NCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem));                                                                  \
    readout_backward_v4_kernel<NCC><<<grid, kWarps * 32, smem, st>>>(                                             \
        *s, fmap, mu, sigma, noise, shifts, features, z, dz, y_true, loss_scale * dloss, d_fmap, feat_dst,        \
        want_small ? ws.small_part : nullptr, d_shifts ? ws.shift_part : nullptr);                                \
  } while (0)
  if (vec4_ok(*s, fmap, d_fmap)) {
    V1T_NC_DISPATCH(s->channels, CALL4)
  } else {
    V1T_NV_DISPATCH(nv, CALL)
  }
#undef CALL4
#undef CALL
  V1T_LAUNCH_CHECK();
  if (d_features && tiles > 1) {
    const int64_t n = (int64_t)s->channels * s->neurons;
    sum_parts_kernel<<<cdiv(n, 256), 256, 0, st>>>(ws.feat_part, tiles, n, d_features);
    V1T_LAUNCH_CHECK();
  }
  if (want_small) {
    small_finish_kernel<<<cdiv(s->neurons, 128), 128, 0, st>>>(ws.small_part, tiles, s->neurons, d_bias, d_mu,
                                                               d_sigma);
    V1T_LAUNCH_CHECK();
  }
  if (d_shifts) {
    const int64_t n = (int64_t)s->batch * 2;
    sum_parts_kernel<<<cdiv(n, 256), 256, 0, st>>>(ws.shift_part, (int)grid.x, n, d_shifts);
    V1T_LAUNCH_CHECK();
  }
  if (sorted) {
    const int Cq = (int)round_up(s->channels, 4);
    const size_t ssm = sort_smem(*s);
    static size_t configured = 0;
    if (ssm > 48 * 1024 && ssm > configured) {
      V1T_CUDA(cudaFuncSetAttribute(readout_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
      configured = ssm;
    }
    readout_sort_kernel<<<s->batch, kSortWarps * 32, ssm, st>>>(*s, mu, sigma, noise, shifts, z, dz, y_true, loss_scale * dloss,
                                                                ws.seg_off, ws.sorted_n, ws.sorted_coef);
    V1T_LAUNCH_CHECK();
    transpose_features_kernel<<<dim3(cdiv(s->neurons, 32), cdiv(Cq, 32)), dim3(32, 8), 0, st>>>(features, ws.feat_t, s->channels,
                                                                                                 s->neurons, Cq);
    V1T_LAUNCH_CHECK();
    const int64_t warps = (int64_t)s->batch * s->gh * s->gw;
#define CALLG(NVV) readout_gather_kernel<NVV><<<(unsigned)cdiv(warps, 8), 256, 0, st>>>(*s, ws.seg_off, ws.sorted_n, ws.sorted_coef, ws.feat_t, Cq, d_fmap_out)
    V1T_NV_DISPATCH(nv, CALLG)
#undef CALLG
    V1T_LAUNCH_CHECK();
  }
  return V1T_OK;
}

extern "C" int v1t_elu1_forward(const float* z, float* y, int64_t n, void* stream) {
  V1T_CHECK_ARG(z && y && n >= 0, "elu1_forward: bad argument");
  if (n == 0) return V1T_OK;
  elu1_fwd_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(z, y, n);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
extern "C" int v1t_elu1_backward(const float* z, const float* dy, float* dz, int64_t n, void* stream) {
  V1T_CHECK_ARG(z && dy && dz && n >= 0, "elu1_backward: bad argument");
  if (n == 0) return V1T_OK;
  elu1_bwd_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(z, dy, dz, n);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

static int poisson_blocks(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 4)); }

extern "C" size_t v1t_poisson_scratch_bytes(int64_t n) { return sizeof(float) * (size_t)poisson_blocks(n); }

extern "C" int v1t_poisson_forward(const float* y_pred, const float* y_true, int64_t n, float eps, float loss_scale,
                                   float* loss_out, void* scratch, void* stream) {
  V1T_CHECK_ARG(y_pred && y_true && loss_out && scratch && n >= 0, "poisson_forward: bad argument");
  const int blocks = poisson_blocks(n);
  poisson_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y_pred, y_true, n, eps, (float*)scratch);
  V1T_LAUNCH_CHECK();
  sum_scale_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const float*)scratch, blocks, loss_scale, loss_out);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
extern "C" int v1t_poisson_backward(const float* y_pred, const float* y_true, int64_t n, float eps, float loss_scale,
                                    const float* dloss, float* dy, void* stream) {
  V1T_CHECK_ARG(y_pred && y_true && dy && n >= 0, "poisson_backward: bad argument");
  if (n == 0) return V1T_OK;
  poisson_bwd_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      y_pred, y_true, n, eps, loss_scale, dloss, dy);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
