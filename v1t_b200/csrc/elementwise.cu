// Row-wise / elementwise kernels of the ViT core: im2col for nn.Unfold (vit.py:68-71), CLS/pos rows
// (vit.py:124-127), BehaviorMLP (vit.py:181-202), LayerNorm fwd/bwd (eps 1e-5, biased variance),
// softmax fwd/bwd over the key dim (vit.py:262), exact-erf GELU fwd/bwd (vit.py:147), inverted dropout
// with a replayable counter-based RNG, and the deterministic column/batch reductions used for the
// bias / LayerNorm-affine / pos-embedding gradients.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"
#include <algorithm>

namespace v1t {
namespace {

constexpr float kLnEps = 1e-5f;

// ---------------------------------------------------------------- patches ------------------------------
__global__ void im2col_kernel(const float* __restrict__ img, float* __restrict__ patches, int B, int C, int H,
                              int W, int p, int s, int gh, int gw) {
  const int pd = C * p * p;
  const int64_t total = (int64_t)B * gh * gw * pd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % pd);
    const int64_t row = i / pd;
    const int l = (int)(row % (gh * gw));
    const int b = (int)(row / (gh * gw));
    const int kw = e % p, kh = (e / p) % p, ch = e / (p * p);
    const int r = l / gw, c = l % gw;
    patches[i] = __ldg(img + (((int64_t)b * C + ch) * H + r * s + kh) * W + c * s + kw);
  }
}

// d_images[b,ch,y,x] = sum over patches covering (y,x) of d_patches  (gather form: atomic-free)
__global__ void col2im_kernel(const float* __restrict__ dpatches, float* __restrict__ dimg, int B, int C, int H,
                              int W, int p, int s, int gh, int gw) {
  const int pd = C * p * p;
  const int64_t total = (int64_t)B * C * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), ch = (int)((i / ((int64_t)W * H)) % C);
    const int b = (int)(i / ((int64_t)W * H * C));
    float acc = 0.f;
    for (int kh = 0; kh < p; ++kh) {
      const int ry = y - kh;
      if (ry < 0 || ry % s != 0 || ry / s >= gh) continue;
      for (int kw = 0; kw < p; ++kw) {
        const int cx = x - kw;
        if (cx < 0 || cx % s != 0 || cx / s >= gw) continue;
        const int l = (ry / s) * gw + cx / s;
        acc += __ldg(dpatches + ((int64_t)b * gh * gw + l) * pd + (ch * p + kh) * p + kw);
      }
    }
    dimg[i] = acc;
  }
}

__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x,
                                int B, int T, int E, int ld) {
  const int b = blockIdx.x;
  for (int e = threadIdx.x; e < E; e += blockDim.x) x[(int64_t)b * T * ld + e] = cls[e] + pos[e];
}

__global__ void dropout_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int cols,
                                    int64_t ld, DropSpec dr) {
  const float inv_keep = 1.f / (1.f - dr.p);
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    dst[r * ld + c] = src[r * ld + c] * dropout_mult(dr.seed, dr.site, (uint64_t)r * drop_stride(cols) + c, dr.p, inv_keep);
  }
}

// 8 columns (one 16-byte plane chunk) per thread: two Philox calls, 128-bit accesses, optional operand planes of
// the result (zero in the pad columns)
__global__ void dropout_rows8_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int cols,
                                     int64_t ld, int chunks, DropSpec dr, PlaneOut pl) {
  const float inv_keep = 1.f / (1.f - dr.p);
  const int64_t total = (pl.hi ? (int64_t)pl.rows_p : rows) * chunks;  // plane pad rows are zero-filled too
  const int64_t drop_ld = drop_stride(cols);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / chunks;
    const int ch = (int)(i % chunks);
    const int c0 = ch * 8;
    float v[8], m8[8];
    if (c0 < cols && r < rows) dropout_mult8(dr.seed, dr.site, ((uint64_t)r * drop_ld + c0) >> 3, dr.p, inv_keep, m8);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = c0 + 4 * q;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (c < cols && r < rows) {
        const float m[4] = {m8[4 * q], m8[4 * q + 1], m8[4 * q + 2], m8[4 * q + 3]};
        if (c + 4 <= cols) {
          const float4 s4 = *reinterpret_cast<const float4*>(src + r * ld + c);
          x[0] = s4.x * m[0]; x[1] = s4.y * m[1]; x[2] = s4.z * m[2]; x[3] = s4.w * m[3];
          *reinterpret_cast<float4*>(dst + r * ld + c) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
          for (int e = 0; e < 4 && c + e < cols; ++e) {
            x[e] = src[r * ld + c + e] * m[e];
            dst[r * ld + c + e] = x[e];
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) v[4 * q + e] = x[e];
    }
    if (pl.hi) {
      uint4 hi, lo;
      tc::split8(v, hi, lo);
      const int64_t off = tc::plane_chunk_off(ch >> 2, pl.rows_p, r, ch & 3);
      *reinterpret_cast<uint4*>(pl.hi + off) = hi;
      if (pl.lo) *reinterpret_cast<uint4*>(pl.lo + off) = lo;
    }
  }
}

// The same, plus the column sums of the result (the bias gradient of the Linear whose output gradient this is):
// a thread keeps ONE chunk (8 columns) and walks rows, so the sums stay in registers; per-CTA partials, fixed order.
// block = 256 threads = (256 / chunks) row lanes x chunks.
__global__ void __launch_bounds__(256) dropout_rows8_colsum_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                   int64_t rows, int cols, int64_t ld, int chunks,
                                                                   DropSpec dr, PlaneOut pl, float* __restrict__ partials) {
  extern __shared__ float red[];  // [row lanes][chunks * 8]
  const float inv_keep = 1.f / (1.f - dr.p);
  const int rpp = 256 / chunks;
  const int ch = threadIdx.x % chunks, ty = threadIdx.x / chunks;
  const int64_t rows_all = pl.hi ? (int64_t)pl.rows_p : rows;
  const int64_t drop_ld = drop_stride(cols);
  const int c0 = ch * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ty < rpp) {
    for (int64_t r = (int64_t)blockIdx.x * rpp + ty; r < rows_all; r += (int64_t)gridDim.x * rpp) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (c0 < cols && r < rows) {
        float m8[8];
        dropout_mult8(dr.seed, dr.site, ((uint64_t)r * drop_ld + c0) >> 3, dr.p, inv_keep, m8);
        if (c0 + 8 <= cols) {
          const float4 a = *reinterpret_cast<const float4*>(src + r * ld + c0);
          const float4 b = *reinterpret_cast<const float4*>(src + r * ld + c0 + 4);
          v[0] = a.x * m8[0]; v[1] = a.y * m8[1]; v[2] = a.z * m8[2]; v[3] = a.w * m8[3];
          v[4] = b.x * m8[4]; v[5] = b.y * m8[5]; v[6] = b.z * m8[6]; v[7] = b.w * m8[7];
          *reinterpret_cast<float4*>(dst + r * ld + c0) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(dst + r * ld + c0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
          for (int e = 0; e < 8 && c0 + e < cols; ++e) {
            v[e] = src[r * ld + c0 + e] * m8[e];
            dst[r * ld + c0 + e] = v[e];
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
      }
      if (pl.hi) {
        uint4 hi, lo;
        tc::split8(v, hi, lo);
        const int64_t off = tc::plane_chunk_off(ch >> 2, pl.rows_p, r, ch & 3);
        *reinterpret_cast<uint4*>(pl.hi + off) = hi;
        if (pl.lo) *reinterpret_cast<uint4*>(pl.lo + off) = lo;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[ty * chunks * 8 + c0 + e] = acc[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float s = 0.f;
    for (int y = 0; y < rpp; ++y) s += red[y * chunks * 8 + c];
    partials[(int64_t)blockIdx.x * cols + c] = s;
  }
}

__global__ void dropout_mask_kernel(float* __restrict__ out, int64_t n, DropSpec dr) {
  const float inv_keep = 1.f / (1.f - dr.p);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = dropout_mult(dr.seed, dr.site, i, dr.p, inv_keep);
}

// ---------------------------------------------------------------- behaviour MLP ------------------------
// grid = B; lat[b,:] = tanh(W3 tanh(W0 beh[b] + b0) + b3)
__global__ void bmlp_forward_kernel(const float* __restrict__ beh, const float* __restrict__ w0,
                                    const float* __restrict__ b0, const float* __restrict__ w3,
                                    const float* __restrict__ b3, float* __restrict__ hid, float* __restrict__ lat,
                                    int bdim, int H, int E) {
  extern __shared__ float sh[];  // [H]
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float a = b0 ? b0[j] : 0.f;
    for (int i = 0; i < bdim; ++i) a = fmaf(w0[j * bdim + i], beh[b * bdim + i], a);
    a = tanhf(a);
    sh[j] = a;
    hid[(int64_t)b * H + j] = a;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float a = b3 ? b3[e] : 0.f;
    for (int j = 0; j < H; ++j) a = fmaf(w3[e * H + j], sh[j], a);
    lat[(int64_t)b * E + e] = tanhf(a);
  }
}

__global__ void tanh_grad_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz,
                                 int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = y[i];
    dz[i] = dy[i] * (1.f - v * v);
  }
}

// Backward of BehaviorMLP (vit.py:181-202), the whole problem is B x 155 x 77: given dlat[b,e] (the gradient of the
// latent added to every token), dz3 = dlat * (1 - lat^2); dW3[e,j] = sum_b dz3[b,e] hid[b,j]; db3 = colsum(dz3);
// dhid = dz3 W3; dz0 = dhid * (1 - hid^2); dW0[j,i] = sum_b dz0[b,j] beh[b,i]; db0 = colsum(dz0).
// gridDim.x CTAs without any cross-CTA dependency: every CTA recomputes dz3 (tiny) and owns a slice of the e rows
// (dW3, db3) and a slice of the hidden units j (dhid -> dz0 -> dW0, db0).  Fixed summation order (deterministic).
// Replaces 8 latency-bound launches per block.
__global__ void __launch_bounds__(256) bmlp_backward_kernel(const float* __restrict__ dlat, const float* __restrict__ lat,
                                                            const float* __restrict__ hid, const float* __restrict__ beh,
                                                            const float* __restrict__ w3, float* __restrict__ dw0,
                                                            float* __restrict__ db0, float* __restrict__ dw3,
                                                            float* __restrict__ db3, int B, int bdim, int H, int E) {
  extern __shared__ __align__(16) float sh[];
  float* dz3 = sh;            // [B][E]
  float* dz0 = sh + B * E;    // [B][jn] (this CTA's hidden units)
  const int G = gridDim.x, g = blockIdx.x;
  const int e0 = (int)((int64_t)E * g / G), e1 = (int)((int64_t)E * (g + 1) / G);
  const int j0 = (int)((int64_t)H * g / G), j1 = (int)((int64_t)H * (g + 1) / G), jn = j1 - j0;
  {  // every CTA recomputes dz3: 128-bit loads, four of them in flight per thread
    const int n = B * E;
    const bool vec = ((reinterpret_cast<uintptr_t>(lat) | reinterpret_cast<uintptr_t>(dlat)) & 15) == 0;
    const int n4 = vec ? n / 4 : 0;
    const float4* l4 = reinterpret_cast<const float4*>(lat);
    const float4* d4 = reinterpret_cast<const float4*>(dlat);
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 y = __ldg(l4 + i), g4 = __ldg(d4 + i);
      reinterpret_cast<float4*>(dz3)[i] = make_float4(g4.x * (1.f - y.x * y.x), g4.y * (1.f - y.y * y.y), g4.z * (1.f - y.z * y.z),
                                                      g4.w * (1.f - y.w * y.w));
    }
    for (int i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const float y = lat[i];
      dz3[i] = dlat[i] * (1.f - y * y);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (e1 - e0) * H && dw3; i += blockDim.x) {
    const int e = e0 + i / H, j = i % H;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dz3[b * E + e], hid[b * H + j], s);
    dw3[e * H + j] = s;
  }
  for (int e = e0 + threadIdx.x; e < e1 && db3; e += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dz3[b * E + e];
    db3[e] = s;
  }
  // one warp per (sample, hidden unit): the lanes split the E-long dot product (a serial loop per thread was a chain of
  // E dependent L2 loads of w3 -- most of this kernel's 89 us at 112 samples); fixed butterfly order
  // (four items per warp and pass, so that their loads overlap)
  for (int i0 = (threadIdx.x >> 5) * 4; i0 < B * jn; i0 += (blockDim.x >> 5) * 4) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int e = threadIdx.x & 31; e < E; e += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = min(i0 + u, B * jn - 1);
        s[u] = fmaf(dz3[(i / jn) * E + e], __ldg(w3 + e * H + j0 + i % jn), s[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float t = warp_sum(s[u]);
      const int i = i0 + u;
      if ((threadIdx.x & 31) == 0 && i < B * jn) {
        const int b = i / jn, j = j0 + i % jn;
        const float y = hid[b * H + j];
        dz0[b * jn + (j - j0)] = t * (1.f - y * y);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < jn * bdim && dw0; i += blockDim.x) {
    const int j = i / bdim, k = i % bdim;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dz0[b * jn + j], beh[b * bdim + k], s);
    dw0[(j0 + j) * bdim + k] = s;
  }
  for (int j = threadIdx.x; j < jn && db0; j += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dz0[b * jn + j];
    db0[j0 + j] = s;
  }
}

// ---------------------------------------------------------------- LayerNorm ----------------------------
// one warp per row; lane owns columns lane + 32*i.  NV = ceil(E/32) rounded to a compiled size.
template <int NV>
__global__ void __launch_bounds__(256) ln_forward_kernel(const float* __restrict__ x_in, const float* __restrict__ add,
                                                         int rows_per_batch, float* __restrict__ x_out,
                                                         const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float* __restrict__ h,
                                                         float* __restrict__ stats, int64_t rows, int E, int ld,
                                                         PlaneOut pl) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float g[NV], bt[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    g[i] = c < E ? gamma[c] : 0.f;
    bt[i] = c < E ? beta[c] : 0.f;
  }
  for (int64_t r = rows + warp; pl.hi && r < pl.rows_p; r += nwarps) {  // zero rows padding the planes
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < 32 * ((E + 31) / 32)) {
        const int64_t off = tc::plane_chunk_off(i, pl.rows_p, r, lane >> 3) + (lane & 7) * 2;
        *reinterpret_cast<uint16_t*>(pl.hi + off) = 0;
        if (pl.lo) *reinterpret_cast<uint16_t*>(pl.lo + off) = 0;
      }
    }
  }
  // U rows per warp and iteration, all loads issued before the first reduction
  constexpr int U = NV <= 8 ? 2 : 1;
  for (int64_t r0 = warp * U; r0 < rows; r0 += nwarps * U) {
    float v[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u;
      const float* xr = x_in + r * ld;
      const float* ar = add ? add + (r / rows_per_batch) * E : nullptr;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        const bool in = r < rows && c < E;
        float t = in ? __ldg(xr + c) : 0.f;
        if (ar && in) t += __ldg(ar + c);
        v[u][i] = t;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u;
      if (r >= rows) break;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) s += v[u][i];
      const float mean = warp_sum(s) / (float)E;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        const float d = c < E ? v[u][i] - mean : 0.f;
        q += d * d;
      }
      const float rstd = rsqrtf(warp_sum(q) / (float)E + kLnEps);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < ld) {
          const bool in = c < E;
          if (x_out) x_out[r * ld + c] = in ? v[u][i] : 0.f;
          if (h) h[r * ld + c] = in ? (v[u][i] - mean) * rstd * g[i] + bt[i] : 0.f;
        }
        // GEMM-operand planes of the normalised row: lane = column within atom i, 8 lanes per 16-byte chunk
        if (pl.hi && c < 32 * ((E + 31) / 32)) {
          const float y = c < E ? (v[u][i] - mean) * rstd * g[i] + bt[i] : 0.f;
          const __nv_bfloat16 hb = __float2bfloat16_rn(y);
          const int64_t off = tc::plane_chunk_off(i, pl.rows_p, r, lane >> 3) + (lane & 7) * 2;
          *reinterpret_cast<__nv_bfloat16*>(pl.hi + off) = hb;
          if (pl.lo) *reinterpret_cast<__nv_bfloat16*>(pl.lo + off) = __float2bfloat16_rn(y - __bfloat162float(hb));
        }
      }
      if (lane == 0 && stats) {
        stats[2 * r] = mean;
        stats[2 * r + 1] = rstd;
      }
    }
  }
}

// dx_accum[r,:] += rstd*(dxh - mean(dxh) - xhat*mean(dxh*xhat)), dxh = dh*gamma; per-CTA partial dgamma/dbeta
template <int NV>
__global__ void __launch_bounds__(256) ln_backward_kernel(const float* __restrict__ dh, const float* __restrict__ x,
                                                          const float* __restrict__ stats,
                                                          const float* __restrict__ gamma,
                                                          float* __restrict__ dx_accum, float* __restrict__ partials,
                                                          int64_t rows, int E, int ld) {
  __shared__ float red[2][8][32 * NV];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * 8 + wid;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  float g[NV], dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    g[i] = c < E ? gamma[c] : 0.f;
    dg[i] = 0.f;
    db[i] = 0.f;
  }
  // U rows per warp and iteration with every load (dh, x, the dx accumulator, the row statistics) issued before the first
  // use: one row at a time left ~10 loads per warp in flight and the kernel at 2.7 TB/s
  constexpr int U = NV <= 8 ? 2 : 1;
  for (int64_t r0 = warp * U; r0 < rows; r0 += nwarps * U) {
    float d[U][NV], xv[U][NV], acc[U][NV], mean[U], rstd[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u;
      const bool row_in = r < rows;
      mean[u] = row_in ? __ldg(stats + 2 * r) : 0.f;
      rstd[u] = row_in ? __ldg(stats + 2 * r + 1) : 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        const bool in = row_in && c < E;
        d[u][i] = in ? __ldg(dh + r * ld + c) : 0.f;
        xv[u][i] = in ? __ldg(x + r * ld + c) : 0.f;
        acc[u][i] = in ? dx_accum[r * ld + c] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u;
      float xh[NV], dxh[NV];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const bool in = r < rows && lane + 32 * i < E;
        xh[i] = in ? (xv[u][i] - mean[u]) * rstd[u] : 0.f;
        dxh[i] = d[u][i] * g[i];
        dg[i] += d[u][i] * xh[i];
        db[i] += d[u][i];
        s1 += dxh[i];
        s2 += dxh[i] * xh[i];
      }
      s1 = warp_sum(s1) / (float)E;
      s2 = warp_sum(s2) / (float)E;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (r < rows && c < E) dx_accum[r * ld + c] = acc[u][i] + rstd[u] * (dxh[i] - s1 - xh[i] * s2);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    red[0][wid][lane + 32 * i] = dg[i];
    red[1][wid][lane + 32 * i] = db[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += red[0][w][c];
      b += red[1][w][c];
    }
    partials[((int64_t)blockIdx.x * 2) * E + c] = a;
    partials[((int64_t)blockIdx.x * 2 + 1) * E + c] = b;
  }
}

// dgamma[c] = sum_p partials[p][0][c], dbeta[c] = sum_p partials[p][1][c].  lane = column (coalesced rows of the
// partials), the 32 warps of a block take every 32nd partial (888 partials: 28 independent loads per lane instead of
// 111 -- the kernel is pure load latency), fixed-order smem reduction (deterministic).
// grid (ceil(E/32), 2): blockIdx.y = 0 -> dgamma, 1 -> dbeta
constexpr int kLnFinishWarps = 32;
__global__ void __launch_bounds__(kLnFinishWarps * 32) ln_finish_kernel(const float* __restrict__ partials,
                                                                        float* __restrict__ dgamma,
                                                                        float* __restrict__ dbeta, int parts, int E) {
  __shared__ float red[kLnFinishWarps][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int which = blockIdx.y;
  float s = 0.f;
  if (c < E) {
#pragma unroll 4
    for (int p = wid; p < parts; p += kLnFinishWarps) s += partials[((int64_t)p * 2 + which) * E + c];
  }
  red[wid][lane] = s;
  __syncthreads();
  if (wid == 0 && c < E) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kLnFinishWarps; ++w) t += red[w][lane];
    float* out = which == 0 ? dgamma : dbeta;
    if (out) out[c] = t;
  }
}

// ---------------------------------------------------------------- softmax ------------------------------
// one warp per row, in place; optional dropout applied to the stored probabilities (vit.py:262-263)
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ S, int64_t rows, int cols, int64_t ld,
                                                           DropSpec dr, int64_t row_offset) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  const float inv_keep = dr.p > 0.f ? 1.f / (1.f - dr.p) : 1.f;
  for (int64_t r = warp; r < rows; r += nwarps) {
    float* row = S + r * ld;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, row[c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float e = expf(row[c] - mx);
      row[c] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    for (int c = lane; c < cols; c += 32) {
      float p = row[c] * inv;
      if (dr.p > 0.f) p *= dropout_mult(dr.seed, dr.site, (uint64_t)(row_offset + r) * drop_stride(cols) + c, dr.p, inv_keep);
      row[c] = p;
    }
  }
}

// in: P = softmax probs (no dropout), dP = dL/d(dropped probs).  out: dP <- dS, P <- dropped probs.
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(float* __restrict__ P, float* __restrict__ dP,
                                                               int64_t rows, int cols, int64_t ld, DropSpec dr,
                                                               int64_t row_offset) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  const float inv_keep = dr.p > 0.f ? 1.f / (1.f - dr.p) : 1.f;
  for (int64_t r = warp; r < rows; r += nwarps) {
    float* p = P + r * ld;
    float* d = dP + r * ld;
    float dot = 0.f;
    for (int c = lane; c < cols; c += 32) {
      float dv = d[c];
      if (dr.p > 0.f) {
        const float m = dropout_mult(dr.seed, dr.site, (uint64_t)(row_offset + r) * drop_stride(cols) + c, dr.p, inv_keep);
        dv *= m;
        d[c] = dv;
      }
      dot += dv * p[c];
    }
    dot = warp_sum(dot);
    for (int c = lane; c < cols; c += 32) {
      const float pv = p[c];
      d[c] = pv * (d[c] - dot);
      if (dr.p > 0.f)
        p[c] = pv * dropout_mult(dr.seed, dr.site, (uint64_t)(row_offset + r) * drop_stride(cols) + c, dr.p, inv_keep);
    }
  }
}

// ---------------------------------------------------------------- GELU ---------------------------------

__global__ void gelu_forward_kernel(const float* __restrict__ u, float* __restrict__ g, int64_t rows, int cols,
                                    int64_t ld, DropSpec dr) {
  const float inv_keep = dr.p > 0.f ? 1.f / (1.f - dr.p) : 1.f;
  const int64_t total = rows * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld;
    const int c = (int)(i % ld);
    float v = 0.f;
    if (c < cols) {
      v = gelu_f(u[i]);
      if (dr.p > 0.f) v *= dropout_mult(dr.seed, dr.site, (uint64_t)r * drop_stride(cols) + c, dr.p, inv_keep);
    }
    g[i] = v;
  }
}

__global__ void gelu_backward_kernel(float* __restrict__ dg, const float* __restrict__ u, int64_t rows, int cols,
                                     int64_t ld, DropSpec dr) {
  const float inv_keep = dr.p > 0.f ? 1.f / (1.f - dr.p) : 1.f;
  const int64_t total = rows * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld;
    const int c = (int)(i % ld);
    float v = 0.f;
    if (c < cols) {
      v = dg[i] * gelu_df(u[i]);
      if (dr.p > 0.f) v *= dropout_mult(dr.seed, dr.site, (uint64_t)r * drop_stride(cols) + c, dr.p, inv_keep);
    }
    dg[i] = v;
  }
}

// ---------------------------------------------------------------- reductions ---------------------------
// partials[(b*strips + strip)*cols + c] = sum of X[b, rows in strip, c];  block (32, 8)
__global__ void colsum_kernel(const float* __restrict__ X, float* __restrict__ partials, int64_t rows, int cols,
                              int64_t xb, int64_t ld, int strips) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int strip = blockIdx.y, b = blockIdx.z;
  const int64_t per = (rows + strips - 1) / strips;
  const int64_t r0 = strip * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  if (c < cols) {
    const float* p = X + b * xb + c;
    int64_t r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {  // four independent loads in flight per thread (fixed summation order)
      const float v0 = p[r * ld], v1 = p[(r + 8) * ld], v2 = p[(r + 16) * ld], v3 = p[(r + 24) * ld];
      s += (v0 + v1) + (v2 + v3);
    }
    for (; r < r1; r += 8) s += p[r * ld];
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partials[((int64_t)b * strips + strip) * cols + c] = t;
  }
}

// out[b*out_ld + c] = sum_strip partials[(b*strips+strip)*cols + c].  lane = column (coalesced), the 32 warps of a
// block take every 32nd strip with 8 independent loads in flight (the kernel is pure load latency: with 8 warps and a
// serial loop the 5784 strips of the MLP bias gradient took 79 us), fixed-order smem reduction.
// grid (ceil(cols/32), batch)
constexpr int kColsumFinishWarps = 32;
__global__ void __launch_bounds__(kColsumFinishWarps * 32) colsum_finish_kernel(const float* __restrict__ partials,
                                                                                float* __restrict__ out, int batch, int cols,
                                                                                int strips, int64_t out_ld) {
  __shared__ float red[kColsumFinishWarps][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane, b = blockIdx.y;
  float s = 0.f;
  if (c < cols) {
    const float* src = partials + (int64_t)b * strips * cols + c;
    int k = wid;
    for (; k + 7 * kColsumFinishWarps < strips; k += 8 * kColsumFinishWarps) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (int64_t)(k + u * kColsumFinishWarps) * cols);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; k < strips; k += kColsumFinishWarps) s += __ldg(src + (int64_t)k * cols);
  }
  red[wid][lane] = s;
  __syncthreads();
  if (wid == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kColsumFinishWarps; ++w) t += red[w][lane];
    out[b * out_ld + c] = t;
  }
}

__global__ void batchsum_kernel(const float* __restrict__ X, float* __restrict__ out, int B, int64_t rows, int cols,
                                int64_t xb, int64_t ld, int64_t out_ld) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += X[b * xb + r * ld + c];
    out[r * out_ld + c] = s;
  }
}

inline int ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int im2col(const float* img, float* patches, int B, int C, int H, int W, int p, int s, int gh, int gw,
           cudaStream_t st) {
  im2col_kernel<<<ew_grid((int64_t)B * gh * gw * C * p * p), 256, 0, st>>>(img, patches, B, C, H, W, p, s, gh, gw);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
int col2im(const float* dpatches, float* dimg, int B, int C, int H, int W, int p, int s, int gh, int gw,
           cudaStream_t st) {
  col2im_kernel<<<ew_grid((int64_t)B * C * H * W), 256, 0, st>>>(dpatches, dimg, B, C, H, W, p, s, gh, gw);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
int cls_rows(const float* cls, const float* pos, float* x, int B, int T, int E, int ld, cudaStream_t st) {
  cls_rows_kernel<<<B, 128, 0, st>>>(cls, pos, x, B, T, E, ld);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
int dropout_rows(const float* src, float* dst, int64_t rows, int cols, int64_t ld, DropSpec dr, cudaStream_t st,
                 PlaneOut pl, float* colsum_out, float* partials, size_t partial_bytes) {
  const bool aligned = ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  if (colsum_out) {  // fused bias gradient
    const int chunks = cdiv(cols, 32) * 4;
    V1T_CHECK_ARG(aligned && chunks <= 256 && partials, "dropout_rows: fused column sums need aligned rows and <= 2048 columns");
    const int rpp = 256 / chunks;
    const int64_t rows_all = pl.hi ? (int64_t)pl.rows_p : rows;
    int grid = (int)std::min<int64_t>(cdiv(rows_all, rpp), (int64_t)kNumSMs * 4);
    grid = (int)std::min<int64_t>(grid, (int64_t)(partial_bytes / (sizeof(float) * (size_t)cols)));
    V1T_CHECK_ARG(grid >= 1, "dropout_rows: partials workspace too small");
    dropout_rows8_colsum_kernel<<<grid, 256, sizeof(float) * rpp * chunks * 8, st>>>(src, dst, rows, cols, ld, chunks, dr,
                                                                                     pl, partials);
    V1T_LAUNCH_CHECK();
    colsum_finish_kernel<<<dim3(cdiv(cols, 32), 1), kColsumFinishWarps * 32, 0, st>>>(partials, colsum_out, 1, cols, grid, 0);
    V1T_LAUNCH_CHECK();
    return V1T_OK;
  }
  if (aligned) {
    const int chunks = cdiv(cols, 32) * 4;  // 16-byte plane chunks (8 columns) per row
    dropout_rows8_kernel<<<ew_grid(rows * chunks), 256, 0, st>>>(src, dst, rows, cols, ld, chunks, dr, pl);
    V1T_LAUNCH_CHECK();
    return V1T_OK;
  }
  V1T_CHECK_ARG(!pl.hi, "dropout_rows: plane output needs 16-byte aligned rows");
  dropout_rows_kernel<<<ew_grid(rows * cols), 256, 0, st>>>(src, dst, rows, cols, ld, dr);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
int dropout_mask(float* out, int64_t n, DropSpec dr, cudaStream_t st) {
  dropout_mask_kernel<<<ew_grid(n), 256, 0, st>>>(out, n, dr);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
int bmlp_forward(const float* beh, const float* w0, const float* b0, const float* w3, const float* b3, float* hid,
                 float* lat, int B, int bdim, int H, int E, cudaStream_t st) {
  bmlp_forward_kernel<<<B, 128, H * sizeof(float), st>>>(beh, w0, b0, w3, b3, hid, lat, bdim, H, E);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
size_t bmlp_backward_smem(int B, int H, int E) { return sizeof(float) * (size_t)B * (E + H); }
int bmlp_backward(const float* dlat, const float* lat, const float* hid, const float* beh, const float* w3, float* dw0,
                  float* db0, float* dw3, float* db3, int B, int bdim, int H, int E, cudaStream_t st) {
  const size_t smem = bmlp_backward_smem(B, H, E);
  V1T_CHECK_ARG(smem <= 200 * 1024, "bmlp_backward: batch too large for the single-CTA kernel");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    V1T_CUDA(cudaFuncSetAttribute(bmlp_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  bmlp_backward_kernel<<<16, 256, smem, st>>>(dlat, lat, hid, beh, w3, dw0, db0, dw3, db3, B, bdim, H, E);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
int tanh_grad(const float* dy, const float* y, float* dz, int64_t n, cudaStream_t st) {
  tanh_grad_kernel<<<ew_grid(n), 256, 0, st>>>(dy, y, dz, n);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

#define V1T_LN_DISPATCH(NVV, KERNEL, ...)                       \
  if (nv <= NVV) {                                              \
    KERNEL<NVV><<<grid, 256, 0, st>>>(__VA_ARGS__);             \
  } else

int ln_forward(const float* x_in, const float* add, int rows_per_batch, float* x_out, const float* gamma,
               const float* beta, float* h, float* stats, int64_t rows, int E, int ld, cudaStream_t st, PlaneOut pl) {
  const int nv = cdiv(ld, 32);
  V1T_CHECK_ARG(nv <= 32 && ld >= E, "layer norm: emb dim %d (ld %d) unsupported (max 1024)", E, ld);
  const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)kNumSMs * 8);
  V1T_LN_DISPATCH(1, ln_forward_kernel, x_in, add, rows_per_batch, x_out, gamma, beta, h, stats, rows, E, ld, pl)
  V1T_LN_DISPATCH(2, ln_forward_kernel, x_in, add, rows_per_batch, x_out, gamma, beta, h, stats, rows, E, ld, pl)
  V1T_LN_DISPATCH(5, ln_forward_kernel, x_in, add, rows_per_batch, x_out, gamma, beta, h, stats, rows, E, ld, pl)
  V1T_LN_DISPATCH(8, ln_forward_kernel, x_in, add, rows_per_batch, x_out, gamma, beta, h, stats, rows, E, ld, pl)
  V1T_LN_DISPATCH(16, ln_forward_kernel, x_in, add, rows_per_batch, x_out, gamma, beta, h, stats, rows, E, ld, pl)
  { ln_forward_kernel<32><<<grid, 256, 0, st>>>(x_in, add, rows_per_batch, x_out, gamma, beta, h, stats, rows, E, ld, pl); }
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int ln_backward(const float* dh, const float* x, const float* stats, const float* gamma, float* dx_accum,
                float* dgamma, float* dbeta, float* partials, size_t partial_bytes, int64_t rows, int E, int ld,
                cudaStream_t st) {
  const int nv = cdiv(ld, 32);
  V1T_CHECK_ARG(nv <= 16 && ld >= E, "layer norm backward: emb dim %d unsupported (max 512)", E);
  int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)kNumSMs * 6);  // 48 warps per SM hide the load latency
  const int64_t max_grid = (int64_t)(partial_bytes / (2 * (size_t)E * sizeof(float)));
  V1T_CHECK_ARG(max_grid >= 1, "layer norm backward: partials workspace too small");
  if (grid > max_grid) grid = (int)max_grid;
  V1T_LN_DISPATCH(1, ln_backward_kernel, dh, x, stats, gamma, dx_accum, partials, rows, E, ld)
  V1T_LN_DISPATCH(2, ln_backward_kernel, dh, x, stats, gamma, dx_accum, partials, rows, E, ld)
  V1T_LN_DISPATCH(5, ln_backward_kernel, dh, x, stats, gamma, dx_accum, partials, rows, E, ld)
  V1T_LN_DISPATCH(8, ln_backward_kernel, dh, x, stats, gamma, dx_accum, partials, rows, E, ld)
  { ln_backward_kernel<16><<<grid, 256, 0, st>>>(dh, x, stats, gamma, dx_accum, partials, rows, E, ld); }
  V1T_LAUNCH_CHECK();
  ln_finish_kernel<<<dim3(cdiv(E, 32), 2), kLnFinishWarps * 32, 0, st>>>(partials, dgamma, dbeta, grid, E);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int softmax_rows(float* S, int64_t rows, int cols, int64_t ld, DropSpec dr, int64_t row_offset, cudaStream_t st) {
  const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)kNumSMs * 16);
  softmax_rows_kernel<<<grid, 256, 0, st>>>(S, rows, cols, ld, dr, row_offset);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
int softmax_bwd_rows(float* P, float* dP, int64_t rows, int cols, int64_t ld, DropSpec dr, int64_t row_offset,
                     cudaStream_t st) {
  const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)kNumSMs * 16);
  softmax_bwd_rows_kernel<<<grid, 256, 0, st>>>(P, dP, rows, cols, ld, dr, row_offset);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
int gelu_forward(const float* u, float* g, int64_t rows, int cols, int64_t ld, DropSpec dr, cudaStream_t st) {
  gelu_forward_kernel<<<ew_grid(rows * ld), 256, 0, st>>>(u, g, rows, cols, ld, dr);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
int gelu_backward(float* dg_inout, const float* u, int64_t rows, int cols, int64_t ld, DropSpec dr,
                  cudaStream_t st) {
  gelu_backward_kernel<<<ew_grid(rows * ld), 256, 0, st>>>(dg_inout, u, rows, cols, ld, dr);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int colsum(const float* X, float* out, int batch, int64_t rows, int cols, int64_t xb, int64_t ld, int64_t out_ld,
           float* partials, size_t partial_bytes, cudaStream_t st) {
  if (cols == 0 || batch == 0) return V1T_OK;
  int strips = (int)std::min<int64_t>(256, std::max<int64_t>(1, rows / 64));
  while (strips > 1 && (size_t)batch * strips * cols * sizeof(float) > partial_bytes) strips /= 2;
  V1T_CHECK_ARG((size_t)batch * strips * cols * sizeof(float) <= partial_bytes, "colsum: partials workspace too small");
  dim3 grid(cdiv(cols, 32), strips, batch), block(32, 8);
  colsum_kernel<<<grid, block, 0, st>>>(X, partials, rows, cols, xb, ld, strips);
  V1T_LAUNCH_CHECK();
  colsum_finish_kernel<<<dim3(cdiv(cols, 32), batch), kColsumFinishWarps * 32, 0, st>>>(partials, out, batch, cols, strips, out_ld);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int colsum_finish(const float* partials, float* out, int cols, int strips, cudaStream_t st) {
  colsum_finish_kernel<<<dim3(cdiv(cols, 32), 1), kColsumFinishWarps * 32, 0, st>>>(partials, out, 1, cols, strips, 0);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int batchsum(const float* X, float* out, int B, int64_t rows, int cols, int64_t xb, int64_t ld, int64_t out_ld,
             cudaStream_t st) {
  batchsum_kernel<<<ew_grid(rows * cols), 256, 0, st>>>(X, out, B, rows, cols, xb, ld, out_ld);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

}  // namespace v1t
