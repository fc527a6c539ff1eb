// "UMMA-ready" bf16 operand planes in HBM for the fused attention kernels.
//
// The attention operands (Q, K, V, dO per head) are re-read by many CTAs, so they are converted ONCE from the
// fp32 GEMM outputs into bf16 hi (+ lo) planes laid out exactly like the shared-memory tiles tcgen05.mma reads:
// 32-element (64-byte) K-atoms, rows 64 B apart, 16-byte chunks XOR-swizzled with ((row >> 1) & 3) — the
// canonical K-major SWIZZLE_64B layout.  A tile of any 8-row-aligned row range of one atom is then a CONTIGUOUS
// byte range, fetched with a single cp.async.bulk (no tensor map, no conversion work in the consumer).
//
//   RM plane  rows = tokens, K = head dim   [B*H][Dp/32 atoms][Tp rows][64 B]     (Q, K, V, dO as A/B operands
//                                                                                  contracted over d)
//   TR plane  rows = head dim, K = tokens   [B*H][Tp/32 atoms][Dp rows][64 B]     (V^T, Q^T, K^T, dO^T: operands
//                                                                                  contracted over tokens)
// Pad rows / columns (t >= T, d >= E) are written as zeros.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace v1t {
namespace {

using namespace tc;

// grid (Tp/32, B*H); block 256.  Each block converts 32 tokens x Dp dims of one (b, h).
__global__ void __launch_bounds__(256) make_planes_kernel(const float* __restrict__ X, int64_t ld, int col0, int B,
                                                          int H, int T, int Tp, int E, int Dp,
                                                          uint8_t* __restrict__ rm_hi, uint8_t* __restrict__ rm_lo,
                                                          uint8_t* __restrict__ tr_hi, uint8_t* __restrict__ tr_lo) {
  extern __shared__ float tile[];  // [32][Dp + 1]
  const int tb = blockIdx.x, bh = blockIdx.y;
  const int b = bh / H, h = bh % H;
  const int t0 = tb * 32;
  const int pitch = Dp + 1;
  for (int i = threadIdx.x; i < 32 * Dp; i += blockDim.x) {
    const int tl = i / Dp, d = i % Dp;
    const int t = t0 + tl;
    float v = 0.f;
    if (t < T && d < E) v = __ldg(X + ((int64_t)b * T + t) * ld + col0 + h * E + d);
    tile[tl * pitch + d] = v;
  }
  __syncthreads();
  const int atoms_d = Dp / 32, atoms_t = Tp / 32;
  // RM: one 16-byte chunk = 8 consecutive d of one token
  if (rm_hi) {
    for (int i = threadIdx.x; i < 32 * (Dp / 8); i += blockDim.x) {
      const int tl = i % 32, ch = i / 32;  // lanes = consecutive tokens
      const int a = ch / 4, c = ch % 4;
      const int t = t0 + tl;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = tile[tl * pitch + a * 32 + c * 8 + e];
      uint4 hi, lo;
      split8(v, hi, lo);
      const int64_t off = attn_plane_off(bh, a, t, Tp, atoms_d) + ((c ^ ((t >> 1) & 3)) << 4);
      *reinterpret_cast<uint4*>(rm_hi + off) = hi;
      if (rm_lo) *reinterpret_cast<uint4*>(rm_lo + off) = lo;
    }
  }
  // TR: one 16-byte chunk = 8 consecutive tokens of one d
  if (tr_hi) {
    const int ka = tb;  // this block is exactly one 32-token atom
    for (int i = threadIdx.x; i < Dp * 4; i += blockDim.x) {
      const int d = i / 4, c = i % 4;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = tile[(c * 8 + e) * pitch + d];
      uint4 hi, lo;
      split8(v, hi, lo);
      const int64_t off = (((int64_t)bh * atoms_t + ka) * Dp + d) * 64 + ((c ^ ((d >> 1) & 3)) << 4);
      *reinterpret_cast<uint4*>(tr_hi + off) = hi;
      if (tr_lo) *reinterpret_cast<uint4*>(tr_lo + off) = lo;
    }
  }
}

// ---- materialised attention on tensor cores with plane operands (head dim > 160) -------------------------------------
// per-(sample, head) matrix planes [bh][Dp/32 atoms][Tq rows][64 B]: one thread per 16-byte chunk; pad rows / columns zero
__global__ void bh_planes_kernel(const float* __restrict__ X, int64_t ld, int col0, int H, int T, int Tq, int E, int Dp,
                                 int64_t chunks, uint8_t* __restrict__ hi, uint8_t* __restrict__ lo) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= chunks) return;
  const int c = (int)(idx & 3);
  int64_t q = idx >> 2;
  const int r = (int)(q % Tq); q /= Tq;
  const int atoms = Dp / 32;
  const int a = (int)(q % atoms);
  const int64_t bh = q / atoms;
  const int b = (int)(bh / H), h = (int)(bh % H);
  const int d0 = a * 32 + c * 8;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e)
    v[e] = (r < T && d0 + e < E) ? __ldg(X + ((int64_t)b * T + r) * ld + col0 + h * E + d0 + e) : 0.f;
  uint4 hv, lv;
  split8(v, hv, lv);
  const int64_t off = ((bh * atoms + a) * Tq + r) * 64 + ((c ^ ((r >> 1) & 3)) << 4);
  *reinterpret_cast<uint4*>(hi + off) = hv;
  if (lo) *reinterpret_cast<uint4*>(lo + off) = lv;
}

// One warp per row of a T x T score matrix; the row lives in registers (kCh chunks of 8 columns per lane).
constexpr int kRowChunks = 8;  // per lane: Tq <= 32 * 8 * 8 = 2048 columns
struct RowRegs { float v[kRowChunks][8]; };
__device__ __forceinline__ void load_row(const float* __restrict__ row, int T, int nch, int lane, RowRegs& x, float fill) {
#pragma unroll
  for (int i = 0; i < kRowChunks; ++i) {
    const int ch = lane + 32 * i, c0 = ch * 8;
    if (ch < nch && c0 + 8 <= T) {  // rows start 16-byte aligned (ld % 4 == 0)
      const float4 lo4 = __ldg(reinterpret_cast<const float4*>(row + c0)), hi4 = __ldg(reinterpret_cast<const float4*>(row + c0 + 4));
      x.v[i][0] = lo4.x; x.v[i][1] = lo4.y; x.v[i][2] = lo4.z; x.v[i][3] = lo4.w;
      x.v[i][4] = hi4.x; x.v[i][5] = hi4.y; x.v[i][6] = hi4.z; x.v[i][7] = hi4.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) x.v[i][e] = (ch < nch && c0 + e < T) ? __ldg(row + c0 + e) : fill;
    }
  }
}
// writes this lane's chunks of row t of matrix `mat` into planes [mat][Tq/32 atoms][Tq rows][64 B]
__device__ __forceinline__ void store_row_planes(const RowRegs& x, int64_t mat, int t, int Tq, int nch, int lane,
                                                 uint8_t* __restrict__ hi, uint8_t* __restrict__ lo) {
  const int atoms = Tq / 32;
#pragma unroll
  for (int i = 0; i < kRowChunks; ++i) {
    const int ch = lane + 32 * i;
    if (ch < nch) {
      uint4 hv, lv;
      split8(x.v[i], hv, lv);
      const int64_t off = ((mat * atoms + (ch >> 2)) * Tq + t) * 64 + (((ch & 3) ^ ((t >> 1) & 3)) << 4);
      *reinterpret_cast<uint4*>(hi + off) = hv;
      if (lo) *reinterpret_cast<uint4*>(lo + off) = lv;
    }
  }
}
__device__ __forceinline__ void softmax_in_regs(RowRegs& x, int T, int nch, int lane) {
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < kRowChunks; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) m = fmaxf(m, x.v[i][e]);  // columns >= T were loaded as -inf
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kRowChunks; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      x.v[i][e] = __expf(x.v[i][e] - m);
      s += x.v[i][e];
    }
  s = warp_sum(s);
  const float inv = 1.f / s;
#pragma unroll
  for (int i = 0; i < kRowChunks; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) x.v[i][e] *= inv;
}
__device__ __forceinline__ void row_mask(float (&mk)[kRowChunks][8], DropSpec dr, int64_t grow, int T, int nch, int lane) {
  const float inv_keep = dr.p > 0.f ? 1.f / (1.f - dr.p) : 1.f;
  const int64_t Tc = drop_stride(T);
#pragma unroll
  for (int i = 0; i < kRowChunks; ++i) {
    const int ch = lane + 32 * i;
    if (dr.p > 0.f && ch < nch && ch * 8 < T) dropout_mult8(dr.seed, dr.site, (uint64_t)(grow * Tc + ch * 8) >> 3, dr.p, inv_keep, mk[i]);
    else {
#pragma unroll
      for (int e = 0; e < 8; ++e) mk[i][e] = 1.f;
    }
  }
}
// rows: mats * Tq (pad rows t >= T are written as zero rows so that contractions over the rows see zeros)
__global__ void __launch_bounds__(256) softmax_rows_planes_kernel(const float* __restrict__ S, int64_t mats, int T, int Tq, int64_t ld,
                                                                  DropSpec dr, int64_t row_offset, uint8_t* __restrict__ p_hi,
                                                                  uint8_t* __restrict__ p_lo) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= mats * Tq) return;
  const int64_t mat = w / Tq;
  const int t = (int)(w % Tq), nch = Tq / 8;
  RowRegs x;
  if (t < T) {
    load_row(S + (mat * T + t) * ld, T, nch, lane, x, -INFINITY);
    softmax_in_regs(x, T, nch, lane);
    float mk[kRowChunks][8];
    row_mask(mk, dr, row_offset + mat * T + t, T, nch, lane);
#pragma unroll
    for (int i = 0; i < kRowChunks; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) x.v[i][e] *= mk[i][e];
  } else {
#pragma unroll
    for (int i = 0; i < kRowChunks; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) x.v[i][e] = 0.f;
  }
  store_row_planes(x, mat, t, Tq, nch, lane, p_hi, p_lo);
}
__global__ void __launch_bounds__(256) softmax_bwd_rows_planes_kernel(const float* __restrict__ S, const float* __restrict__ dPd,
                                                                      int64_t mats, int T, int Tq, int64_t ld, DropSpec dr,
                                                                      int64_t row_offset, uint8_t* __restrict__ pd_hi,
                                                                      uint8_t* __restrict__ pd_lo, uint8_t* __restrict__ ds_hi,
                                                                      uint8_t* __restrict__ ds_lo) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= mats * Tq) return;
  const int64_t mat = w / Tq;
  const int t = (int)(w % Tq), nch = Tq / 8;
  RowRegs p, g;
  if (t < T) {
    load_row(S + (mat * T + t) * ld, T, nch, lane, p, -INFINITY);
    softmax_in_regs(p, T, nch, lane);
    load_row(dPd + (mat * T + t) * ld, T, nch, lane, g, 0.f);
    float mk[kRowChunks][8];
    row_mask(mk, dr, row_offset + mat * T + t, T, nch, lane);
    float delta = 0.f;
#pragma unroll
    for (int i = 0; i < kRowChunks; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        g.v[i][e] *= mk[i][e];                 // dP = dPd * mask
        delta = fmaf(p.v[i][e], g.v[i][e], delta);
      }
    delta = warp_sum(delta);
#pragma unroll
    for (int i = 0; i < kRowChunks; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        g.v[i][e] = p.v[i][e] * (g.v[i][e] - delta);  // dS
        p.v[i][e] *= mk[i][e];                        // Pd
      }
  } else {
#pragma unroll
    for (int i = 0; i < kRowChunks; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) { p.v[i][e] = 0.f; g.v[i][e] = 0.f; }
  }
  store_row_planes(p, mat, t, Tq, nch, lane, pd_hi, pd_lo);
  store_row_planes(g, mat, t, Tq, nch, lane, ds_hi, ds_lo);
}

// GEMM-operand planes of a row-major matrix X[rows, cols]: one thread per 16-byte chunk (8 columns of one row)
__device__ __forceinline__ void matrix_planes_chunk(int64_t idx, const float* __restrict__ X, int64_t ld, int64_t rows,
                                                    int64_t cols, int64_t rows_p, uint8_t* __restrict__ hi,
                                                    uint8_t* __restrict__ lo, int gin, int gout, int cgin, int cgout) {
  const int c = (int)(idx & 3);
  const int64_t r = (idx >> 2) % rows_p, a = (idx >> 2) / rows_p;
  const int64_t col0 = a * 32 + c * 8;
  int64_t sr = r;  // source row (-1: padding)
  if (gout > 0) sr = (r % gout < gin) ? (r / gout) * gin + r % gout : -1;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    int64_t sc = col0 + e;  // source column (-1: padding)
    if (cgout > 0) sc = (sc % cgout < cgin) ? (sc / cgout) * cgin + sc % cgout : -1;
    v[e] = (sr >= 0 && sr < rows && sc >= 0 && sc < cols) ? __ldg(X + sr * ld + sc) : 0.f;
  }
  uint4 h, l;
  split8(v, h, l);
  const int64_t off = (a * rows_p + r) * 64 + ((c ^ (int)((r >> 1) & 3)) << 4);
  *reinterpret_cast<uint4*>(hi + off) = h;
  if (lo) *reinterpret_cast<uint4*>(lo + off) = l;
}
__global__ void matrix_planes_kernel(const float* __restrict__ X, int64_t ld, int64_t rows, int64_t cols,
                                     int64_t rows_p, int64_t chunks, uint8_t* __restrict__ hi,
                                     uint8_t* __restrict__ lo, int gin, int gout, int cgin, int cgout) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < chunks) matrix_planes_chunk(idx, X, ld, rows, cols, rows_p, hi, lo, gin, gout, cgin, cgout);
}
// several matrices in one launch (the weights of a block): job j owns chunk indices [first[j], first[j+1])
__global__ void matrix_planes_batch_kernel(const PlaneJobs jobs) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= jobs.first[jobs.n]) return;
  int j = 0;
  while (idx >= jobs.first[j + 1]) ++j;
  const PlaneJob& q = jobs.job[j];
  matrix_planes_chunk(idx - jobs.first[j], q.X, q.ld, q.rows, q.cols, q.rows_p, q.hi, q.lo, q.row_gin, q.row_gout,
                      q.col_gin, q.col_gout);
}

struct PadPlanes { uint8_t* p[6]; };
// zero the pad rows t in [T, Tq) of every slab of each plane.  attn_ad == 0: matrix planes, slab = column atom (Tq
// rows of 64 bytes each);  attn_ad > 0: attention planes, slab = (sample*head) * attn_ad + atom (attn_plane_off)
__global__ void zero_pad_rows_kernel(PadPlanes pp, int n_planes, int64_t slabs, int Tq, int T, int attn_ad) {
  const int pad = Tq - T;
  const int64_t total = slabs * pad * 4;  // 16-byte pieces per plane
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int piece = (int)(i & 3);
    const int64_t row = (i >> 2) % pad, slab = (i >> 2) / pad;
    const int64_t off = (attn_ad > 0 ? tc::attn_plane_off(slab / attn_ad, (int)(slab % attn_ad), T + (int)row, Tq, attn_ad)
                                     : (slab * Tq + T + row) * 64) + piece * 16;
    for (int p = 0; p < n_planes; ++p) *reinterpret_cast<uint4*>(pp.p[p] + off) = make_uint4(0, 0, 0, 0);
  }
}

// qkv[(b*T+t), (s*H+h)*E + d] = hi + lo of the attention planes (attention-map hooks only)
__global__ void planes_to_qkv_kernel(HeadPlanes hp, int B, int E, float* __restrict__ qkv) {
  const int64_t total = (int64_t)B * hp.T * 3 * hp.H * E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % E);
    int64_t r = i / E;
    const int h = (int)(r % hp.H); r /= hp.H;
    const int s = (int)(r % 3); r /= 3;
    const int t = (int)(r % hp.T), b = (int)(r / hp.T);
    const int64_t off = tc::attn_plane_off((int64_t)b * hp.H + h, d / 32, t, hp.Tq, hp.AD) +
                        ((((d & 31) >> 3) ^ ((t >> 1) & 3)) << 4) + (d & 7) * 2;
    float v = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(hp.p[s][0] + off));
    if (hp.p[s][1]) v += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(hp.p[s][1] + off));
    qkv[i] = v;
  }
}

// delta[b,h,t] = sum_d O[b,t,h*E+d] * dO[b,t,h*E+d]   (softmax backward row term); one warp per (b,t,h)
__global__ void attn_delta_kernel(const float* __restrict__ O, const float* __restrict__ dO, float* __restrict__ delta,
                                  int B, int H, int T, int Tp, int E, int64_t ld) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t total = (int64_t)B * T * H;
  if (w >= total) return;
  const int h = (int)(w % H);
  const int64_t r = w / H;  // b*T + t
  const int b = (int)(r / T), t = (int)(r % T);
  const float* o = O + r * ld + h * E;
  const float* g = dO + r * ld + h * E;
  float s = 0.f;
  for (int d = lane; d < E; d += 32) s = fmaf(o[d], g[d], s);
  s = warp_sum(s);
  if (lane == 0) delta[((int64_t)b * H + h) * Tp + t] = s;
}

}  // namespace

size_t plane_bytes(int B, int H, int Tp, int Dp) { return (size_t)B * H * Tp * Dp * 2; }

size_t matrix_plane_bytes(int64_t rows, int64_t cols) {
  return (size_t)cdiv(cols, 32) * (size_t)round_up(rows, 32) * 64;
}

int matrix_planes(const float* X, int64_t ld, int64_t rows, int64_t cols, void* hi, void* lo, PlaneOp* out,
                  cudaStream_t st, int row_gin, int row_gout, int col_gin, int col_gout) {
  V1T_CHECK_ARG(row_gout == 0 || (row_gin > 0 && row_gout >= row_gin && rows % row_gin == 0),
                "matrix_planes: bad row grouping");
  V1T_CHECK_ARG(col_gout == 0 || (col_gin > 0 && col_gout >= col_gin && cols % col_gin == 0),
                "matrix_planes: bad column grouping");
  const int64_t prow = row_gout ? rows / row_gin * row_gout : rows;
  const int64_t pcol = col_gout ? cols / col_gin * col_gout : cols;
  const int64_t rows_p = round_up(prow, 32), catoms = cdiv(pcol, 32);
  const int64_t chunks = catoms * rows_p * 4;
  V1T_CHECK_ARG(X && hi && out && catoms * rows_p < (1ll << 31), "matrix_planes: bad argument");
  matrix_planes_kernel<<<(unsigned)cdiv(chunks, 256), 256, 0, st>>>(X, ld, rows, cols, rows_p, chunks, (uint8_t*)hi,
                                                                      (uint8_t*)lo, row_gin, row_gout, col_gin, col_gout);
  V1T_LAUNCH_CHECK();
  out->hi = (const uint8_t*)hi; out->lo = (const uint8_t*)lo; out->rows_p = (int)rows_p; out->catoms = (int)catoms; out->batch_bytes = 0;
  return V1T_OK;
}

int matrix_planes_batch(PlaneJobs& jobs, cudaStream_t st) {
  V1T_CHECK_ARG(jobs.n >= 0 && jobs.n <= PlaneJobs::kMax, "matrix_planes_batch: too many jobs");
  jobs.first[0] = 0;
  for (int j = 0; j < jobs.n; ++j) {
    PlaneJob& q = jobs.job[j];
    const int64_t prow = q.row_gout ? q.rows / q.row_gin * q.row_gout : q.rows;
    const int64_t pcol = q.col_gout ? q.cols / q.col_gin * q.col_gout : q.cols;
    q.rows_p = round_up(prow, 32);
    jobs.first[j + 1] = jobs.first[j] + (int64_t)cdiv(pcol, 32) * q.rows_p * 4;
  }
  if (jobs.first[jobs.n] == 0) return V1T_OK;
  matrix_planes_batch_kernel<<<(unsigned)cdiv(jobs.first[jobs.n], 256), 256, 0, st>>>(jobs);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int zero_plane_pad_rows(uint8_t* const* planes, int n_planes, int64_t slabs, int Tq, int T, cudaStream_t st, int attn_ad) {
  V1T_CHECK_ARG(n_planes >= 0 && n_planes <= 6 && Tq >= T, "zero_plane_pad_rows: bad argument");
  if (n_planes == 0 || Tq == T) return V1T_OK;
  PadPlanes pp{};
  for (int i = 0; i < n_planes; ++i) pp.p[i] = planes[i];
  const int64_t total = slabs * (Tq - T) * 4;
  zero_pad_rows_kernel<<<(unsigned)std::min<int64_t>(cdiv(total, 256), 4096), 256, 0, st>>>(pp, n_planes, slabs, Tq, T,
                                                                                            attn_ad);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int planes_to_qkv(const HeadPlanes& hp, int B, int E, float* qkv, cudaStream_t st) {
  const int64_t total = (int64_t)B * hp.T * 3 * hp.H * E;
  planes_to_qkv_kernel<<<(unsigned)std::min<int64_t>(cdiv(total, 256), 65535), 256, 0, st>>>(hp, B, E, qkv);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int make_planes(const float* X, int64_t ld, int col0, int B, int H, int T, int Tp, int E, int Dp, void* rm_hi,
                void* rm_lo, void* tr_hi, void* tr_lo, cudaStream_t st) {
  V1T_CHECK_ARG(Tp % 128 == 0 && Dp % 32 == 0 && Tp >= T && Dp >= E, "make_planes: bad padded sizes");
  dim3 grid(Tp / 32, B * H);
  V1T_CHECK_ARG(grid.y <= 65535, "make_planes: too many (batch, head) pairs");
  const size_t smem = sizeof(float) * 32 * (Dp + 1);
  make_planes_kernel<<<grid, 256, smem, st>>>(X, ld, col0, B, H, T, Tp, E, Dp, (uint8_t*)rm_hi, (uint8_t*)rm_lo,
                                              (uint8_t*)tr_hi, (uint8_t*)tr_lo);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

// delta from the planes alone: O as operand planes of the head-padded [B*T, H*Dp] matrix (attention forward
// epilogue), dO as attention planes ([b*H+h][atoms][Tp][64 B]); one warp per (b, t, h), lane = column within an atom
__global__ void attn_delta_planes_kernel(PlaneOp o, const uint8_t* __restrict__ do_hi, const uint8_t* __restrict__ do_lo,
                                         float* __restrict__ delta, int B, int H, int T, int Tp, int AD) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)B * T * H) return;
  const int h = (int)(w % H);
  const int64_t r = w / H;  // b*T + t
  const int b = (int)(r / T), t = (int)(r % T);
  float s = 0.f;
  if (lane < AD * 4) {  // one 16-byte chunk (8 head-dim columns) per lane, 128-bit loads from all four planes
    const int a = lane >> 2, ch = lane & 3;
    const int64_t oo = tc::plane_chunk_off(h * AD + a, o.rows_p, r, ch);
    const int64_t go = tc::attn_plane_off((int64_t)b * H + h, a, t, Tp, AD) + ((ch ^ ((t >> 1) & 3)) << 4);
    const uint4 z = make_uint4(0, 0, 0, 0);
    const uint4 oh = __ldg(reinterpret_cast<const uint4*>(o.hi + oo));
    const uint4 ol = o.lo ? __ldg(reinterpret_cast<const uint4*>(o.lo + oo)) : z;
    const uint4 gh = __ldg(reinterpret_cast<const uint4*>(do_hi + go));
    const uint4 gl = do_lo ? __ldg(reinterpret_cast<const uint4*>(do_lo + go)) : z;
    const uint32_t ohw[4] = {oh.x, oh.y, oh.z, oh.w}, olw[4] = {ol.x, ol.y, ol.z, ol.w};
    const uint32_t ghw[4] = {gh.x, gh.y, gh.z, gh.w}, glw[4] = {gl.x, gl.y, gl.z, gl.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {  // a bf16 is the upper half of the fp32 with the same value
      const float o0 = __uint_as_float(ohw[e] << 16) + __uint_as_float(olw[e] << 16);
      const float o1 = __uint_as_float(ohw[e] & 0xffff0000u) + __uint_as_float(olw[e] & 0xffff0000u);
      const float g0 = __uint_as_float(ghw[e] << 16) + __uint_as_float(glw[e] << 16);
      const float g1 = __uint_as_float(ghw[e] & 0xffff0000u) + __uint_as_float(glw[e] & 0xffff0000u);
      s = fmaf(o0, g0, s);
      s = fmaf(o1, g1, s);
    }
  }
  s = warp_sum(s);
  if (lane == 0) delta[((int64_t)b * H + h) * Tp + t] = s;
}

int attn_delta_planes(const PlaneOp& o, const void* do_hi, const void* do_lo, float* delta, int B, int H, int T, int Tp,
                      int AD, cudaStream_t st) {
  const int64_t warps = (int64_t)B * T * H;
  V1T_CUDA(cudaMemsetAsync(delta, 0, sizeof(float) * (size_t)B * H * Tp, st));  // zero the padded rows
  attn_delta_planes_kernel<<<cdiv(warps, 8), 256, 0, st>>>(o, (const uint8_t*)do_hi, (const uint8_t*)do_lo, delta, B, H, T,
                                                          Tp, AD);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int attn_delta(const float* O, const float* dO, float* delta, int B, int H, int T, int Tp, int E, int64_t ld,
               cudaStream_t st) {
  const int64_t warps = (int64_t)B * T * H;
  V1T_CUDA(cudaMemsetAsync(delta, 0, sizeof(float) * (size_t)B * H * Tp, st));  // zero the padded rows
  attn_delta_kernel<<<cdiv(warps, 8), 256, 0, st>>>(O, dO, delta, B, H, T, Tp, E, ld);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}


size_t bh_plane_bytes(int B, int H, int Tq, int Dp) { return (size_t)B * H * (size_t)(Dp / 32) * Tq * 64; }
int bh_planes(const float* X, int64_t ld, int col0, int B, int H, int T, int Tq, int E, int Dp, void* hi, void* lo,
              cudaStream_t st) {
  V1T_CHECK_ARG(X && hi && Dp % 32 == 0 && Tq % 32 == 0 && Tq >= T && E <= Dp, "bh_planes: bad argument");
  const int64_t chunks = (int64_t)B * H * (Dp / 32) * Tq * 4;
  bh_planes_kernel<<<cdiv(chunks, 256), 256, 0, st>>>(X, ld, col0, H, T, Tq, E, Dp, chunks, (uint8_t*)hi, (uint8_t*)lo);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
size_t prob_plane_bytes(int64_t mats, int Tq) { return (size_t)mats * (size_t)(Tq / 32) * Tq * 64; }
int softmax_rows_planes(const float* S, int64_t mats, int T, int Tq, int64_t ld, DropSpec dr, int64_t row_offset, void* p_hi,
                        void* p_lo, cudaStream_t st) {
  V1T_CHECK_ARG(S && p_hi && Tq % 32 == 0 && Tq >= T && Tq <= 32 * 8 * kRowChunks && ld % 4 == 0, "softmax_rows_planes: T up to %d",
                32 * 8 * kRowChunks);
  softmax_rows_planes_kernel<<<cdiv(mats * Tq, 8), 256, 0, st>>>(S, mats, T, Tq, ld, dr, row_offset, (uint8_t*)p_hi,
                                                                 (uint8_t*)p_lo);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
int softmax_bwd_rows_planes(const float* S, const float* dPd, int64_t mats, int T, int Tq, int64_t ld, DropSpec dr,
                            int64_t row_offset, void* pd_hi, void* pd_lo, void* ds_hi, void* ds_lo, cudaStream_t st) {
  V1T_CHECK_ARG(S && dPd && pd_hi && ds_hi && Tq % 32 == 0 && Tq >= T && Tq <= 32 * 8 * kRowChunks && ld % 4 == 0,
                "softmax_bwd_rows_planes: T up to %d", 32 * 8 * kRowChunks);
  softmax_bwd_rows_planes_kernel<<<cdiv(mats * Tq, 8), 256, 0, st>>>(S, dPd, mats, T, Tq, ld, dr, row_offset, (uint8_t*)pd_hi,
                                                                     (uint8_t*)pd_lo, (uint8_t*)ds_hi, (uint8_t*)ds_lo);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

}  // namespace v1t
