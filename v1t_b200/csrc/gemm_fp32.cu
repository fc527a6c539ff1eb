// fp32 CUDA-core GEMM with arbitrary operand strides, two batch dims, split-K and a fused
// bias/residual epilogue.  This is the V1T_IMPL_FP32 workhorse: every contraction of the ViT core
// (vit.py:143-150,253-275) and its autograd can be expressed with it; it is also the on-device
// cross-check for the tcgen05 kernels.
#include "common.cuh"
#include "kernels.cuh"

namespace v1t {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;
constexpr int kThreads = (BM / TM) * (BN / TN);  // 256

__global__ void __launch_bounds__(kThreads) sgemm_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  int z = blockIdx.z;
  const int split = z % g.splits;
  z /= g.splits;
  const int b2 = z % g.d.batch2, b1 = z / g.d.batch2;
  const float* __restrict__ A = g.A + b1 * g.d.a_b1 + b2 * g.d.a_b2;
  const float* __restrict__ B = g.B + b1 * g.d.b_b1 + b2 * g.d.b_b2;
  float* __restrict__ C = g.C + b1 * g.d.c_b1 + b2 * g.d.c_b2 + (int64_t)split * g.c_split;
  const float* __restrict__ R = g.R ? g.R + b1 * g.d.r_b1 + b2 * g.d.r_b2 : nullptr;

  const int k_begin = split * g.k_chunk;
  const int k_end = min(g.d.k, k_begin + g.k_chunk);

  const bool a_kc = (g.d.a_k == 1);   // K contiguous in A
  const bool b_nc = (g.d.b_n == 1);   // N contiguous in B
  constexpr int A_PER = BM * BK / kThreads;  // 8
  constexpr int B_PER = BN * BK / kThreads;  // 4

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[A_PER], rb[B_PER];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int mm, kk;
      if (a_kc) { kk = tid % BK; mm = tid / BK + i * (kThreads / BK); }
      else      { mm = tid % BM; kk = tid / BM + i * (kThreads / BM); }
      const int m = m0 + mm, k = k0 + kk;
      ra[i] = (m < g.d.m && k < k_end) ? __ldg(A + (int64_t)m * g.d.a_m + (int64_t)k * g.d.a_k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int nn, kk;
      if (b_nc) { nn = tid % BN; kk = tid / BN + i * (kThreads / BN); }
      else      { kk = tid % BK; nn = tid / BK + i * (kThreads / BK); }
      const int n = n0 + nn, k = k0 + kk;
      rb[i] = (n < g.d.n && k < k_end) ? __ldg(B + (int64_t)k * g.d.b_k + (int64_t)n * g.d.b_n) : 0.f;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int mm, kk;
      if (a_kc) { kk = tid % BK; mm = tid / BK + i * (kThreads / BK); }
      else      { mm = tid % BM; kk = tid / BM + i * (kThreads / BM); }
      As[kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int nn, kk;
      if (b_nc) { nn = tid % BN; kk = tid / BN + i * (kThreads / BN); }
      else      { kk = tid % BK; nn = tid / BK + i * (kThreads / BK); }
      Bs[kk][nn] = rb[i];
    }
  };

  if (k_begin < k_end) load_tiles(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < k_end) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.d.m) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.d.n) continue;
      float v = g.d.alpha * acc[i][j];
      if (g.bias) v += __ldg(g.bias + n);
      if (g.drop.p > 0.f)
        v *= dropout_mult(g.drop.seed, g.drop.site, (uint64_t)m * drop_stride(g.d.n) + n, g.drop.p, 1.f / (1.f - g.drop.p));
      if (R) v += __ldg(R + (int64_t)m * g.d.r_m + n);
      float* c = C + (int64_t)m * g.d.c_m + n;
      if (g.d.accumulate) v += *c;
      *c = v;
    }
  }
}

// out[i] (+)= sum_p partials[p*n + i]   (deterministic split-K / column-sum finish)
__global__ void reduce_partials_kernel(const float* __restrict__ partials, float* __restrict__ out, int64_t n,
                                       int parts, int64_t rows, int64_t cols, int64_t ld_out, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < parts; ++p) s += partials[(int64_t)p * n + i];
  const int64_t r = i / cols, c = i % cols;
  (void)rows;
  float* o = out + r * ld_out + c;
  *o = accumulate ? *o + s : s;
}

// un-pad a grouped index: groups of `gout` (padded) hold `gin` real entries; -1 for padding
__device__ __forceinline__ int64_t ungroup(int64_t i, int gin, int gout) {
  if (gout <= 0) return i;
  return (i % gout < gin) ? (i / gout) * gin + i % gout : -1;
}
// block (64, 4): threadIdx.x = consecutive outputs (coalesced rows of every partial), the 4 y-lanes take every 4th
// partial so that 4x as many loads are in flight; fixed-order smem reduction over y (deterministic)
__global__ void __launch_bounds__(256) reduce_partials_ld_kernel(const float* __restrict__ partials, float* __restrict__ out,
                                                                 int parts, int64_t rows, int64_t cols, int64_t in_ld,
                                                                 int64_t out_ld, int accumulate, GroupMap gm) {
  __shared__ float red[4][64];
  const int64_t i = (int64_t)blockIdx.x * 64 + threadIdx.x;
  const bool in = i < rows * cols;
  const int64_t r = in ? i / cols : 0, c = in ? i % cols : 0;
  float s = 0.f;
  if (in) {
    const float* pp = partials + r * in_ld + c;
    const int64_t ps = rows * in_ld;
    int p = threadIdx.y;
    for (; p + 12 < parts; p += 16) {  // four loads in flight per thread
      const float v0 = pp[p * ps], v1 = pp[(p + 4) * ps], v2 = pp[(p + 8) * ps], v3 = pp[(p + 12) * ps];
      s += (v0 + v1) + (v2 + v3);
    }
    for (; p < parts; p += 4) s += pp[p * ps];
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y != 0 || !in) return;
  const int64_t ro = ungroup(r, gm.row_gin, gm.row_gout), co = ungroup(c, gm.col_gin, gm.col_gout);
  if (ro < 0 || co < 0) return;  // padded row / column of a head-padded product
  const float t = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
  float* o = out + ro * out_ld + co;
  *o = accumulate ? *o + t : t;
}

}  // namespace

int reduce_partials_ld(const float* partials, float* out, int parts, int64_t rows, int64_t cols, int64_t in_ld,
                       int64_t out_ld, int accumulate, cudaStream_t st, GroupMap gm) {
  if (rows * cols == 0) return V1T_OK;
  reduce_partials_ld_kernel<<<cdiv(rows * cols, 64), dim3(64, 4), 0, st>>>(partials, out, parts, rows, cols, in_ld,
                                                                           out_ld, accumulate, gm);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int reduce_partials(const float* partials, float* out, int parts, int64_t rows, int64_t cols, int64_t ld_out,
                    int accumulate, cudaStream_t st) {
  const int64_t n = rows * cols;
  if (n == 0) return V1T_OK;
  reduce_partials_kernel<<<cdiv(n, 256), 256, 0, st>>>(partials, out, n, parts, rows, cols, ld_out, accumulate);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

int gemm_fp32(const v1t_gemm_desc& d, const float* A, const float* B, float* C, const float* bias, const float* R,
              cudaStream_t st, DropSpec drop) {
  V1T_CHECK_ARG(d.m >= 0 && d.n >= 0 && d.k >= 0 && d.batch1 >= 1 && d.batch2 >= 1, "gemm: bad sizes");
  if (d.m == 0 || d.n == 0) return V1T_OK;
  GemmArgs g;
  g.d = d;
  g.A = A; g.B = B; g.C = C; g.bias = bias; g.R = R;
  g.drop = drop;
  g.epi = no_epi();
  g.splits = 1; g.k_chunk = d.k; g.c_split = 0;
  dim3 grid(cdiv(d.n, BN), cdiv(d.m, BM), d.batch1 * d.batch2);
  V1T_CHECK_ARG(grid.z <= 65535 && grid.y <= 65535, "gemm: grid too large");
  sgemm_kernel<<<grid, kThreads, 0, st>>>(g);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

// Split-K GEMM for the weight gradients (tiny M x N, huge K = B*T rows): partial products go to
// `partials` [splits, m, n] and are summed in a fixed order -> bitwise deterministic.
int gemm_fp32_splitk(const v1t_gemm_desc& d, const float* A, const float* B, float* C, float* partials,
                     size_t partial_bytes, cudaStream_t st) {
  V1T_CHECK_ARG(d.batch1 == 1 && d.batch2 == 1, "splitk gemm: no batch");
  if (d.m == 0 || d.n == 0) return V1T_OK;
  const int tiles = cdiv(d.n, BN) * cdiv(d.m, BM);
  int splits = (4 * kNumSMs + tiles - 1) / tiles;
  const int max_by_k = cdiv(d.k, 4 * BK);
  if (splits > max_by_k) splits = max_by_k;
  const int64_t per = (int64_t)d.m * d.n * sizeof(float);
  if ((int64_t)splits * per > (int64_t)partial_bytes) splits = (int)(partial_bytes / per);
  if (splits < 1) splits = 1;
  if (splits == 1) return gemm_fp32(d, A, B, C, nullptr, nullptr, st);
  int k_chunk = (int)round_up(cdiv(d.k, splits), BK);
  splits = cdiv(d.k, k_chunk);
  GemmArgs g;
  g.d = d;
  g.d.accumulate = 0;
  g.d.c_m = d.n;
  g.A = A; g.B = B; g.C = partials; g.bias = nullptr; g.R = nullptr;
  g.drop = no_drop();
  g.epi = no_epi();
  g.splits = splits; g.k_chunk = k_chunk; g.c_split = (int64_t)d.m * d.n;
  dim3 grid(cdiv(d.n, BN), cdiv(d.m, BM), splits);
  sgemm_kernel<<<grid, kThreads, 0, st>>>(g);
  V1T_LAUNCH_CHECK();
  return reduce_partials(partials, C, splits, d.m, d.n, d.c_m, d.accumulate, st);
}

}  // namespace v1t

extern "C" int v1t_gemm_fp32(const v1t_gemm_desc* d, const float* A, const float* B, float* C, const float* bias,
                             const float* R, void* stream) {
  V1T_CHECK_ARG(d && A && B && C, "v1t_gemm_fp32: null argument");
  return v1t::gemm_fp32(*d, A, B, C, bias, R, (cudaStream_t)stream);
}
