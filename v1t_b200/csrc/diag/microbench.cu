// Measurement helper (not on the product path): cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N and
// of the operand source (A from shared memory "SS" vs A from tensor memory "TS").  Used to size the attention /
// GEMM tiles ("MMA cost model").
#include <stdarg.h>

#include "../common.cuh"
#include "../tc_common.cuh"
#include "../../../include/v1t_b200_diag.h"

// libv1t_b200_diag.so is self-contained (it is NOT part of the product library): local error plumbing
namespace v1t {
static thread_local char g_diag_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_diag_err, sizeof(g_diag_err), fmt, ap);
  va_end(ap);
}
void count_launch() {}
}  // namespace v1t
extern "C" const char* v1t_diag_last_error(void) { return v1t::g_diag_err; }

namespace v1t {
namespace {
using namespace tc;

__global__ void __launch_bounds__(128, 1) mma_microbench_kernel(int N, int ts, int iters, int mn_b, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = idesc_bf16(128, N, 0, mn_b);
    const uint64_t da = kDescK64 | (smem_u32(smem) >> 4);
    const uint64_t db = (mn_b ? desc_mn_sw64_base(2048) : kDescK64) | (smem_u32(smem + 32768) >> 4);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {  // rep 0 = warm-up
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (leader) {
          // 8 distinct k-steps like a real main loop (A: 4 atoms x 2 halves)
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t ao = (ks >> 1) * 512 + (ks & 1) * 2, bo = mn_b ? ks * 64 : (ks >> 1) * 1024 + (ks & 1) * 2;
            if (ts) umma_bf16_ts(tmem + 256, tmem + ks * 8, db + bo, idesc, 1u);
            else umma_bf16(tmem + 256, da + ao, db + bo, idesc, 1u);
          }
        }
      }
      if (leader) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (leader) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// Self-test of the TS form (A operand in tensor memory): C[128 x N] = A[128 x K] B[N x K]^T with A written to TMEM
// by tcgen05.st (thread = row, two bf16 per 32-bit column), B staged K-major SWIZZLE_64B in shared memory.
__global__ void __launch_bounds__(128, 1) ts_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             float* __restrict__ C, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = threadIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  // A row -> TMEM columns [256, 256 + K/2): 8 columns (16 bf16) per tcgen05.st
  for (int k0 = 0; k0 < K; k0 += 16) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __nv_bfloat162 v = __floats2bfloat162_rn(A[row * K + k0 + 2 * i], A[row * K + k0 + 2 * i + 1]);
      r[i] = *reinterpret_cast<const uint32_t*>(&v);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_off + 256 + k0 / 2),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  }
  tmem_st_wait();
  // B -> smem, K-major SW64: atoms of 32 k, each [N rows][64 B]
  for (int i = threadIdx.x; i < N * (K / 8); i += blockDim.x) {
    const int n = i % N, ch = i / N;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = B[n * K + ch * 8 + e];
    uint4 hi, lo;
    split8(x, hi, lo);
    *reinterpret_cast<uint4*>(smem + (ch >> 2) * (N * 64) + sw64_offset(n, ch & 3)) = hi;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = idesc_bf16(128, N, 0, 0);
    const uint64_t db = kDescK64 | (smem_u32(smem) >> 4);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t bo = (ks >> 1) * (N * 64 / 16) + (ks & 1) * 2;
      if (leader) umma_bf16_ts(tmem, tmem + 256 + ks * 8, db + bo, idesc, ks > 0 ? 1u : 0u);
    }
    if (leader) umma_commit(&bar);
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + lane_off + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 16; ++c) C[row * N + c0 + c] = __uint_as_float(v[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
}  // namespace
}  // namespace v1t

extern "C" int v1t_ts_selftest(const float* A, const float* B, float* C, int N, int K, void* stream) {
  using namespace v1t;
  V1T_CHECK_ARG(A && B && C && N % 16 == 0 && N >= 16 && N <= 256 && K % 32 == 0 && K >= 32 && K <= 256,
                "ts_selftest: bad argument");
  V1T_CUDA(cudaFuncSetAttribute(ts_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  ts_selftest_kernel<<<1, 128, 160 * 1024, (cudaStream_t)stream>>>(A, B, C, N, K);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

// cycles for `iters * 8` MMAs of shape 128 x N x 16 on every SM concurrently (grid = 148); returns the max over SMs
extern "C" int v1t_mma_microbench(int N, int ts, int iters, int mn_b, long long* out_dev, void* stream) {
  using namespace v1t;
  V1T_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256 && out_dev, "mma_microbench: bad argument");
  V1T_CUDA(cudaFuncSetAttribute(mma_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  mma_microbench_kernel<<<kNumSMs, 128, 100 * 1024, (cudaStream_t)stream>>>(N, ts, iters, mn_b, out_dev);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

// ---------------------------------------------------------------------------------------------------------
// cp.async.bulk fill-rate microbenchmark: every SM streams `iters` rounds of `copies` copies of `bytes` bytes from a
// global buffer (`span` bytes, wrapped) into a ring of `slots` shared-memory slots; a slot is re-filled as soon as
// its previous contents landed (nothing consumes the data).  out_dev[sm] = cycles.  Answers "how many GB/s can one
// SM pull in, as a function of the copy size and the bytes in flight" -- the bound of the bf16x3 attention / GEMM
// main loops.
namespace v1t {
namespace {
__global__ void __launch_bounds__(32, 1) bulk_microbench_kernel(const uint8_t* __restrict__ src, int64_t span, int bytes,
                                                                int copies, int slots, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8];
  using namespace tc;
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncwarp();
  const int64_t slot_bytes = (int64_t)bytes * copies;
  const int64_t per_cta = slot_bytes * iters;
  long long t0 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    for (int it = 0; it < iters + slots; ++it) {
      const int s = it % slots;
      if (it >= slots) mbar_wait(&full[s], ((it / slots) - 1) & 1);  // previous fill of this slot has landed
      if (it < iters) {
        mbar_expect_tx(&full[s], (uint32_t)slot_bytes);
        int64_t off = ((int64_t)blockIdx.x * per_cta + (int64_t)it * slot_bytes) % (span - slot_bytes);
        off &= ~(int64_t)127;
        for (int cidx = 0; cidx < copies; ++cidx) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(smem + (int64_t)s * slot_bytes + (int64_t)cidx * bytes)),
                       "l"(src + off + (int64_t)cidx * bytes), "r"(bytes), "r"(smem_u32(&full[s]))
                       : "memory");
        }
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
}  // namespace
}  // namespace v1t

extern "C" int v1t_bulk_microbench(const void* src, long long span, int bytes, int copies, int slots, int iters,
                                   long long* out_dev, void* stream) {
  using namespace v1t;
  V1T_CHECK_ARG(src && out_dev && bytes > 0 && bytes % 16 == 0 && copies >= 1 && slots >= 1 && slots <= 8 && iters >= 1 &&
                    (long long)bytes * copies * slots <= 200 * 1024 && span >= 2ll * bytes * copies,
                "bulk_microbench: bad argument");
  V1T_CUDA(cudaFuncSetAttribute(bulk_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024));
  bulk_microbench_kernel<<<kNumSMs, 32, 201 * 1024 + 1024, (cudaStream_t)stream>>>((const uint8_t*)src, span, bytes, copies,
                                                                                    slots, iters, out_dev);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}
