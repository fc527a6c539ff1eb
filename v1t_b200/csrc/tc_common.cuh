// sm_100a primitives used by the tensor-core kernels: mbarrier, proxy fences, TMEM allocation,
// tcgen05.mma / commit / ld, UMMA shared-memory and instruction descriptors.  Raw inline PTX (no CUTLASS).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace v1t {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait is a potentially BLOCKING test: the hardware may suspend the thread until the phase completes or a
// system-dependent time limit passes.  A suspendTimeHint operand (compile with -DV1T_MBAR_HINT_NS=<ns>) stretches that
// limit -- SASS: `@!P0 NANOSLEEP.SYNCS <ns>` after the first failed check.  Measured with 20 us on the bench step: the
// polling instructions (SYNCS / ISETP / BRA / CS2R / IADD3: ~40 % of the executed warp instructions of the attention
// backward in the ncu source view) disappear, but every kernel got 3-5 % SLOWER (attention forward 320 -> 330 us,
// backward 1251 -> 1317 us): the wake-up adds latency on the MMA-issue and softmax critical paths.  Default: no hint.
#ifndef V1T_MBAR_HINT_NS
#define V1T_MBAR_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#if V1T_MBAR_HINT_NS > 0
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)V1T_MBAR_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error surfaced to Python) instead of hanging the GPU.  The clock is read
// once per 256 failed polls only.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++polls & 255u) == 0u && clock64() - t0 > 4000000000ll) __trap();
  }
}
// explicit shared-space 128-bit accesses (pointers derived through integer casts otherwise compile to generic LD/ST)
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (UMMA operand fetch, TMA store)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; single thread issues
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 16 consecutive fp32 columns: thread `lane` gets TMEM lane (lane_base + lane), columns col..col+15
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// stores: thread `lane` writes TMEM lane (lane_base + lane), 8 / 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
// 8 fp32 -> 4 packed bf16x2 words (hi) and 4 words (lo = x - hi)
__device__ __forceinline__ void split8_words(const float (&x)[8], uint32_t (&h)[4], uint32_t (&l)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hb);
    const __nv_bfloat162 lb = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hb);
    l[i] = *reinterpret_cast<const uint32_t*>(&lb);
  }
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------------------
// K-major operand tile stored as [rows][64 B] (32 bf16 of K per row), 64-byte swizzle (Swizzle<2,4,3>):
// 16-byte chunk c of row r lives at  r*64 + ((c ^ ((r >> 1) & 3)) << 4).  Tile base must be 512 B aligned.
// SBO = 8 rows * 64 B = 512 B; LBO unused for swizzled K-major; version = 1 (sm_100); layout type 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t desc_k_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                    // leading byte offset (ignored)
  d |= (uint64_t)(512 >> 4) << 32;           // stride byte offset
  d |= (uint64_t)1 << 46;                    // descriptor version
  d |= (uint64_t)4 << 61;                    // SWIZZLE_64B
  return d;
}
// MN-major operand stored as [atoms along MN][rows along K][64 B] (32 bf16 of MN per row), 64-byte swizzle:
// LBO = byte stride between 32-element atoms along MN, SBO = stride between groups of 8 K-rows (= 512 B).
__device__ __forceinline__ uint64_t desc_mn_sw64(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// constant (address-independent) parts: descriptor(addr) = base | (addr >> 4); advancing the start address by
// `bytes` is a plain add of (bytes >> 4) to the low word.
constexpr uint64_t kDescK64 = ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
__device__ __forceinline__ uint64_t desc_mn_sw64_base(uint32_t lbo_bytes) {
  return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint32_t sw64_offset(int row, int chunk) {
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M x N tile, operand major-ness (0 = K, 1 = MN)
__host__ __device__ inline uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                       // C format F32
  d |= 1u << 7;                       // A format BF16
  d |= 1u << 10;                      // B format BF16
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// fp32 -> bf16 hi + bf16 lo (x ~= hi + lo, ~16 mantissa bits); 8 values -> two 16-byte chunks
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hb);
    const __nv_bfloat162 lb = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hb);
    l[i] = *reinterpret_cast<const uint32_t*>(&lb);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Attention planes (Q, K, V, dO per (sample, head)): [b*H + h][Tq/64 key/query tiles][AD atoms][64 rows][64 B], i.e. the
// AD head-dim atoms of a 64-token tile are CONTIGUOUS (AD * 4 KB): a streamed K / V / Q / dO tile is one bulk copy
// per plane that lands in shared memory exactly as [atom][64 rows][64 B].  (Bulk copies cost ~90 cycles each on
// top of the bytes -- scripts/bulk_microbench.py: 4 KB copies cap an SM at 68 GB/s, 20 KB copies at 105-114 -- and
// the forward needs 67 GB/s.)  Returns the byte offset of row t of atom `atom`; Tq is a multiple of 128.
__host__ __device__ __forceinline__ int64_t attn_plane_off(int64_t bh, int atom, int t, int Tq, int AD) {
  return ((((bh * (Tq >> 6) + (t >> 6)) * AD + atom) << 6) + (t & 63)) << 6;
}
// Row `t` of a 128-row RESIDENT operand (attention plane of head `bh`) -> tensor memory columns [tcol, tcol + AD*16) of
// this thread's lane, for the head-dim atoms at0, at0 + astep, ...  All global loads are issued before the first
// tcgen05.st: the stores are asm volatile, so a load-store-load-store sequence serialised one global round trip per
// atom (the prologue of the attention kernels measured ~12 k cycles per CTA, scripts/pair_trace.py).
template <int AD>
__device__ __forceinline__ void plane_row_to_tmem(const uint8_t* plane, int64_t bh, int t, int Tp, uint32_t taddr,
                                                  int at0, int astep) {
  uint4 ph[AD][4];
#pragma unroll
  for (int at_i = 0; at_i < AD; ++at_i) {
    if (at_i >= at0 && (at_i - at0) % astep == 0) {
      const uint4* src = reinterpret_cast<const uint4*>(plane + attn_plane_off(bh, at_i, t, Tp, AD));
#pragma unroll
      for (int p = 0; p < 4; ++p) ph[at_i][p] = __ldg(src + p);
    }
  }
  // logical chunk c sits at physical position c ^ sw.  tcgen05.st takes ONE address per warp (lane i writes TMEM lane
  // base + i at the same columns), so the per-row swizzle must permute the DATA, not the column: an XOR network of
  // conditional swaps (no runtime register index, i.e. no local-memory array)
  const int sw = (t >> 1) & 3;
  auto cswap = [](bool on, uint4& x, uint4& y) {
    const uint4 a = x, b = y;
    x = on ? b : a;
    y = on ? a : b;
  };
#pragma unroll
  for (int at_i = 0; at_i < AD; ++at_i) {
    if (at_i >= at0 && (at_i - at0) % astep == 0) {
      cswap(sw & 1, ph[at_i][0], ph[at_i][1]);
      cswap(sw & 1, ph[at_i][2], ph[at_i][3]);
      cswap(sw & 2, ph[at_i][0], ph[at_i][2]);
      cswap(sw & 2, ph[at_i][1], ph[at_i][3]);
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st4(taddr + at_i * 16 + c * 4, ph[at_i][c].x, ph[at_i][c].y, ph[at_i][c].z, ph[at_i][c].w);
    }
  }
}
// The same from a copy of the 128-row tile that a bulk copy staged in shared memory ([2 tiles of 64 rows][AD atoms][64 rows]
// [64 B], exactly as it lies in the plane): the per-row swizzle is resolved by the (per-lane) shared-memory ADDRESS, so the
// tensor-memory address stays warp-uniform.  One TMA fetch of 40 KB per plane replaces 128 x AD scattered 64-byte row
// pieces per plane (the register-path prologue measured 7.9 k cycles per CTA).
template <int AD>
__device__ __forceinline__ void smem_row_to_tmem(uint32_t smem_plane, int row, uint32_t taddr, int at0, int astep) {
  const uint32_t base = smem_plane + (uint32_t)(row >> 6) * (AD * 4096u) + (uint32_t)(row & 63) * 64u;
  const uint32_t sw = (uint32_t)(row >> 1) & 3u;
  // one atom at a time (4 chunks in flight): shared-memory latency is short, and 16 live registers instead of 16 x atoms
  // keep this out of local memory inside the persistent kernels
#pragma unroll
  for (int at_i = 0; at_i < AD; ++at_i)
    if (at_i >= at0 && (at_i - at0) % astep == 0) {
      uint4 v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = lds128(base + at_i * 4096u + ((c ^ sw) << 4));
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st4(taddr + at_i * 16 + c * 4, v[c].x, v[c].y, v[c].z, v[c].w);
    }
}
// byte offset of 16-byte chunk `chunk` of row `row` in column atom `atom` of a matrix plane ([atoms][rows_p][64 B])
__device__ __forceinline__ int64_t plane_chunk_off(int64_t atom, int64_t rows_p, int64_t row, int chunk) {
  return (atom * rows_p + row) * 64 + ((chunk ^ (int)((row >> 1) & 3)) << 4);
}
// 4 adjacent values -> 8 bytes of the hi plane and 8 bytes of the lo plane
__device__ __forceinline__ void split4(const float (&x)[4], uint2& hi, uint2& lo) {
  uint32_t h[2], l[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const __nv_bfloat162 hb = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hb);
    const __nv_bfloat162 lb = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hb);
    l[i] = *reinterpret_cast<const uint32_t*>(&lb);
  }
  hi = make_uint2(h[0], h[1]);
  lo = make_uint2(l[0], l[1]);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

}  // namespace tc
}  // namespace v1t
