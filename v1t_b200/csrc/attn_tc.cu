// Fused attention forward / backward on tcgen05 (reference: Attention.scaled_dot_product_attention,
// vit.py:253-265, and its autograd).
//
//   P = softmax(Q K^T * E^-0.5) (dropout)   O = P V        per (sample, head), T = 1654 tokens, head dim E = 155
//
// FORWARD: one CTA per (b, h, 128-query tile); K / V^T stream through shared memory in 64-key tiles fetched with
// cp.async.bulk from the pre-swizzled bf16 planes (planes.cu); S and O accumulate in TMEM; nothing of size T x T
// reaches HBM.  bf16x3 mode multiplies hi/lo split operands (Q, K, P, V all split) -> fp32-class accuracy.
// Softmax is TWO-PASS instead of online rescaling of the O accumulator in TMEM:
//   pass 1: S = Qhi Khi^T only (1 MMA product) -> per-row reference maximum m (any value near the true max works)
//   pass 2: S (full precision), P = exp2(S*c - m*c), l += rowsum(P), O += P V   (no correction step / TMEM RMW)
// Warp roles (320 threads): warps 0-7 softmax/epilogue (TMEM lane quarter = warp & 3, column half = warp >> 2, so
// every scheduler has two softmax warps to hide latencies), warp 8 TMEM alloc + MMA issue (descriptors are
// base + constant, loops fully unrolled on the head-dim atom count AD), warp 9 bulk-copy producer.
// S is double-buffered in TMEM so QK^T of tile j+1 overlaps softmax of tile j.
//
// BACKWARD: two launches of ONE kernel template, both atomic-free and deterministic:
//   KV = true : CTA per (b, h, 128-key tile), streams query tiles      -> dV = Pd^T dO,  dK = scale * dS^T Q
//   KV = false: CTA per (b, h, 128-query tile), streams key tiles      -> dQ = scale * dS K
// with P = exp2(S*c - lse) recomputed from the saved log-sum-exp, Pd = P * dropout, dP = (dO V^T) * dropout,
// dS = P * (dP - delta), delta = rowsum(dO * O).  Per streamed tile j (N tokens) the tensor core computes
//   S' = X x_j^T,  dP' = Y y_j^T          (X, Y: resident 128-row operands; x_j, y_j: streamed, K-major over d)
//   out1 += Pd' y_j (KV only),  out2 += dS' x_j     (x_j / y_j re-used un-transposed through MN-major descriptors)
// S'/dP' double-buffered in TMEM, streamed tiles double-buffered in shared memory; Pd' and dS' return to shared
// memory as bf16 hi/lo A-operands (they share one 64-byte-row atom when N = 16).  bf16x3: N = 16 (2 x 80 KB of
// resident hi+lo operands leave room for nothing larger), bf16: N = 32.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace v1t {
namespace {

using namespace tc;

constexpr int BQ = 128, BKEY = 64;
constexpr int kSmWarps = 8;                       // softmax warps
constexpr int kSmThreads = kSmWarps * 32;
constexpr int kThreadsAttn = (kSmWarps + 2) * 32;  // + MMA warp + loader warp
constexpr int kMmaWarp = kSmWarps, kLoadWarp = kSmWarps + 1;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
struct FwdSmem {
  uint32_t q_hi, q_lo, k_hi, k_lo, v_hi, v_lo, p_hi, p_lo, bars, xch, total;
};
__host__ __device__ inline FwdSmem fwd_smem_layout(int Dp) {
  FwdSmem s;
  const uint32_t ad = Dp / 32;
  uint32_t o = 0;
  s.q_hi = o; o += ad * BQ * 64;
  s.q_lo = o; o += ad * BQ * 64;
  s.k_hi = o; o += ad * BKEY * 64;   // k_hi, k_lo, v_hi, v_lo are contiguous and equally sized:
  s.k_lo = o; o += ad * BKEY * 64;   // pass 1 uses them as a 4-slot ring of hi-only K tiles
  s.v_hi = o; o += 2 * Dp * 64;
  s.v_lo = o; o += 2 * Dp * 64;
  s.p_hi = o; o += 2 * BQ * 64;
  s.p_lo = o; o += 2 * BQ * 64;
  s.bars = o; o += 256;
  s.xch = o; o += 2 * BQ * 4;        // cross-half exchange of row max / row sum
  s.total = o + 1024;                // + alignment slack
  return s;
}

template <int AD>
__global__ void __launch_bounds__(kThreadsAttn, 1) attn_fwd_kernel(const AttnFwdArgs a) {
  constexpr int Dp = AD * 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const FwdSmem L = fwd_smem_layout(Dp);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* p_full = bars + 5;
  uint64_t* p_empty = bars + 6;
  uint64_t* o_full = bars + 7;
  uint64_t* s_full = bars + 8;    // [2]
  uint64_t* s_empty = bars + 10;  // [2]
  uint64_t* r_full = bars + 12;   // [4] pass-1 K ring
  uint64_t* r_empty = bars + 16;  // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* xch = reinterpret_cast<float*>(smem + L.xch);
  constexpr uint32_t kRingSlot = AD * BKEY * 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.y;
  const int q0 = qt * BQ;
  const int at = a.Tp / 32;
  const int nk = (a.T + BKEY - 1) / BKEY;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(p_full, kSmThreads);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], kSmThreads);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * BKEY;

  if (warp == kLoadWarp) {
    // ============================== BULK-COPY PRODUCER ==============================
    if (lane == 0) {
      constexpr uint32_t q_bytes = AD * BQ * 64;
      mbar_expect_tx(q_full, a.x3 ? 2 * q_bytes : q_bytes);
#pragma unroll
      for (int at_i = 0; at_i < AD; ++at_i) {
        const int64_t src = (((int64_t)bh * AD + at_i) * a.Tp + q0) * 64;
        bulk_g2s(smem + L.q_hi + at_i * BQ * 64, a.q_hi + src, BQ * 64, q_full);
        if (a.x3) bulk_g2s(smem + L.q_lo + at_i * BQ * 64, a.q_lo + src, BQ * 64, q_full);
      }
      constexpr uint32_t k_bytes = AD * BKEY * 64, v_bytes = 2 * Dp * 64;
      auto load_k = [&](uint32_t dst_hi, uint32_t dst_lo, int j, bool lo, uint64_t* bar) {
        mbar_expect_tx(bar, lo ? 2 * k_bytes : k_bytes);
#pragma unroll
        for (int at_i = 0; at_i < AD; ++at_i) {
          const int64_t src = (((int64_t)bh * AD + at_i) * a.Tp + j * BKEY) * 64;
          bulk_g2s(smem + dst_hi + at_i * BKEY * 64, a.k_hi + src, BKEY * 64, bar);
          if (lo) bulk_g2s(smem + dst_lo + at_i * BKEY * 64, a.k_lo + src, BKEY * 64, bar);
        }
      };
      // pass 1: hi planes of K only, 4-slot ring over the (still unused) K/V regions
      for (int j = 0; j < nk; ++j) {
        const int slot = j & 3;
        mbar_wait(&r_empty[slot], ((j >> 2) & 1) ^ 1);
        load_k(L.k_hi + slot * kRingSlot, 0, j, false, &r_full[slot]);
      }
      for (int slot = 0; slot < 4 && slot < nk; ++slot) {  // all pass-1 MMAs have released their slots
        const int last = ((nk - 1 - slot) / 4) * 4 + slot;
        mbar_wait(&r_empty[slot], (last >> 2) & 1);
      }
      // pass 2: K runs one tile ahead of V (K(j+1) is free after S(j), V(j) after P V(j-1))
      load_k(L.k_hi, L.k_lo, 0, a.x3 != 0, k_full);
      for (int j = 0; j < nk; ++j) {
        if (j + 1 < nk) {
          mbar_wait(k_empty, j & 1);
          load_k(L.k_hi, L.k_lo, j + 1, a.x3 != 0, k_full);
        }
        mbar_wait(v_empty, (j & 1) ^ 1);
        mbar_expect_tx(v_full, a.x3 ? 2 * v_bytes : v_bytes);
#pragma unroll
        for (int ka = 0; ka < 2; ++ka) {
          const int64_t src = (((int64_t)bh * at + (j * 2 + ka)) * Dp) * 64;
          bulk_g2s(smem + L.v_hi + ka * Dp * 64, a.vt_hi + src, Dp * 64, v_full);
          if (a.x3) bulk_g2s(smem + L.v_lo + ka * Dp * 64, a.vt_lo + src, Dp * 64, v_full);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA ISSUER ==============================
    const uint32_t idesc_s = idesc_bf16(BQ, BKEY, 0, 0);
    const uint32_t idesc_o = idesc_bf16(BQ, Dp, 0, 0);
    const uint64_t dq_hi = kDescK64 | (smem_u32(smem + L.q_hi) >> 4), dq_lo = kDescK64 | (smem_u32(smem + L.q_lo) >> 4);
    const uint64_t dk_hi = kDescK64 | (smem_u32(smem + L.k_hi) >> 4), dk_lo = kDescK64 | (smem_u32(smem + L.k_lo) >> 4);
    const uint64_t dv_hi = kDescK64 | (smem_u32(smem + L.v_hi) >> 4), dv_lo = kDescK64 | (smem_u32(smem + L.v_lo) >> 4);
    const uint64_t dp_hi = kDescK64 | (smem_u32(smem + L.p_hi) >> 4), dp_lo = kDescK64 | (smem_u32(smem + L.p_lo) >> 4);
    uint32_t itk = 0, its = 0;

    const bool leader = elect_one();  // the same lane issues every tcgen05.mma / commit of this CTA

    // S[buf] = Q K^T from the K tile described by (kh, kl); commits `k_done` and s_full[buf]
    auto issue_s = [&](uint64_t kh, uint64_t kl, bool full_precision, uint64_t* k_ready, uint32_t k_parity,
                       uint64_t* k_done) {
      const uint32_t buf = its & 1;
      mbar_wait(k_ready, k_parity);
      mbar_wait(&s_empty[buf], ((its >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + buf * BKEY;
#pragma unroll
      for (int ks = 0; ks < 2 * AD; ++ks) {
        constexpr uint32_t kQA = BQ * 64 / 16, kKA = BKEY * 64 / 16;  // atom strides in 16-byte units
        const uint32_t qo = (ks >> 1) * kQA + (ks & 1) * 2, ko = (ks >> 1) * kKA + (ks & 1) * 2;
        if (leader) {
          umma_bf16(d, dq_hi + qo, kh + ko, idesc_s, ks > 0 ? 1u : 0u);
          if (full_precision) {
            umma_bf16(d, dq_lo + qo, kh + ko, idesc_s, 1u);
            umma_bf16(d, dq_hi + qo, kl + ko, idesc_s, 1u);
          }
        }
      }
      if (leader) {
        umma_commit(k_done);
        umma_commit(&s_full[buf]);
      }
      __syncwarp();
      ++its;
    };
    auto issue_s2 = [&]() {  // pass-2 tile from the K buffer
      issue_s(dk_hi, dk_lo, a.x3 != 0, k_full, itk & 1, k_empty);
      ++itk;
    };
    auto issue_pv = [&](int j, bool last) {
      mbar_wait(p_full, j & 1);
      mbar_wait(v_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < BKEY / 16; ++ks) {
        constexpr uint32_t kPA = BQ * 64 / 16, kVA = Dp * 64 / 16;
        const uint32_t po = (ks >> 1) * kPA + (ks & 1) * 2, vo = (ks >> 1) * kVA + (ks & 1) * 2;
        if (leader) {
          umma_bf16(tmem_o, dp_hi + po, dv_hi + vo, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          if (a.x3) {
            umma_bf16(tmem_o, dp_lo + po, dv_hi + vo, idesc_o, 1u);
            umma_bf16(tmem_o, dp_hi + po, dv_lo + vo, idesc_o, 1u);
          }
        }
      }
      if (leader) {
        umma_commit(p_empty);
        umma_commit(v_empty);
        if (last) umma_commit(o_full);
      }
      __syncwarp();
    };

    mbar_wait(q_full, 0);
    for (int j = 0; j < nk; ++j) {  // pass 1: reference row max from the hi planes
      const int slot = j & 3;
      issue_s(dk_hi + slot * (kRingSlot >> 4), 0, false, &r_full[slot], (j >> 2) & 1, &r_empty[slot]);
    }
    issue_s2();  // pass 2, tile 0
    for (int j = 0; j + 1 < nk; ++j) {
      issue_s2();  // S(j+1) overlaps softmax(j)
      issue_pv(j, false);
    }
    issue_pv(nk - 1, true);
  } else {
    // ============================== SOFTMAX / EPILOGUE ==============================
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;       // TMEM lane = query row of the tile
    const int qi = q0 + row;                   // token index
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int b = bh / a.H, h = bh % a.H;
    uint32_t its = 0;
    float m = -INFINITY;
    // ---- pass 1: row max over this warp's 32 of the 64 key columns
    for (int j = 0; j < nk; ++j, ++its) {
      const uint32_t buf = its & 1;
      mbar_wait(&s_full[buf], (its >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_off + buf * BKEY + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[buf]);
      const int jb = j * BKEY + half * 32;
      if (jb + 32 <= a.T) {
#pragma unroll
        for (int c = 0; c < 32; ++c) m = fmaxf(m, __uint_as_float(v[c]));
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (jb + c < a.T) m = fmaxf(m, __uint_as_float(v[c]));
      }
    }
    xch[half * BQ + row] = m;
    named_bar_sync(1, kSmThreads);
    m = fmaxf(xch[row], xch[BQ + row]);
    named_bar_sync(1, kSmThreads);
    const float m2 = m * a.scale_log2;
    float l = 0.f;
    const float inv_keep = a.drop.p > 0.f ? 1.f / (1.f - a.drop.p) : 1.f;
    const int Tc = (a.T + 3) & ~3;
    const uint64_t drop_row = ((uint64_t)bh * a.T + (uint64_t)min(qi, a.T - 1)) * (uint64_t)Tc;
    // ---- pass 2: P = exp2(S*c - m*c), l += rowsum(P), P (dropout) -> smem as the A operand of P V
    for (int j = 0; j < nk; ++j, ++its) {
      const uint32_t buf = its & 1;
      mbar_wait(&s_full[buf], (its >> 1) & 1);
      tc_fence_after();
      float p[32];
      {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + buf * BKEY + half * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) p[c] = __uint_as_float(v[c]);
      }
      tc_fence_before();
      mbar_arrive(&s_empty[buf]);
      const int jb = j * BKEY + half * 32;
      if (jb + 32 <= a.T) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          p[c] = fast_exp2(fmaf(p[c], a.scale_log2, -m2));
          l += p[c];
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          p[c] = (jb + c < a.T) ? fast_exp2(fmaf(p[c], a.scale_log2, -m2)) : 0.f;
          l += p[c];
        }
      }
      if (a.drop.p > 0.f) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float mk[4];
          dropout_mult4(a.drop.seed, a.drop.site, (drop_row + (uint64_t)(jb + 4 * g)) >> 2, a.drop.p, inv_keep, mk);
#pragma unroll
          for (int e = 0; e < 4; ++e) p[4 * g + e] *= mk[e];
        }
      }
      mbar_wait(p_empty, (j & 1) ^ 1);  // P V of the previous tile has consumed the buffer
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = p[ch * 8 + e];
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = half * (BQ * 64) + sw64_offset(row, ch);
        *reinterpret_cast<uint4*>(smem + L.p_hi + off) = hi;
        if (a.x3) *reinterpret_cast<uint4*>(smem + L.p_lo + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(p_full);
    }
    xch[half * BQ + row] = l;
    named_bar_sync(1, kSmThreads);
    l = xch[row] + xch[BQ + row];
    // ---- epilogue: O / l -> global (each half writes AD*16 of the Dp columns), base-2 log-sum-exp
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.f / l;
    float* orow = a.O + ((int64_t)b * a.T + qi) * a.o_ld + h * a.E;
#pragma unroll
    for (int cc = 0; cc < AD; ++cc) {
      const int c0 = half * (AD * 16) + cc * 16;
      uint32_t v[16];
      tmem_ld16(tmem_o + lane_off + c0, v);
      tmem_ld_wait();
      if (qi < a.T) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c0 + c < a.E) orow[c0 + c] = __uint_as_float(v[c]) * inv_l;
      }
    }
    // padded query rows get +inf so that the backward's exp2(S*c - lse) vanishes there without bounds checks
    if (half == 0 && a.lse) a.lse[(int64_t)bh * a.Tp + qi] = qi < a.T ? m2 + log2f(l) : INFINITY;
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

template <int AD>
int launch_fwd(const AttnFwdArgs& a, cudaStream_t st) {
  const FwdSmem L = fwd_smem_layout(AD * 32);
  V1T_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<AD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  dim3 grid(cdiv(a.T, BQ), a.B * a.H);
  attn_fwd_kernel<AD><<<grid, kThreadsAttn, L.total, st>>>(a);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

// ---------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------
struct BwdSmem {
  uint32_t x_hi, x_lo, y_hi, y_lo, st[2][4] /* xj_hi, xj_lo, yj_hi, yj_lo */, ps_hi, ps_lo, bars, total;
};
__host__ __device__ inline BwdSmem bwd_smem_layout(int Dp, int N, int x3) {
  BwdSmem s;
  const uint32_t ad = Dp / 32;
  uint32_t o = 0;
  s.x_hi = o; o += ad * 128 * 64;
  s.x_lo = o; if (x3) o += ad * 128 * 64;
  s.y_hi = o; o += ad * 128 * 64;
  s.y_lo = o; if (x3) o += ad * 128 * 64;
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < 4; ++k) {
      s.st[i][k] = o;
      if (x3 || (k & 1) == 0) o += ad * N * 64;
    }
  const uint32_t ps_atoms = (2 * N + 31) / 32;
  s.ps_hi = o; o += ps_atoms * 128 * 64;
  s.ps_lo = o; if (x3) o += ps_atoms * 128 * 64;
  s.bars = o; o += 256;
  s.total = o + 1024;
  return s;
}

template <bool KV, int N, int AD>
__global__ void __launch_bounds__(kThreadsAttn, 1) attn_bwd_kernel(const AttnBwdArgs a) {
  constexpr int Dp = AD * 32;
  constexpr int NH = N / 2;  // columns per softmax thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const BwdSmem L = bwd_smem_layout(Dp, N, a.x3);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* res_full = bars + 0;
  uint64_t* ps_full = bars + 1;
  uint64_t* ps_empty = bars + 2;
  uint64_t* o_full = bars + 3;
  uint64_t* st_full = bars + 4;    // [2]
  uint64_t* st_empty = bars + 6;   // [2]
  uint64_t* sp_full = bars + 8;    // [2]
  uint64_t* sp_empty = bars + 10;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 128, bh = blockIdx.y;  // first resident row (key for KV, query otherwise)
  const int nt = (a.T + N - 1) / N;
  const uint8_t* X_hi = KV ? a.k_hi : a.q_hi;   const uint8_t* X_lo = KV ? a.k_lo : a.q_lo;
  const uint8_t* Y_hi = KV ? a.v_hi : a.do_hi;  const uint8_t* Y_lo = KV ? a.v_lo : a.do_lo;
  const uint8_t* xs_hi = KV ? a.q_hi : a.k_hi;  const uint8_t* xs_lo = KV ? a.q_lo : a.k_lo;
  const uint8_t* ys_hi = KV ? a.do_hi : a.v_hi; const uint8_t* ys_lo = KV ? a.do_lo : a.v_lo;

  if (threadIdx.x == 0) {
    mbar_init(res_full, 1);
    mbar_init(ps_full, kSmThreads);
    mbar_init(ps_empty, 1);
    mbar_init(o_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&st_full[i], 1);
      mbar_init(&st_empty[i], 1);
      mbar_init(&sp_full[i], 1);
      mbar_init(&sp_empty[i], kSmThreads);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S'[2][N] | dP'[2][N] | out1[Dp] | out2[Dp]
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 2 * N, tm_o1 = tmem_base + 4 * N, tm_o2 = tm_o1 + Dp;
  constexpr uint32_t tb = N * 64;  // bytes of one head-dim atom of a streamed tile

  if (warp == kLoadWarp) {
    // ============================== BULK-COPY PRODUCER ==============================
    if (lane == 0) {
      constexpr uint32_t rb = AD * 128 * 64;
      mbar_expect_tx(res_full, (a.x3 ? 4 : 2) * rb);
#pragma unroll
      for (int at_i = 0; at_i < AD; ++at_i) {
        const int64_t src = (((int64_t)bh * AD + at_i) * a.Tp + r0) * 64;
        bulk_g2s(smem + L.x_hi + at_i * 8192, X_hi + src, 8192, res_full);
        bulk_g2s(smem + L.y_hi + at_i * 8192, Y_hi + src, 8192, res_full);
        if (a.x3) {
          bulk_g2s(smem + L.x_lo + at_i * 8192, X_lo + src, 8192, res_full);
          bulk_g2s(smem + L.y_lo + at_i * 8192, Y_lo + src, 8192, res_full);
        }
      }
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        mbar_wait(&st_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&st_full[s], (a.x3 ? 4 : 2) * AD * tb);
#pragma unroll
        for (int at_i = 0; at_i < AD; ++at_i) {
          const int64_t src = (((int64_t)bh * AD + at_i) * a.Tp + (int64_t)j * N) * 64;
          bulk_g2s(smem + L.st[s][0] + at_i * tb, xs_hi + src, tb, &st_full[s]);
          bulk_g2s(smem + L.st[s][2] + at_i * tb, ys_hi + src, tb, &st_full[s]);
          if (a.x3) {
            bulk_g2s(smem + L.st[s][1] + at_i * tb, xs_lo + src, tb, &st_full[s]);
            bulk_g2s(smem + L.st[s][3] + at_i * tb, ys_lo + src, tb, &st_full[s]);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA ISSUER ==============================
    const uint32_t idesc_s = idesc_bf16(128, N, 0, 0);
    const uint32_t idesc_o = idesc_bf16(128, Dp, 0, 1);  // B = streamed tile viewed MN-major (d contiguous)
    const uint64_t dX_hi = kDescK64 | (smem_u32(smem + L.x_hi) >> 4), dX_lo = kDescK64 | (smem_u32(smem + L.x_lo) >> 4);
    const uint64_t dY_hi = kDescK64 | (smem_u32(smem + L.y_hi) >> 4), dY_lo = kDescK64 | (smem_u32(smem + L.y_lo) >> 4);
    const uint64_t dPS_hi = kDescK64 | (smem_u32(smem + L.ps_hi) >> 4), dPS_lo = kDescK64 | (smem_u32(smem + L.ps_lo) >> 4);
    const uint64_t mn_base = desc_mn_sw64_base(tb);
    // K-major / MN-major descriptors of the stage-0 streamed buffers; stage 1 = + st_stride (16-byte units)
    uint64_t dk[4], dm[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t addr = smem_u32(smem + L.st[0][k]) >> 4;
      dk[k] = kDescK64 | addr;
      dm[k] = mn_base | addr;
    }
    const uint32_t st_stride = (L.st[1][0] - L.st[0][0]) >> 4;

    const bool leader = elect_one();  // the same lane issues every tcgen05.mma / commit of this CTA

    auto issue_scores = [&](int j) {  // S'(j) = X x_j^T, dP'(j) = Y y_j^T
      const int s = j & 1;
      mbar_wait(&st_full[s], (j >> 1) & 1);
      mbar_wait(&sp_empty[s], ((j >> 1) & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint32_t d = (which ? tm_dp : tm_s) + s * N;
        const uint64_t Ah = which ? dY_hi : dX_hi, Al = which ? dY_lo : dX_lo;
        const uint64_t Bh = dk[which * 2] + s * st_stride, Bl = dk[which * 2 + 1] + s * st_stride;
#pragma unroll
        for (int ks = 0; ks < 2 * AD; ++ks) {
          const uint32_t ao = (ks >> 1) * (8192 / 16) + (ks & 1) * 2, bo = (ks >> 1) * (tb / 16) + (ks & 1) * 2;
          if (leader) {
            umma_bf16(d, Ah + ao, Bh + bo, idesc_s, ks > 0 ? 1u : 0u);
            if (a.x3) {
              umma_bf16(d, Al + ao, Bh + bo, idesc_s, 1u);
              umma_bf16(d, Ah + ao, Bl + bo, idesc_s, 1u);
            }
          }
        }
      }
      if (leader) umma_commit(&sp_full[s]);
      __syncwarp();
    };
    auto issue_out = [&](int j, bool last) {  // out1 += Pd' y_j (KV), out2 += dS' x_j
      const int s = j & 1;
      mbar_wait(ps_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int which = KV ? 0 : 1; which < 2; ++which) {
        const uint32_t d = which ? tm_o2 : tm_o1;
        const uint64_t Bh = (which ? dm[0] : dm[2]) + s * st_stride, Bl = (which ? dm[1] : dm[3]) + s * st_stride;
#pragma unroll
        for (int ks = 0; ks < N / 16; ++ks) {
          const int kel = (which ? N : 0) + ks * 16;  // element offset along the packed [Pd' | dS'] K axis
          const uint32_t ao = (kel >> 5) * (8192 / 16) + ((kel >> 4) & 1) * 2;
          if (leader) {
            umma_bf16(d, dPS_hi + ao, Bh + ks * 64, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
            if (a.x3) {
              umma_bf16(d, dPS_lo + ao, Bh + ks * 64, idesc_o, 1u);
              umma_bf16(d, dPS_hi + ao, Bl + ks * 64, idesc_o, 1u);
            }
          }
        }
      }
      if (leader) {
        umma_commit(ps_empty);
        umma_commit(&st_empty[s]);
        if (last) umma_commit(o_full);
      }
      __syncwarp();
    };

    mbar_wait(res_full, 0);
    issue_scores(0);
    for (int j = 0; j + 1 < nt; ++j) {
      issue_scores(j + 1);  // overlaps the softmax-backward of tile j
      issue_out(j, false);
    }
    issue_out(nt - 1, true);
  } else {
    // ============================== SOFTMAX-BACKWARD / EPILOGUE ==============================
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const int ri = r0 + row;  // key index (KV) or query index
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int b = bh / a.H, h = bh % a.H;
    const float inv_keep = a.drop.p > 0.f ? 1.f / (1.f - a.drop.p) : 1.f;
    const int Tc = (a.T + 3) & ~3;
    const float* lse = a.lse + (int64_t)bh * a.Tp;
    const float* delta = a.delta + (int64_t)bh * a.Tp;
    float lse_r = 0.f, delta_r = 0.f;
    if (!KV && ri < a.T) { lse_r = lse[ri]; delta_r = delta[ri]; }
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      mbar_wait(&sp_full[s], (j >> 1) & 1);
      tc_fence_after();
      float sv[NH], dv[NH];
      {
        uint32_t v1[NH], v2[NH];
        if constexpr (NH == 8) {
          tmem_ld8(tm_s + lane_off + s * N + half * NH, v1);
          tmem_ld8(tm_dp + lane_off + s * N + half * NH, v2);
        } else {
          tmem_ld16(tm_s + lane_off + s * N + half * NH, v1);
          tmem_ld16(tm_dp + lane_off + s * N + half * NH, v2);
        }
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < NH; ++c) {
          sv[c] = __uint_as_float(v1[c]);
          dv[c] = __uint_as_float(v2[c]);
        }
      }
      tc_fence_before();
      mbar_arrive(&sp_empty[s]);
      const int c0 = j * N + half * NH;
      float mult[NH];
#pragma unroll
      for (int c = 0; c < NH; ++c) mult[c] = 1.f;
      if (a.drop.p > 0.f) {
        if constexpr (KV) {  // thread = key row, columns = queries: one mask row per column
#pragma unroll
          for (int c = 0; c < NH; ++c)
            if (c0 + c < a.T && ri < a.T)
              mult[c] = dropout_mult(a.drop.seed, a.drop.site, ((uint64_t)bh * a.T + (c0 + c)) * (uint64_t)Tc + ri,
                                     a.drop.p, inv_keep);
        } else {             // thread = query row, columns = keys: 4 adjacent keys share one Philox call
          const uint64_t rowb = ((uint64_t)bh * a.T + (uint64_t)min(ri, a.T - 1)) * (uint64_t)Tc;
#pragma unroll
          for (int g = 0; g < NH / 4; ++g) {
            float mk[4];
            dropout_mult4(a.drop.seed, a.drop.site, (rowb + (uint64_t)(c0 + 4 * g)) >> 2, a.drop.p, inv_keep, mk);
#pragma unroll
            for (int e = 0; e < 4; ++e) mult[4 * g + e] = mk[e];
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NH; ++c) {
        const int ci = c0 + c;  // query index (KV) or key index
        const bool valid = (ci < a.T) && (ri < a.T);
        const float l2 = KV ? (ci < a.T ? __ldg(lse + ci) : 0.f) : lse_r;
        const float dl = KV ? (ci < a.T ? __ldg(delta + ci) : 0.f) : delta_r;
        const float p = valid ? fast_exp2(fmaf(sv[c], a.scale_log2, -l2)) : 0.f;
        dv[c] = p * (dv[c] * mult[c] - dl);  // dS'
        sv[c] = p * mult[c];                 // Pd'
      }
      mbar_wait(ps_empty, (j & 1) ^ 1);
#pragma unroll
      for (int hsel = KV ? 0 : 1; hsel < 2; ++hsel) {
#pragma unroll
        for (int ch = 0; ch < NH / 8; ++ch) {
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = hsel ? dv[ch * 8 + e] : sv[ch * 8 + e];
          uint4 hi, lo;
          split8(x, hi, lo);
          const int kel = hsel * N + half * NH + ch * 8;
          const uint32_t off = (kel >> 5) * 8192 + sw64_offset(row, (kel >> 3) & 3);
          *reinterpret_cast<uint4*>(smem + L.ps_hi + off) = hi;
          if (a.x3) *reinterpret_cast<uint4*>(smem + L.ps_lo + off) = lo;
        }
      }
      fence_proxy_async();
      mbar_arrive(ps_full);
    }
    // ---- epilogue: accumulators -> d_qkv (fp32, packed [B,T,3*H*E]); each half writes AD*16 columns
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int I = a.H * a.E;
    float* grow = a.dqkv + ((int64_t)b * a.T + ri) * (3 * I) + h * a.E;
#pragma unroll
    for (int which = KV ? 0 : 1; which < 2; ++which) {
      // KV: out1 = dV (col block 2I), out2 = dK (col block I, scaled);  !KV: out2 = dQ (col block 0, scaled)
      float* dst = grow + (KV ? (which ? I : 2 * I) : 0);
      const float sc = which ? a.scale : 1.f;
      const uint32_t tm = which ? tm_o2 : tm_o1;
#pragma unroll
      for (int cc = 0; cc < AD; ++cc) {
        const int d0 = half * (AD * 16) + cc * 16;
        uint32_t v[16];
        tmem_ld16(tm + lane_off + d0, v);
        tmem_ld_wait();
        if (ri < a.T) {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (d0 + c < a.E) dst[d0 + c] = __uint_as_float(v[c]) * sc;
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

template <bool KV, int N, int AD>
int launch_bwd(const AttnBwdArgs& a, cudaStream_t st) {
  const BwdSmem L = bwd_smem_layout(AD * 32, N, a.x3);
  V1T_CHECK_ARG(L.total <= 232448, "attn_bwd_tc: shared memory budget exceeded (%u bytes)", L.total);
  V1T_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<KV, N, AD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  dim3 grid(cdiv(a.T, 128), a.B * a.H);
  attn_bwd_kernel<KV, N, AD><<<grid, kThreadsAttn, L.total, st>>>(a);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

template <int AD>
int bwd_both(const AttnBwdArgs& a, cudaStream_t st) {
  const char* force = getenv("V1T_BWD_N");  // debug: force the streamed tile size
  if (force && atoi(force) == 32) {
    V1T_TRY((launch_bwd<true, 32, AD>(a, st)));
    return launch_bwd<false, 32, AD>(a, st);
  }
  if (force && atoi(force) == 16) {
    V1T_TRY((launch_bwd<true, 16, AD>(a, st)));
    return launch_bwd<false, 16, AD>(a, st);
  }
  if (a.x3) {
    V1T_TRY((launch_bwd<true, 16, AD>(a, st)));
    return launch_bwd<false, 16, AD>(a, st);
  }
  V1T_TRY((launch_bwd<true, 32, AD>(a, st)));
  return launch_bwd<false, 32, AD>(a, st);
}

}  // namespace

#define V1T_AD_DISPATCH(ad, CALL) \
  switch (ad) {                   \
    case 1: return CALL(1);       \
    case 2: return CALL(2);       \
    case 3: return CALL(3);       \
    case 4: return CALL(4);       \
    default: return CALL(5);      \
  }

int attn_fwd_tc(const AttnFwdArgs& a, cudaStream_t st) {
  V1T_CHECK_ARG(a.Dp % 32 == 0 && a.Dp >= 32 && a.Dp <= 160 && a.Tp % 128 == 0 && a.Tp >= a.T && a.E <= a.Dp,
                "attn_fwd_tc: unsupported dims (Dp %d, Tp %d)", a.Dp, a.Tp);
  V1T_CHECK_ARG(a.B * a.H <= 65535, "attn_fwd_tc: too many (batch, head) pairs");
#define CALL(AD) launch_fwd<AD>(a, st)
  V1T_AD_DISPATCH(a.Dp / 32, CALL)
#undef CALL
}

int attn_bwd_tc(const AttnBwdArgs& a, cudaStream_t st) {
  V1T_CHECK_ARG(a.Dp % 32 == 0 && a.Dp >= 32 && a.Dp <= 160 && a.Tp % 128 == 0 && a.Tp >= a.T && a.E <= a.Dp,
                "attn_bwd_tc: unsupported dims (Dp %d, Tp %d)", a.Dp, a.Tp);
  V1T_CHECK_ARG(a.B * a.H <= 65535, "attn_bwd_tc: too many (batch, head) pairs");
#define CALL(AD) bwd_both<AD>(a, st)
  V1T_AD_DISPATCH(a.Dp / 32, CALL)
#undef CALL
}

}  // namespace v1t

// ---------------------------------------------------------------------------------------------------------
// C-ABI: fused attention on a packed fp32 qkv tensor [B, T, 3*H*E] (the layout to_qkv produces, vit.py:269)
// ---------------------------------------------------------------------------------------------------------
namespace v1t {
AttnPlanes carve_attn_planes(void* base, int B, int H, int Tp, int Dp, bool with_backward) {
  AttnPlanes p{};
  char* c = (char*)base;
  size_t off = 0;
  const size_t pb = (size_t)round_up((int64_t)plane_bytes(B, H, Tp, Dp), 1024);
  auto take = [&]() {
    uint8_t* q = c ? (uint8_t*)(c + off) : nullptr;
    off += pb;
    return q;
  };
  // V^T (forward, token-contiguous) and V (backward, head-dim-contiguous) are never live together: one slot
  for (int i = 0; i < 2; ++i) { p.q[i] = take(); p.k[i] = take(); p.vt[i] = take(); p.v[i] = p.vt[i]; }
  if (with_backward) {
    for (int i = 0; i < 2; ++i) p.dO[i] = take();
  }
  p.lse = (float*)(c ? c + off : nullptr);
  off += (size_t)round_up((int64_t)B * H * Tp * 4, 1024);
  p.delta = (float*)(c ? c + off : nullptr);
  off += (size_t)round_up((int64_t)B * H * Tp * 4, 1024);
  p.total = off;
  return p;
}
}  // namespace v1t

namespace v1t {
int attn_fwd_dispatch(const AttnFwdArgs& a, cudaStream_t st) {
  const char* gen = getenv("V1T_ATTN_FWD");  // "1" selects the first-generation (smem-resident) kernel
  if (gen && atoi(gen) == 1) return attn_fwd_tc(a, st);
  return attn_fwd2_tc(a, st);
}
int attn_bwd_dispatch(const AttnBwdArgs& a, cudaStream_t st) {
  const char* gen = getenv("V1T_ATTN_BWD");  // "1" selects the first-generation (smem-resident) kernels
  if (gen && atoi(gen) == 1) return attn_bwd_tc(a, st);
  return attn_bwd2_tc(a, st);
}
}  // namespace v1t

extern "C" size_t v1t_attn_scratch_bytes(int B, int H, int T, int E) {
  const int Tp = (int)v1t::round_up(T, 128), Dp = (int)v1t::round_up(E, 32);
  return v1t::carve_attn_planes(nullptr, B, H, Tp, Dp, true).total;
}

extern "C" int v1t_attn_forward(const float* qkv, int B, int H, int T, int E, int impl, float p_drop, uint64_t seed,
                                uint32_t site, float* out, float* lse_out, void* scratch, void* stream) {
  using namespace v1t;
  V1T_CHECK_ARG(qkv && out && scratch && B > 0 && H > 0 && T > 0 && E > 0, "attn_forward: bad argument");
  V1T_CHECK_ARG(impl == V1T_IMPL_BF16X3 || impl == V1T_IMPL_BF16, "attn_forward: impl must be BF16X3 or BF16");
  V1T_CHECK_ARG(E <= 160, "attn_forward: fused kernel supports head dim <= 160 (got %d)", E);
  cudaStream_t st = (cudaStream_t)stream;
  const int Tp = (int)round_up(T, 128), Dp = (int)round_up(E, 32);
  const int I = H * E;
  const int x3 = impl == V1T_IMPL_BF16X3;
  AttnPlanes p = carve_attn_planes(scratch, B, H, Tp, Dp, true);
  V1T_TRY(make_planes(qkv, 3 * I, 0, B, H, T, Tp, E, Dp, p.q[0], x3 ? p.q[1] : nullptr, nullptr, nullptr, st));
  V1T_TRY(make_planes(qkv, 3 * I, I, B, H, T, Tp, E, Dp, p.k[0], x3 ? p.k[1] : nullptr, nullptr, nullptr, st));
  V1T_TRY(make_planes(qkv, 3 * I, 2 * I, B, H, T, Tp, E, Dp, nullptr, nullptr, p.vt[0], x3 ? p.vt[1] : nullptr, st));
  AttnFwdArgs a{};
  a.q_hi = p.q[0]; a.q_lo = p.q[1]; a.k_hi = p.k[0]; a.k_lo = p.k[1]; a.vt_hi = p.vt[0]; a.vt_lo = p.vt[1];
  a.O = out; a.o_ld = I; a.lse = lse_out ? lse_out : p.lse;
  a.B = B; a.H = H; a.T = T; a.Tp = Tp; a.E = E; a.Dp = Dp;
  a.scale_log2 = (1.0f / sqrtf((float)E)) * 1.4426950408889634f;
  a.x3 = x3;
  a.drop = DropSpec{seed, site, p_drop};
  return attn_fwd_dispatch(a, st);
}

extern "C" int v1t_attn_backward(const float* qkv, const float* out, const float* d_out, const float* lse, int B, int H,
                                 int T, int E, int impl, float p_drop, uint64_t seed, uint32_t site, float* d_qkv,
                                 void* scratch, void* stream) {
  using namespace v1t;
  V1T_CHECK_ARG(qkv && out && d_out && lse && d_qkv && scratch && B > 0 && H > 0 && T > 0 && E > 0,
                "attn_backward: bad argument");
  V1T_CHECK_ARG(impl == V1T_IMPL_BF16X3 || impl == V1T_IMPL_BF16, "attn_backward: impl must be BF16X3 or BF16");
  V1T_CHECK_ARG(E <= 160, "attn_backward: fused kernel supports head dim <= 160 (got %d)", E);
  cudaStream_t st = (cudaStream_t)stream;
  const int Tp = (int)round_up(T, 128), Dp = (int)round_up(E, 32);
  const int I = H * E;
  const int x3 = impl == V1T_IMPL_BF16X3;
  AttnPlanes p = carve_attn_planes(scratch, B, H, Tp, Dp, true);
  V1T_TRY(make_planes(qkv, 3 * I, 0, B, H, T, Tp, E, Dp, p.q[0], x3 ? p.q[1] : nullptr, nullptr, nullptr, st));
  V1T_TRY(make_planes(qkv, 3 * I, I, B, H, T, Tp, E, Dp, p.k[0], x3 ? p.k[1] : nullptr, nullptr, nullptr, st));
  V1T_TRY(make_planes(qkv, 3 * I, 2 * I, B, H, T, Tp, E, Dp, p.v[0], x3 ? p.v[1] : nullptr, nullptr, nullptr, st));
  V1T_TRY(make_planes(d_out, I, 0, B, H, T, Tp, E, Dp, p.dO[0], x3 ? p.dO[1] : nullptr, nullptr, nullptr, st));
  V1T_TRY(attn_delta(out, d_out, p.delta, B, H, T, Tp, E, I, st));
  AttnBwdArgs a{};
  a.q_hi = p.q[0]; a.q_lo = p.q[1]; a.k_hi = p.k[0]; a.k_lo = p.k[1]; a.v_hi = p.v[0]; a.v_lo = p.v[1];
  a.do_hi = p.dO[0]; a.do_lo = p.dO[1];
  a.lse = lse; a.delta = p.delta; a.dqkv = d_qkv;
  a.B = B; a.H = H; a.T = T; a.Tp = Tp; a.E = E; a.Dp = Dp;
  a.scale = 1.0f / sqrtf((float)E);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.x3 = x3;
  a.drop = DropSpec{seed, site, p_drop};
  return attn_bwd_dispatch(a, st);
}
