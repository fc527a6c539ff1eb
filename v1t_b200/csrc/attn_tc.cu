// Fused attention forward on tcgen05 (reference: Attention.scaled_dot_product_attention, vit.py:253-265).
//
//   P = softmax(Q K^T * E^-0.5) (dropout)   O = P V        per (sample, head), T = 1654 tokens, head dim E = 155
//
// One CTA per (b, h, 128-query tile); K / V^T stream through shared memory in 64-key tiles fetched with
// cp.async.bulk from the pre-swizzled bf16 planes (planes.cu); S and O accumulate in TMEM; nothing of size T x T
// ever reaches HBM.  bf16x3 mode multiplies hi/lo split operands (Q, K, P, V all split) -> fp32-class accuracy.
//
// Softmax uses a TWO-PASS scheme instead of online rescaling of the O accumulator in TMEM:
//   pass 1: S = Qhi Khi^T only (1 MMA product) -> per-row reference maximum m (any value near the true max works)
//   pass 2: S (full precision), P = exp2(S*c - m*c), l += rowsum(P), O += P V     (no correction step, no TMEM
//           read-modify-write; costs one extra bf16 QK^T product)
// Warp roles (192 threads): warps 0-3 softmax/epilogue (TMEM lane quarter = warp id), warp 4 TMEM alloc + MMA
// issue, warp 5 bulk-copy producer.  S is double-buffered in TMEM so QK^T of tile j+1 overlaps softmax of tile j.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace v1t {
namespace {

using namespace tc;

constexpr int BQ = 128, BKEY = 64;
constexpr int kSoftmaxWarps = 4;
constexpr int kFwdThreads = (kSoftmaxWarps + 2) * 32;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct FwdSmem {
  uint32_t q_hi, q_lo, k_hi, k_lo, v_hi, v_lo, p_hi, p_lo, bars, total;
};
__host__ __device__ inline FwdSmem fwd_smem_layout(int Dp) {
  FwdSmem s;
  const uint32_t ad = Dp / 32;
  uint32_t o = 0;
  s.q_hi = o; o += ad * BQ * 64;
  s.q_lo = o; o += ad * BQ * 64;
  s.k_hi = o; o += ad * BKEY * 64;
  s.k_lo = o; o += ad * BKEY * 64;
  s.v_hi = o; o += 2 * Dp * 64;
  s.v_lo = o; o += 2 * Dp * 64;
  s.p_hi = o; o += 2 * BQ * 64;
  s.p_lo = o; o += 2 * BQ * 64;
  s.bars = o; o += 256;
  s.total = o + 1024;  // + alignment slack
  return s;
}

__global__ void __launch_bounds__(kFwdThreads, 1) attn_fwd_kernel(const AttnFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const FwdSmem L = fwd_smem_layout(a.Dp);
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
  uint64_t* r_full = bars + 12;   // [4] pass-1 K ring (slots = the k_hi, k_lo, v_hi, v_lo regions)
  uint64_t* r_empty = bars + 16;  // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  const uint32_t ring_off[4] = {L.k_hi, L.k_lo, L.v_hi, L.v_lo};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.y;
  const int q0 = qt * BQ;
  const int ad = a.Dp / 32, at = a.Tp / 32;
  const int nk = (a.T + BKEY - 1) / BKEY;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(p_full, kSoftmaxWarps * 32);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], kSoftmaxWarps * 32);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kSoftmaxWarps) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * BKEY;

  if (warp == kSoftmaxWarps + 1) {
    // ============================== BULK-COPY PRODUCER ==============================
    if (lane == 0) {
      const uint32_t q_bytes = ad * BQ * 64;
      mbar_expect_tx(q_full, a.x3 ? 2 * q_bytes : q_bytes);
      for (int at_i = 0; at_i < ad; ++at_i) {
        const int64_t src = (((int64_t)bh * ad + at_i) * a.Tp + q0) * 64;
        bulk_g2s(smem + L.q_hi + at_i * BQ * 64, a.q_hi + src, BQ * 64, q_full);
        if (a.x3) bulk_g2s(smem + L.q_lo + at_i * BQ * 64, a.q_lo + src, BQ * 64, q_full);
      }
      const uint32_t k_bytes = ad * BKEY * 64, v_bytes = 2 * a.Dp * 64;
      auto load_k = [&](uint32_t dst_hi, uint32_t dst_lo, int j, bool lo, uint64_t* bar) {
        mbar_expect_tx(bar, lo ? 2 * k_bytes : k_bytes);
        for (int at_i = 0; at_i < ad; ++at_i) {
          const int64_t src = (((int64_t)bh * ad + at_i) * a.Tp + j * BKEY) * 64;
          bulk_g2s(smem + dst_hi + at_i * BKEY * 64, a.k_hi + src, BKEY * 64, bar);
          if (lo) bulk_g2s(smem + dst_lo + at_i * BKEY * 64, a.k_lo + src, BKEY * 64, bar);
        }
      };
      // pass 1: hi planes of K only, 4-slot ring over the (still unused) K/V regions
      for (int j = 0; j < nk; ++j) {
        const int slot = j & 3;
        mbar_wait(&r_empty[slot], ((j >> 2) & 1) ^ 1);
        load_k(ring_off[slot], 0, j, false, &r_full[slot]);
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
        for (int ka = 0; ka < 2; ++ka) {
          const int64_t src = (((int64_t)bh * at + (j * 2 + ka)) * a.Dp) * 64;
          bulk_g2s(smem + L.v_hi + ka * a.Dp * 64, a.vt_hi + src, a.Dp * 64, v_full);
          if (a.x3) bulk_g2s(smem + L.v_lo + ka * a.Dp * 64, a.vt_lo + src, a.Dp * 64, v_full);
        }
      }
    }
  } else if (warp == kSoftmaxWarps) {
    // ============================== MMA ISSUER ==============================
    const uint32_t idesc_s = idesc_bf16(BQ, BKEY, 0, 0);
    const uint32_t idesc_o = idesc_bf16(BQ, a.Dp, 0, 0);
    const uint32_t sq_hi = smem_u32(smem + L.q_hi), sq_lo = smem_u32(smem + L.q_lo);
    const uint32_t sk_hi = smem_u32(smem + L.k_hi), sk_lo = smem_u32(smem + L.k_lo);
    const uint32_t sv_hi = smem_u32(smem + L.v_hi), sv_lo = smem_u32(smem + L.v_lo);
    const uint32_t sp_hi = smem_u32(smem + L.p_hi), sp_lo = smem_u32(smem + L.p_lo);
    uint32_t itk = 0, its = 0;

    // S[buf] = Q K^T from the K tile at smem offsets (k_hi_addr, k_lo_addr); commits `k_done` and s_full[buf]
    auto issue_s = [&](uint32_t k_hi_addr, uint32_t k_lo_addr, bool full_precision, uint64_t* k_ready,
                       uint32_t k_parity, uint64_t* k_done) {
      const uint32_t buf = its & 1;
      mbar_wait(k_ready, k_parity);
      mbar_wait(&s_empty[buf], ((its >> 1) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d = tmem_base + buf * BKEY;
        for (int ks = 0; ks < a.Dp / 16; ++ks) {
          const uint32_t qo = (ks >> 1) * (BQ * 64) + (ks & 1) * 32, ko = (ks >> 1) * (BKEY * 64) + (ks & 1) * 32;
          const uint64_t qh = desc_k_sw64(sq_hi + qo), kh = desc_k_sw64(k_hi_addr + ko);
          umma_bf16(d, qh, kh, idesc_s, ks > 0 ? 1u : 0u);
          if (full_precision) {
            const uint64_t ql = desc_k_sw64(sq_lo + qo), kl = desc_k_sw64(k_lo_addr + ko);
            umma_bf16(d, ql, kh, idesc_s, 1u);
            umma_bf16(d, qh, kl, idesc_s, 1u);
          }
        }
        umma_commit(k_done);
        umma_commit(&s_full[buf]);
      }
      __syncwarp();
      ++its;
    };
    auto issue_s2 = [&]() {  // pass-2 tile from the K buffer
      issue_s(sk_hi, sk_lo, a.x3 != 0, k_full, itk & 1, k_empty);
      ++itk;
    };

    mbar_wait(q_full, 0);
    for (int j = 0; j < nk; ++j) {  // pass 1: reference row max from the hi planes
      const int slot = j & 3;
      issue_s(smem_u32(smem + ring_off[slot]), 0, false, &r_full[slot], (j >> 2) & 1, &r_empty[slot]);
    }
    issue_s2();  // pass 2, tile 0
    for (int j = 0; j < nk; ++j) {
      if (j + 1 < nk) issue_s2();                 // S(j+1) overlaps softmax(j)
      mbar_wait(p_full, j & 1);
      mbar_wait(v_full, j & 1);
      tc_fence_after();
      if (lane == 0) {
        for (int ks = 0; ks < BKEY / 16; ++ks) {
          const uint32_t po = (ks >> 1) * (BQ * 64) + (ks & 1) * 32, vo = (ks >> 1) * (a.Dp * 64) + (ks & 1) * 32;
          const uint64_t ph = desc_k_sw64(sp_hi + po), vh = desc_k_sw64(sv_hi + vo);
          umma_bf16(tmem_o, ph, vh, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          if (a.x3) {
            const uint64_t pl = desc_k_sw64(sp_lo + po), vl = desc_k_sw64(sv_lo + vo);
            umma_bf16(tmem_o, pl, vh, idesc_o, 1u);
            umma_bf16(tmem_o, ph, vl, idesc_o, 1u);
          }
        }
        umma_commit(p_empty);
        umma_commit(v_empty);
        if (j == nk - 1) umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ============================== SOFTMAX / EPILOGUE ==============================
    const int row = warp * 32 + lane;          // TMEM lane = query row of the tile
    const int qi = q0 + row;                   // token index
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const int b = bh / a.H, h = bh % a.H;
    uint32_t its = 0;
    float m = -INFINITY;
    // ---- pass 1: row max
    for (int j = 0; j < nk; ++j, ++its) {
      const uint32_t buf = its & 1;
      mbar_wait(&s_full[buf], (its >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tmem_ld32(tmem_base + lane_off + buf * BKEY + half * 32, v);
        tmem_ld_wait();
        const int jb = j * BKEY + half * 32;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (jb + c < a.T) m = fmaxf(m, __uint_as_float(v[c]));
      }
      tc_fence_before();
      mbar_arrive(&s_empty[buf]);
    }
    const float m2 = m * a.scale_log2;
    float l = 0.f;
    const float inv_keep = a.drop.p > 0.f ? 1.f / (1.f - a.drop.p) : 1.f;
    const uint64_t drop_row = ((uint64_t)bh * a.T + (uint64_t)min(qi, a.T - 1)) * (uint64_t)a.T;
    // ---- pass 2: P = exp2(S*c - m*c), l += rowsum(P), P (dropout) -> smem as the A operand of P V
    for (int j = 0; j < nk; ++j, ++its) {
      const uint32_t buf = its & 1;
      mbar_wait(&s_full[buf], (its >> 1) & 1);
      tc_fence_after();
      float p[64];
      {
        uint32_t v[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld32(tmem_base + lane_off + buf * BKEY + half * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) p[half * 32 + c] = __uint_as_float(v[c]);
        }
      }
      tc_fence_before();
      mbar_arrive(&s_empty[buf]);
      const int jb = j * BKEY;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        const float e = (jb + c < a.T) ? exp2f(fmaf(p[c], a.scale_log2, -m2)) : 0.f;
        l += e;
        p[c] = e;
      }
      if (a.drop.p > 0.f) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          p[c] *= dropout_mult(a.drop.seed, a.drop.site, drop_row + (uint64_t)(jb + c), a.drop.p, inv_keep);
      }
      mbar_wait(p_empty, (j & 1) ^ 1);  // P V of the previous tile has consumed the buffer
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = p[ch * 8 + e];
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = (ch >> 2) * (BQ * 64) + sw64_offset(row, ch & 3);
        *reinterpret_cast<uint4*>(smem + L.p_hi + off) = hi;
        if (a.x3) *reinterpret_cast<uint4*>(smem + L.p_lo + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> global, log-sum-exp (base 2) for the backward
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.f / l;
    float* orow = a.O + ((int64_t)b * a.T + qi) * a.o_ld + h * a.E;
    for (int c0 = 0; c0 < a.Dp; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_o + lane_off + c0, v);
      tmem_ld_wait();
      if (qi < a.T) {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c0 + c < a.E) orow[c0 + c] = __uint_as_float(v[c]) * inv_l;
      }
    }
    if (qi < a.T && a.lse) a.lse[(int64_t)bh * a.Tp + qi] = m2 + log2f(l);
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kSoftmaxWarps) tmem_dealloc<512>(tmem_base);
}

}  // namespace

int attn_fwd_tc(const AttnFwdArgs& a, cudaStream_t st) {
  V1T_CHECK_ARG(a.Dp % 32 == 0 && a.Dp >= 32 && a.Dp <= 160 && a.Tp % 128 == 0 && a.Tp >= a.T && a.E <= a.Dp,
                "attn_fwd_tc: unsupported dims (Dp %d, Tp %d)", a.Dp, a.Tp);
  const FwdSmem L = fwd_smem_layout(a.Dp);
  static int attr_smem = 0;
  if ((int)L.total > attr_smem) {
    V1T_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    attr_smem = (int)L.total;
  }
  dim3 grid(cdiv(a.T, BQ), a.B * a.H);
  V1T_CHECK_ARG(grid.y <= 65535, "attn_fwd_tc: too many (batch, head) pairs");
  attn_fwd_kernel<<<grid, kFwdThreads, L.total, st>>>(a);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
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
  for (int i = 0; i < 2; ++i) { p.q[i] = take(); p.k[i] = take(); p.vt[i] = take(); }
  if (with_backward) {
    for (int i = 0; i < 2; ++i) { p.v[i] = take(); p.qt[i] = take(); p.kt[i] = take(); p.dO[i] = take(); p.dOt[i] = take(); }
  }
  p.lse = (float*)(c ? c + off : nullptr);
  off += (size_t)round_up((int64_t)B * H * Tp * 4, 1024);
  p.delta = (float*)(c ? c + off : nullptr);
  off += (size_t)round_up((int64_t)B * H * Tp * 4, 1024);
  p.total = off;
  return p;
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
  return attn_fwd_tc(a, st);
}
