// Attention forward, second generation: Q (the 128-row resident operand) and P (the A operand of P V) live in
// TENSOR MEMORY, so every tcgen05.mma runs in its TS form (~10 + N/2 cycles instead of 43 + N/2, measured by
// scripts/mma_microbench.py) and shared memory only holds the streamed K and V tiles, each in its own
// double-buffered ring (K(j) is free after S(j), V(j) after P V(j)), which hides the bulk-copy latency.  V comes
// from the same token-major planes as K and is read through an MN-major descriptor (N = head dim is the
// contiguous direction), so no transposed copy of V is ever made.
// Two-pass softmax: pass 1 = row maxima from the hi planes only (any value near the true maximum works as the
// reference point), pass 2 = exact scores, P = exp2, row sums, O += P V (no rescaling of the accumulator in TMEM).
//
// TMEM columns (Dp = 160): Q_hi[80] | Q_lo[80] | S[2][64] | P_hi[32] | P_lo[32] | O[160]  = 512.
// Warp roles (608 threads): warps 0-15 softmax/epilogue (lane quarter = warp & 3, column slot = warp >> 2: 16 of a
// tile's 64 key columns each -- with 8 warps the exp/dropout/split chain of a tile did not fit under the MMAs of the
// next one), warp 16 MMA issue + TMEM alloc, warp 17 K-ring producer, warp 18 V-ring producer.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace v1t {
namespace {

using namespace tc;

constexpr int BQ = 128, BKEY = 64;
constexpr int kSmWarps = 16;
constexpr int kSlots = kSmWarps / 4;   // column slots per TMEM lane quarter
constexpr int NW = BKEY / kSlots;      // key columns of a tile per softmax warp
constexpr int kSmThreads = kSmWarps * 32;
constexpr int kThreadsAttn = (kSmWarps + 3) * 32;
constexpr int kMmaWarp = kSmWarps, kLoadWarpK = kSmWarps + 1, kLoadWarpV = kSmWarps + 2;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int AD>
struct FSmem {
  static constexpr uint32_t kKTile = AD * BKEY * 64;       // one plane of a K tile (64 keys x Dp)
  static constexpr uint32_t kVTile = AD * BKEY * 64;       // one plane of a V tile (64 keys x Dp, MN-major operand)
  // K ring depth: 3 slots were tried (more bytes in flight) and were not faster than 2 -- the copies are not
  // latency-bound
  static constexpr int kKS = 2;
  static constexpr uint32_t k_ring = 0;                    // kKS slots x (hi, lo); pass 1: 4 hi-only slots
  static constexpr uint32_t v_ring = 2 * kKS * kKTile;     // 2 slots x (hi, lo)
  static constexpr uint32_t bars = v_ring + 4 * kVTile;
  static constexpr uint32_t xch = bars + 256;              // cross-half exchange of row max / row sum
  static constexpr uint32_t total = xch + kSlots * BQ * 4 + 1024;
  // EMIT instantiation: per-warp 32 x 17 fp32 transpose tiles behind everything else (the rings are too small for them at
  // small head dims)
  static constexpr uint32_t emit_tiles = xch + kSlots * BQ * 4;
  static constexpr uint32_t total_emit = emit_tiles + kSmWarps * 32 * 17 * 4 + 1024;
};

template <int AD, bool EMIT>
__global__ void __launch_bounds__(kThreadsAttn, 1) attn_fwd2_kernel(const AttnFwdArgs a) {
  constexpr int Dp = AD * 32, HC = AD * 16;
  using L = FSmem<AD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  uint64_t* q_ready = bars + 0;   // Q rows stored to TMEM by the 256 softmax threads
  uint64_t* p_full = bars + 1;
  uint64_t* p_empty = bars + 2;
  uint64_t* o_full = bars + 3;
  uint64_t* k_full = bars + 4;    // [kKS <= 4]
  uint64_t* k_empty = bars + 8;   // [kKS <= 4]
  uint64_t* v_full = bars + 12;   // [2]
  uint64_t* v_empty = bars + 14;  // [2]
  uint64_t* s_full = bars + 16;   // [2]
  uint64_t* s_empty = bars + 18;  // [2]
  uint64_t* r_full = bars + 20;   // [4] pass-1 ring
  uint64_t* r_empty = bars + 24;  // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  uint64_t* stage_full = bars + 30;  // the 128 Q rows have landed in the V ring (bulk copy)
  constexpr int kKS = L::kKS;
  float* xch = reinterpret_cast<float*>(smem + L::xch);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.y;
  const int q0 = qt * BQ;
  const int at = a.Tp / 32;
  const int nk = (a.T + BKEY - 1) / BKEY;
  const bool p_lo = a.x3 && !(a.prec & 1);  // P_lo * V_hi term of P V (precision-budget experiment, kernels.cuh)

  if (threadIdx.x == 0) {
    mbar_init(q_ready, kSmThreads);
    mbar_init(stage_full, 1);
    mbar_init(p_full, kSmThreads);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], kSmThreads);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], 1);
    }
    for (int i = 0; i < kKS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t cQ_hi = 0, cQ_lo = HC, cS = 2 * HC, cP_hi = cS + 2 * BKEY, cP_lo = cP_hi + BKEY / 2,
                     cO = cP_lo + BKEY / 2;
  static_assert(cO + Dp <= 512, "TMEM budget exceeded");

  if (warp == kLoadWarpK) {
    // ============================== K PRODUCER ==============================
    if (lane == 0) {
      auto load_k = [&](uint32_t dst_hi, uint32_t dst_lo, int j, bool lo, uint64_t* bar) {
        mbar_expect_tx(bar, lo ? 2 * L::kKTile : L::kKTile);
        // the AD atoms of a 64-key tile are contiguous in the plane: ONE copy per plane ([atom][64 keys][64 B])
        const int64_t src = attn_plane_off(bh, 0, j * BKEY, a.Tp, AD);
        bulk_g2s(smem + dst_hi, a.k_hi + src, L::kKTile, bar);
        if (lo) bulk_g2s(smem + dst_lo, a.k_lo + src, L::kKTile, bar);
      };
      // pass 1: hi planes only, 4-slot ring over the K region
      for (int j = 0; j < nk && !EMIT; ++j) {
        const int slot = j & 3;
        mbar_wait(&r_empty[slot], ((j >> 2) & 1) ^ 1);
        load_k(L::k_ring + slot * L::kKTile, 0, j, false, &r_full[slot]);
      }
      for (int slot = 0; slot < 4 && slot < nk && !EMIT; ++slot) {  // all pass-1 MMAs have released their slots
        const int last = ((nk - 1 - slot) / 4) * 4 + slot;
        mbar_wait(&r_empty[slot], (last >> 2) & 1);
      }
      // pass 2: 2 slots of (hi, lo)
      for (int j = 0; j < nk; ++j) {
        const int s = j % kKS;
        mbar_wait(&k_empty[s], ((j / kKS) & 1) ^ 1);
        load_k(L::k_ring + s * 2 * L::kKTile, L::k_ring + s * 2 * L::kKTile + L::kKTile, j, a.x3 != 0, &k_full[s]);
      }
    }
  } else if (warp == kLoadWarpV) {
    // ============================== V^T PRODUCER ==============================
    if (lane == 0) {
      {  // the V ring is idle during pass 1: it first carries the 128 Q rows (two consecutive 64-row plane tiles = one
         // contiguous copy per plane) to the softmax warps, which move them into tensor memory
        const int64_t qoff = attn_plane_off(bh, 0, q0, a.Tp, AD);
        mbar_expect_tx(stage_full, (a.x3 ? 2 : 1) * 2 * L::kVTile);
        bulk_g2s(smem + L::v_ring, a.q_hi + qoff, 2 * L::kVTile, stage_full);
        if (a.x3) bulk_g2s(smem + L::v_ring + 2 * L::kVTile, a.q_lo + qoff, 2 * L::kVTile, stage_full);
        mbar_wait(q_ready, 0);  // every row has been read out of the ring
      }
      for (int j = 0; j < nk && !EMIT; ++j) {
        const int s = j & 1;
        mbar_wait(&v_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[s], a.x3 ? 2 * L::kVTile : L::kVTile);
        uint8_t* base = smem + L::v_ring + s * 2 * L::kVTile;
        const int64_t src = attn_plane_off(bh, 0, j * BKEY, a.Tp, AD);  // [atom][64 keys][64 B], one copy per plane
        bulk_g2s(base, a.v_hi + src, L::kVTile, &v_full[s]);
        if (a.x3) bulk_g2s(base + L::kVTile, a.v_lo + src, L::kVTile, &v_full[s]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA ISSUER ==============================
    const bool leader = elect_one();
    const uint32_t idesc_s = idesc_bf16(BQ, BKEY, 0, 0);
    const uint32_t idesc_o = idesc_bf16(BQ, Dp, 0, 1);  // B = V is MN-major (head dim contiguous)
    const uint32_t kr0 = smem_u32(smem + L::k_ring) >> 4, vr0 = smem_u32(smem + L::v_ring) >> 4;
    const uint32_t tQ_hi = tmem_base + cQ_hi, tQ_lo = tmem_base + cQ_lo;
    const uint32_t tP_hi = tmem_base + cP_hi, tP_lo = tmem_base + cP_lo, tO = tmem_base + cO;
    uint32_t its = 0;

    auto issue_s = [&](uint64_t kh, uint64_t kl, bool full_precision, uint64_t* k_ready, uint32_t k_parity,
                       uint64_t* k_done) {
      const uint32_t buf = its & 1;
      mbar_wait(k_ready, k_parity);
      mbar_wait(&s_empty[buf], ((its >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + cS + buf * BKEY;
#pragma unroll
      for (int ks = 0; ks < 2 * AD; ++ks) {
        constexpr uint32_t kKA = BKEY * 64 / 16;
        const uint32_t ko = (ks >> 1) * kKA + (ks & 1) * 2;
        if (leader) {
          umma_bf16_ts(d, tQ_hi + ks * 8, kh + ko, idesc_s, ks > 0 ? 1u : 0u);
          if (full_precision) {
            umma_bf16_ts(d, tQ_lo + ks * 8, kh + ko, idesc_s, 1u);
            umma_bf16_ts(d, tQ_hi + ks * 8, kl + ko, idesc_s, 1u);
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
    auto issue_s2 = [&](int j) {
      const int s = j % kKS;
      const uint64_t kh = kDescK64 | (uint64_t)(kr0 + s * 2 * (L::kKTile >> 4));
      issue_s(kh, kh + (L::kKTile >> 4), a.x3 != 0, &k_full[s], (j / kKS) & 1, &k_empty[s]);
    };
    auto issue_pv = [&](int j, bool last) {
      const int s = j & 1;
      mbar_wait(&v_full[s], (j >> 1) & 1);
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      // MN-major SWIZZLE_64B: head-dim atoms BKEY*64 bytes apart (LBO), 16 keys = 16 rows of 64 bytes per MMA
      const uint64_t vh = desc_mn_sw64_base(BKEY * 64) | (uint64_t)(vr0 + s * 2 * (L::kVTile >> 4)), vl = vh + (L::kVTile >> 4);
#pragma unroll
      for (int ks = 0; ks < BKEY / 16; ++ks) {
        const uint32_t vo = ks * (16 * 64 / 16);
        if (leader) {
          umma_bf16_ts(tO, tP_hi + ks * 8, vh + vo, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          if (p_lo) umma_bf16_ts(tO, tP_lo + ks * 8, vh + vo, idesc_o, 1u);
          if (a.x3) umma_bf16_ts(tO, tP_hi + ks * 8, vl + vo, idesc_o, 1u);
        }
      }
      if (leader) {
        umma_commit(p_empty);
        umma_commit(&v_empty[s]);
        if (last) umma_commit(o_full);
      }
      __syncwarp();
    };

    mbar_wait(q_ready, 0);
    tc_fence_after();
    if constexpr (EMIT) {
      for (int j = 0; j < nk; ++j) issue_s2(j);  // exact scores only
    } else {
      for (int j = 0; j < nk; ++j) {  // pass 1: reference row max from the hi planes
        const int slot = j & 3;
        issue_s(kDescK64 | (uint64_t)(kr0 + slot * (L::kKTile >> 4)), 0, false, &r_full[slot], (j >> 2) & 1, &r_empty[slot]);
      }
      issue_s2(0);
      for (int j = 0; j + 1 < nk; ++j) {
        issue_s2(j + 1);  // S(j+1) overlaps softmax(j)
        issue_pv(j, false);
      }
      issue_pv(nk - 1, true);
    }
  } else {
    // ============================== SOFTMAX / EPILOGUE ==============================
    const int quarter = warp & 3, slot = warp >> 2;
    const int row = quarter * 32 + lane;
    const int qi = q0 + row;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int b = bh / a.H, h = bh % a.H;
    // ---- Q rows -> TMEM: slot s takes plane (s & 1) and the head-dim atoms of parity (s >> 1)
    {
      const bool lo = (slot & 1) != 0;
      mbar_wait(stage_full, 0);
      if (!lo || a.x3)
        smem_row_to_tmem<AD>(smem_u32(smem + L::v_ring) + (lo ? 2 * L::kVTile : 0), row, tmem_base + lane_off + (lo ? cQ_lo : cQ_hi),
                             slot >> 1, 2);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(q_ready);
    }
    if constexpr (EMIT) {
      // ---- probabilities only: P = exp2(S c - lse) with the lse a forward saved; each warp transposes its 32 x 16 piece
      // through a private shared-memory tile so that half a warp writes 64 contiguous bytes of one row of the [T, T] matrix
      const float lse_r = a.lse_in[(int64_t)bh * a.Tp + qi];
      float* tile = reinterpret_cast<float*>(smem + L::emit_tiles) + warp * (32 * 17);
      float* prow = a.probs + ((int64_t)bh * a.T + q0 + quarter * 32) * a.T;
      for (int j = 0; j < nk; ++j) {
        const uint32_t buf = j & 1;
        const int jb = j * BKEY + slot * NW;
        mbar_wait(&s_full[buf], (j >> 1) & 1);
        tc_fence_after();
        uint32_t v[NW];
        tmem_ld16(tmem_base + lane_off + cS + buf * BKEY + slot * NW, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_empty[buf]);
#pragma unroll
        for (int c = 0; c < NW; ++c) tile[lane * 17 + c] = fast_exp2(fmaf(__uint_as_float(v[c]), a.scale_log2, -lse_r));
        __syncwarp();
        const int cc = lane & 15, rr = lane >> 4;
#pragma unroll
        for (int r2 = 0; r2 < 32; r2 += 2) {
          const int r = r2 + rr;
          if (q0 + quarter * 32 + r < a.T && jb + cc < a.T) prow[(int64_t)r * a.T + jb + cc] = tile[r * 17 + cc];
        }
        __syncwarp();
      }
      tc_fence_before();
    } else {
    uint32_t its = 0;
    float m = -INFINITY;
    // ---- pass 1: row max over this warp's NW of the 64 key columns
    for (int j = 0; j < nk; ++j, ++its) {
      const uint32_t buf = its & 1;
      mbar_wait(&s_full[buf], (its >> 1) & 1);
      tc_fence_after();
      uint32_t v[NW];
      tmem_ld16(tmem_base + lane_off + cS + buf * BKEY + slot * NW, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[buf]);
      const int jb = j * BKEY + slot * NW;
      if (jb + NW <= a.T) {
#pragma unroll
        for (int c = 0; c < NW; ++c) m = fmaxf(m, __uint_as_float(v[c]));
      } else {
#pragma unroll
        for (int c = 0; c < NW; ++c)
          if (jb + c < a.T) m = fmaxf(m, __uint_as_float(v[c]));
      }
    }
    xch[slot * BQ + row] = m;
    named_bar_sync(1, kSmThreads);
#pragma unroll
    for (int s = 0; s < kSlots; ++s) m = fmaxf(m, xch[s * BQ + row]);
    named_bar_sync(1, kSmThreads);
    const float m2 = m * a.scale_log2;
    float l = 0.f;
    const float inv_keep = a.drop.p > 0.f ? 1.f / (1.f - a.drop.p) : 1.f;
    const int Tc = (int)drop_stride(a.T);
    const uint64_t drop_row = ((uint64_t)bh * a.T + (uint64_t)min(qi, a.T - 1)) * (uint64_t)Tc;
    // ---- pass 2
    for (int j = 0; j < nk; ++j, ++its) {
      const uint32_t buf = its & 1;
      const int jb = j * BKEY + slot * NW;
      // the dropout multipliers do not depend on S: draw them BEFORE waiting for the scores, so that the chain from
      // "S(j) complete" to "P(j) in TMEM" (which the P V MMA waits for) holds no Philox latency
      uint32_t keep = 0xffffu;  // bit c: key jb + c is kept
      if (a.drop.p > 0.f) {
        keep = 0;
#pragma unroll
        for (int g = 0; g < NW / 8; ++g)  // one Philox call per 8 adjacent keys
          keep |= dropout_keep8(a.drop.seed, a.drop.site, (drop_row + (uint64_t)(jb + 8 * g)) >> 3, a.drop.p) << (8 * g);
        // the backward reads the decisions instead of re-drawing them (its softmax warps are instruction-bound and the
        // key-major passes would also have to transpose the calls): 16 bits per thread per tile, 22 MB per layer at B = 16
        if (a.drop_bits)
          *reinterpret_cast<uint16_t*>(a.drop_bits + ((int64_t)bh * a.Tp + qi) * (a.Tp >> 3) + (jb >> 3)) = (uint16_t)keep;
      }
      mbar_wait(&s_full[buf], (its >> 1) & 1);
      tc_fence_after();
      float p[NW];
      {
        uint32_t v[NW];
        tmem_ld16(tmem_base + lane_off + cS + buf * BKEY + slot * NW, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < NW; ++c) p[c] = __uint_as_float(v[c]);
      }
      tc_fence_before();
      mbar_arrive(&s_empty[buf]);
      if (jb + NW <= a.T) {
#pragma unroll
        for (int c = 0; c < NW; ++c) {
          p[c] = fast_exp2(fmaf(p[c], a.scale_log2, -m2));
          l += p[c];
        }
      } else {
#pragma unroll
        for (int c = 0; c < NW; ++c) {
          p[c] = (jb + c < a.T) ? fast_exp2(fmaf(p[c], a.scale_log2, -m2)) : 0.f;
          l += p[c];
        }
      }
      if (a.drop.p > 0.f) {
#pragma unroll
        for (int c = 0; c < NW; ++c) p[c] = (keep >> c) & 1u ? p[c] * inv_keep : 0.f;
      }
      mbar_wait(p_empty, (j & 1) ^ 1);  // P V of the previous tile has consumed the operand
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < NW / 8; ++ch) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = p[ch * 8 + e];
        uint32_t hw[4], lw[4];
        split8_words(x, hw, lw);
        const uint32_t col = (slot * NW + ch * 8) / 2;
        tmem_st4(tmem_base + lane_off + cP_hi + col, hw[0], hw[1], hw[2], hw[3]);
        if (p_lo) tmem_st4(tmem_base + lane_off + cP_lo + col, lw[0], lw[1], lw[2], lw[3]);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    xch[slot * BQ + row] = l;
    named_bar_sync(1, kSmThreads);
    l = 0.f;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) l += xch[s * BQ + row];
    // ---- epilogue
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.f / l;
    float* orow = a.O ? a.O + ((int64_t)b * a.T + qi) * a.o_ld + h * a.E : nullptr;
    const int64_t prow = (int64_t)b * a.T + qi;  // row of the head-padded [B*T, H*Dp] output matrix
    const uint32_t stage = smem_u32(smem + L::k_ring);
#pragma unroll
    for (int cc = 0; cc < AD; ++cc) {
      const int c0 = slot * (AD * 8) + cc * 8;  // each slot drains AD*8 of the Dp output columns
      uint32_t v[8];
      tmem_ld8(tmem_base + lane_off + cO + c0, v);
      tmem_ld_wait();
      if (qi < a.T && orow) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c0 + c < a.E) orow[c0 + c] = __uint_as_float(v[c]) * inv_l;
      }
      if (a.o_pl.hi) {
        // operand planes for the projection GEMM and its weight gradient (pad columns are 0): staged in the (now
        // idle) K ring in plane layout -- rows are consecutive plane rows, so each head-dim atom of the tile is ONE
        // contiguous block that a bulk store writes out (full lines instead of 16-byte pieces at a 64-byte stride)
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[e]) * inv_l;
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t so = (uint32_t)(c0 >> 5) * (BQ * 64) + row * 64 + ((((c0 & 31) >> 3) ^ (int)((prow >> 1) & 3)) << 4);
        sts128(stage + so, hi);
        if (a.o_pl.lo) sts128(stage + AD * BQ * 64 + so, lo);
      }
    }
    if (a.o_pl.hi) {
      fence_proxy_async();              // generic-proxy stores above -> visible to the bulk (async proxy) store
      named_bar_sync(1, kSmThreads);
      const int rows_valid = min(BQ, a.T - q0);  // rows past T belong to the next sample of the flat matrix
      if (threadIdx.x < AD * 2 && rows_valid > 0) {
        const int at_i = threadIdx.x >> 1, pln = threadIdx.x & 1;
        uint8_t* dstp = pln ? a.o_pl.lo : a.o_pl.hi;
        if (dstp) {
          const int64_t off = ((int64_t)(h * AD + at_i) * a.o_pl.rows_p + ((int64_t)b * a.T + q0)) * 64;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstp + off),
                       "r"(stage + (uint32_t)(pln * AD + at_i) * (BQ * 64)), "r"(rows_valid * 64)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must stay intact until it was read
        }
      }
    }
    // padded query rows get +inf so that the backward's exp2(S*c - lse) vanishes there without bounds checks
    if (slot == 0 && a.lse) a.lse[(int64_t)bh * a.Tp + qi] = qi < a.T ? m2 + log2f(l) : INFINITY;
    tc_fence_before();
    }  // !EMIT
  }

  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

template <int AD>
int launch_fwd2(const AttnFwdArgs& a, cudaStream_t st) {
  using L = FSmem<AD>;
  static_assert(L::total <= 232448, "shared memory budget exceeded");
  dim3 grid(cdiv(a.T, BQ), a.B * a.H);
  if (a.probs) {
    static_assert(L::total_emit <= 232448, "shared memory budget exceeded");
    V1T_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<AD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total_emit));
    attn_fwd2_kernel<AD, true><<<grid, kThreadsAttn, L::total_emit, st>>>(a);
  } else {
    V1T_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<AD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
    attn_fwd2_kernel<AD, false><<<grid, kThreadsAttn, L::total, st>>>(a);
  }
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

}  // namespace

int attn_emit_probs_tc(const AttnFwdArgs& a, cudaStream_t st) {
  V1T_CHECK_ARG(a.probs && a.lse_in, "attn_emit_probs_tc: probs / lse_in missing");
  return attn_fwd2_tc(a, st);
}

int attn_fwd2_tc(const AttnFwdArgs& a, cudaStream_t st) {
  V1T_CHECK_ARG(a.Dp % 32 == 0 && a.Dp >= 32 && a.Dp <= 160 && a.Tp % 128 == 0 && a.Tp >= a.T && a.E <= a.Dp,
                "attn_fwd2_tc: unsupported dims (Dp %d, Tp %d)", a.Dp, a.Tp);
  V1T_CHECK_ARG(a.B * a.H <= 65535, "attn_fwd2_tc: too many (batch, head) pairs");
  switch (a.Dp / 32) {
    case 1: return launch_fwd2<1>(a, st);
    case 2: return launch_fwd2<2>(a, st);
    case 3: return launch_fwd2<3>(a, st);
    case 4: return launch_fwd2<4>(a, st);
    default: return launch_fwd2<5>(a, st);
  }
}

}  // namespace v1t
