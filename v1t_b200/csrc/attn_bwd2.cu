// Attention backward, second generation: the 128-row RESIDENT operands live in TENSOR MEMORY and feed tcgen05.mma
// in its TS form (A from TMEM costs ~10 + N/2 cycles per MMA instead of 43 + N/2 from shared memory, measured by
// scripts/mma_microbench.py), which also frees shared memory for larger, double-buffered streamed tiles.
//
//   MODE_V  (CTA per 128-key tile)            dV = Pd^T dO                  X = K in TMEM; streams (Q_j, dO_j), N = 64
//   MODE_S, KV = true  (per 128-key tile)     dK = scale * dS^T Q           X = K, Y = V;  streams (Q_j, dO_j), N = 32
//   MODE_S, KV = false (per 128-query tile)   dQ = scale * dS K             X = Q, Y = dO; streams (K_j, V_j),  N = 32
//
// Per streamed tile j:   S' = X x_j^T        (TS: X hi/lo in TMEM)
//                        dP' = Y y_j^T       (TS for Y hi; the Y lo plane stays in shared memory, SS form)
//                        P' = exp2(S' c - lse),  Pd' = P' * dropout,  dS' = P' (dP' * dropout - delta)
//                        out += Pd' y_j (MODE_V)  |  out += dS' x_j (MODE_S)   — A operand (Pd'/dS' as bf16 hi/lo)
//                        written by the softmax warps straight into TMEM (tcgen05.st), B = the streamed tile
//                        re-used un-transposed through an MN-major descriptor.
// Atomic-free, deterministic; P recomputed from the forward's base-2 log-sum-exp.
//
// TMEM columns (HC = Dp/2 columns per 128 x Dp bf16 plane):
//   MODE_S: X_hi | X_lo | Y_hi | S'[N] | dP'[N] | dS'_hi[N/2] | dS'_lo[N/2] | out[Dp]     = 3 HC + 3 N + Dp  (496 @ N=32)
//   MODE_V: X_hi | X_lo | S'[2][N] | Pd'_hi[N/2] | Pd'_lo[N/2] | out[Dp]                  = 2 HC + 3 N + Dp  (512 @ N=64)
// Warp roles (608 threads): warps 0-15 softmax-backward/epilogue (lane quarter = warp & 3, column slot = warp >> 2:
// each warp owns a quarter of a tile's columns; with 8 warps the element-wise work was latency-bound at ~0.4 IPC and
// the tensor pipe sat at 35-45 %), warp 16 MMA issue + TMEM alloc, warps 17-18 bulk-copy producers (one per ring).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace v1t {
namespace {

using namespace tc;

constexpr int kSmWarps = 16;
constexpr int kSlots = kSmWarps / 4;  // column slots per TMEM lane quarter
constexpr int kSmThreads = kSmWarps * 32;
constexpr int kThreadsAttn = (kSmWarps + 3) * 32;
constexpr int kMmaWarp = kSmWarps, kLoadWarpX = kSmWarps + 1, kLoadWarpY = kSmWarps + 2;
constexpr int MODE_V = 0, MODE_S = 1;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int MODE, int N, int AD>
struct Smem2 {
  // The two streamed operands have different lifetimes, so each gets its own ring:
  //   MODE_V: x_j (queries, for S') is free after the score MMAs, y_j (dO, for out) after the output MMAs
  //   MODE_S: y_j (for dP') is free after the score MMAs, x_j (for S' and out) after the output MMAs
  static constexpr int kXS = MODE == MODE_V ? 2 : 4;  // ring slots
  static constexpr int kYS = 2;
  static constexpr uint32_t kTile = AD * N * 64;      // one bf16 plane of one streamed operand tile
  static constexpr uint32_t kSlot = 2 * kTile;        // hi + lo planes
  static constexpr uint32_t kYlo = AD * 128 * 64;     // resident Y lo plane (MODE_S, bf16x3)
  static constexpr uint32_t y_lo = 0;
  static constexpr uint32_t x_ring = MODE == MODE_S ? kYlo : 0;
  static constexpr uint32_t y_ring = x_ring + kXS * kSlot;
  static constexpr uint32_t bars = y_ring + kYS * kSlot;
  static constexpr uint32_t total = bars + 256 + 1024;
};

template <int MODE, bool KV, int N, int AD>
__device__ __forceinline__ void attn_bwd2_body(const AttnBwdArgs& a, const int tile, const int bh) {
  constexpr int Dp = AD * 32, HC = AD * 16;
  constexpr int NH = N / kSlots;  // tile columns per softmax warp
  constexpr int OPC = N / 2;      // TMEM columns of one bf16 plane of the Pd' / dS' operand
  using L = Smem2<MODE, N, AD>;
  constexpr int kXS = L::kXS, kYS = L::kYS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  uint64_t* res_full = bars + 0;   // resident operands ready (TMEM stores by 256 threads + Y_lo bulk copy)
  uint64_t* ps_full = bars + 1;
  uint64_t* ps_empty = bars + 2;
  uint64_t* o_full = bars + 3;
  uint64_t* x_full = bars + 4;     // [kXS]
  uint64_t* x_empty = bars + 8;    // [kXS]
  uint64_t* y_full = bars + 12;    // [kYS]
  uint64_t* y_empty = bars + 14;   // [kYS]
  uint64_t* sp_full = bars + 16;   // [2]
  uint64_t* sp_empty = bars + 18;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = tile * 128;
  // precision-budget experiment (kernels.cuh: attn_prec_env); 0 = all three bf16x3 terms everywhere
  const bool out_lo = a.x3 && !(a.prec & 1);   // A_lo * B_hi term of the accumulating contraction
  const bool dp_ylo = a.x3 && !(a.prec & 6);   // Y_lo * y_hi term of dP' (resident lo plane in shared memory, SS form)
  const bool dp_ysl = a.x3 && !(a.prec & 4);   // Y_hi * y_lo term of dP'
  const int nt = (a.T + N - 1) / N;
  constexpr bool kv_roles = MODE == MODE_V ? true : KV;  // rows = keys, streamed = queries
  const uint8_t* X_hi = kv_roles ? a.k_hi : a.q_hi;   const uint8_t* X_lo = kv_roles ? a.k_lo : a.q_lo;
  const uint8_t* Y_hi = kv_roles ? a.v_hi : a.do_hi;  const uint8_t* Y_lo = kv_roles ? a.v_lo : a.do_lo;
  const uint8_t* xs_hi = kv_roles ? a.q_hi : a.k_hi;  const uint8_t* xs_lo = kv_roles ? a.q_lo : a.k_lo;
  const uint8_t* ys_hi = kv_roles ? a.do_hi : a.v_hi; const uint8_t* ys_lo = kv_roles ? a.do_lo : a.v_lo;

  if (threadIdx.x == 0) {
    mbar_init(res_full, kSmThreads + 1);
    mbar_init(ps_full, kSmThreads);
    mbar_init(ps_empty, 1);
    mbar_init(o_full, 1);
    for (int i = 0; i < kXS; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < kYS; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
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
  // column map
  constexpr uint32_t cX_hi = 0, cX_lo = HC;
  constexpr uint32_t cY_hi = 2 * HC;                                     // MODE_S only
  constexpr uint32_t cS = MODE == MODE_S ? 3 * HC : 2 * HC;              // S' (MODE_V: two buffers of N)
  constexpr uint32_t cDP = cS + N;                                       // MODE_S: dP'
  constexpr uint32_t cPS_hi = cS + 2 * N, cPS_lo = cPS_hi + OPC;         // A operand of the output MMA
  constexpr uint32_t cOut = cPS_lo + OPC;
  static_assert(cOut + Dp <= 512, "TMEM budget exceeded");

  if (warp == kLoadWarpX || warp == kLoadWarpY) {
    // ============================== BULK-COPY PRODUCERS (one warp per ring) ==============================
    if (lane == 0) {
      const bool is_x = warp == kLoadWarpX;
      if (is_x) {  // also fetches the resident Y lo plane
        if (MODE == MODE_S && dp_ylo) {
          mbar_expect_tx(res_full, L::kYlo);
#pragma unroll
          for (int at_i = 0; at_i < AD; ++at_i) {  // resident layout [atom][128 rows][64 B] from two 64-row plane tiles
            bulk_g2s(smem + L::y_lo + at_i * 8192, Y_lo + attn_plane_off(bh, at_i, r0, a.Tp, AD), 4096, res_full);
            bulk_g2s(smem + L::y_lo + at_i * 8192 + 4096, Y_lo + attn_plane_off(bh, at_i, r0 + 64, a.Tp, AD), 4096, res_full);
          }
        } else {
          mbar_arrive(res_full);
        }
      }
      constexpr uint32_t tb = N * 64;
      const int slots = is_x ? kXS : kYS;
      uint64_t* fullb = is_x ? x_full : y_full;
      uint64_t* emptyb = is_x ? x_empty : y_empty;
      const uint8_t* src_hi = is_x ? xs_hi : ys_hi;
      const uint8_t* src_lo = is_x ? xs_lo : ys_lo;
      uint8_t* ring = smem + (is_x ? L::x_ring : L::y_ring);
      for (int j = 0; j < nt; ++j) {
        const int s = j % slots;
        mbar_wait(&emptyb[s], ((j / slots) & 1) ^ 1);
        uint8_t* base = ring + s * L::kSlot;
        mbar_expect_tx(&fullb[s], (a.x3 ? 2 : 1) * L::kTile);
        if constexpr (N == 64) {  // a whole 64-row plane tile: its AD atoms are contiguous, one copy per plane
          const int64_t src = attn_plane_off(bh, 0, j * N, a.Tp, AD);
          bulk_g2s(base, src_hi + src, L::kTile, &fullb[s]);
          if (a.x3) bulk_g2s(base + L::kTile, src_lo + src, L::kTile, &fullb[s]);
        } else {
#pragma unroll
          for (int at_i = 0; at_i < AD; ++at_i) {
            const int64_t src = attn_plane_off(bh, at_i, j * N, a.Tp, AD);
            bulk_g2s(base + at_i * tb, src_hi + src, tb, &fullb[s]);
            if (a.x3) bulk_g2s(base + L::kTile + at_i * tb, src_lo + src, tb, &fullb[s]);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA ISSUER ==============================
    const bool leader = elect_one();
    const uint32_t idesc_s = idesc_bf16(128, N, 0, 0);
    const uint32_t idesc_o = idesc_bf16(128, Dp, 0, 1);  // B = streamed tile viewed MN-major (head dim contiguous)
    constexpr uint32_t tb = N * 64;
    const uint32_t xr0 = smem_u32(smem + L::x_ring) >> 4, yr0 = smem_u32(smem + L::y_ring) >> 4;
    const uint64_t dYlo = kDescK64 | (smem_u32(smem + L::y_lo) >> 4);
    const uint64_t mn_base = desc_mn_sw64_base(tb);
    const uint32_t tX_hi = tmem_base + cX_hi, tX_lo = tmem_base + cX_lo, tY_hi = tmem_base + cY_hi;
    const uint32_t tPS_hi = tmem_base + cPS_hi, tPS_lo = tmem_base + cPS_lo, tOut = tmem_base + cOut;

    auto issue_scores = [&](int j) {
      const int xs = j % kXS, ys = j % kYS;
      const uint32_t xb = xr0 + xs * (L::kSlot >> 4), yb = yr0 + ys * (L::kSlot >> 4);
      const int buf = MODE == MODE_V ? (j & 1) : 0;
      mbar_wait(&x_full[xs], (j / kXS) & 1);
      if (MODE == MODE_S) mbar_wait(&y_full[ys], (j / kYS) & 1);
      if (MODE == MODE_V) mbar_wait(&sp_empty[buf], ((j >> 1) & 1) ^ 1);
      else mbar_wait(&sp_empty[0], (j & 1) ^ 1);
      tc_fence_after();
      const uint64_t xh = kDescK64 | (uint64_t)xb, xl = kDescK64 | (uint64_t)(xb + (L::kTile >> 4));
      const uint64_t yh = kDescK64 | (uint64_t)yb, yl = kDescK64 | (uint64_t)(yb + (L::kTile >> 4));
      const uint32_t dS = tmem_base + cS + buf * N;
#pragma unroll
      for (int ks = 0; ks < 2 * AD; ++ks) {
        const uint32_t bo = (ks >> 1) * (tb / 16) + (ks & 1) * 2;
        if (leader) {
          umma_bf16_ts(dS, tX_hi + ks * 8, xh + bo, idesc_s, ks > 0 ? 1u : 0u);
          if (a.x3) {
            umma_bf16_ts(dS, tX_lo + ks * 8, xh + bo, idesc_s, 1u);
            umma_bf16_ts(dS, tX_hi + ks * 8, xl + bo, idesc_s, 1u);
          }
        }
      }
      if (MODE == MODE_S) {
        const uint32_t dDP = tmem_base + cDP;
#pragma unroll
        for (int ks = 0; ks < 2 * AD; ++ks) {
          const uint32_t bo = (ks >> 1) * (tb / 16) + (ks & 1) * 2, ao = (ks >> 1) * (8192 / 16) + (ks & 1) * 2;
          if (leader) {
            umma_bf16_ts(dDP, tY_hi + ks * 8, yh + bo, idesc_s, ks > 0 ? 1u : 0u);
            if (dp_ylo) umma_bf16(dDP, dYlo + ao, yh + bo, idesc_s, 1u);
            if (dp_ysl) umma_bf16_ts(dDP, tY_hi + ks * 8, yl + bo, idesc_s, 1u);
          }
        }
      }
      if (leader) {
        if (MODE == MODE_V) umma_commit(&x_empty[xs]);   // queries no longer needed
        else umma_commit(&y_empty[ys]);                  // dP' operand no longer needed
        umma_commit(&sp_full[buf]);
      }
      __syncwarp();
    };
    auto issue_out = [&](int j, bool last) {
      const int xs = j % kXS, ys = j % kYS;
      if (MODE == MODE_V) mbar_wait(&y_full[ys], (j / kYS) & 1);
      mbar_wait(ps_full, j & 1);
      tc_fence_after();
      // MODE_V: B = y_j (dO_j); MODE_S: B = x_j
      const uint32_t sb = MODE == MODE_V ? yr0 + ys * (L::kSlot >> 4) : xr0 + xs * (L::kSlot >> 4);
      const uint64_t bh_ = mn_base | (uint64_t)sb, bl_ = mn_base | (uint64_t)(sb + (L::kTile >> 4));
#pragma unroll
      for (int ks = 0; ks < N / 16; ++ks) {
        if (leader) {
          umma_bf16_ts(tOut, tPS_hi + ks * 8, bh_ + ks * 64, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          if (out_lo) umma_bf16_ts(tOut, tPS_lo + ks * 8, bh_ + ks * 64, idesc_o, 1u);
          if (a.x3) umma_bf16_ts(tOut, tPS_hi + ks * 8, bl_ + ks * 64, idesc_o, 1u);
        }
      }
      if (leader) {
        umma_commit(ps_empty);
        if (MODE == MODE_V) umma_commit(&y_empty[ys]);
        else umma_commit(&x_empty[xs]);
        if (last) umma_commit(o_full);
      }
      __syncwarp();
    };

    mbar_wait(res_full, 0);
    tc_fence_after();
    issue_scores(0);
    for (int j = 0; j + 1 < nt; ++j) {
      issue_scores(j + 1);
      issue_out(j, false);
    }
    issue_out(nt - 1, true);
  } else {
    // ============================== SOFTMAX-BACKWARD / EPILOGUE ==============================
    const int quarter = warp & 3, slot = warp >> 2;
    const int row = quarter * 32 + lane;
    const int ri = r0 + row;  // key index (kv_roles) or query index
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int b = bh / a.H, h = bh % a.H;

    // ---- resident operands: plane rows -> TMEM (thread = row)
    {
      if constexpr (MODE == MODE_S) {  // three planes: slot 0: X hi, slot 1: X lo, slot 2: Y hi (all atoms each)
        if (slot == 0) plane_row_to_tmem<AD>(X_hi, bh, ri, a.Tp, tmem_base + lane_off + cX_hi, 0, 1);
        if (slot == 1 && a.x3) plane_row_to_tmem<AD>(X_lo, bh, ri, a.Tp, tmem_base + lane_off + cX_lo, 0, 1);
        if (slot == 2) plane_row_to_tmem<AD>(Y_hi, bh, ri, a.Tp, tmem_base + lane_off + cY_hi, 0, 1);
      } else {  // two planes: slot s takes plane (s & 1) and the head-dim atoms of parity (s >> 1)
        const bool lo = (slot & 1) != 0;
        if (!lo || a.x3)
          plane_row_to_tmem<AD>(lo ? X_lo : X_hi, bh, ri, a.Tp, tmem_base + lane_off + (lo ? cX_lo : cX_hi), slot >> 1, 2);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(res_full);
    }

    const float inv_keep = a.drop.p > 0.f ? 1.f / (1.f - a.drop.p) : 1.f;
    const float* lse = a.lse + (int64_t)bh * a.Tp;
    const float* delta = a.delta + (int64_t)bh * a.Tp;
    float lse_r = 0.f, delta_r = 0.f;
    if (!kv_roles) { lse_r = lse[ri]; delta_r = delta[ri]; }  // padded rows: lse = +inf, delta = 0

    for (int j = 0; j < nt; ++j) {
      const int buf = MODE == MODE_V ? (j & 1) : 0;
      // per-column softmax statistics of this tile, fetched before the scores arrive
      float lsv[kv_roles ? NH : 1], dlv[(kv_roles && MODE == MODE_S) ? NH : 1];
      if constexpr (kv_roles) {
        const float4* l4 = reinterpret_cast<const float4*>(lse + j * N + slot * NH);
        const float4* d4 = reinterpret_cast<const float4*>(delta + j * N + slot * NH);
#pragma unroll
        for (int g = 0; g < NH / 4; ++g) {
          const float4 lv = __ldg(l4 + g);
          lsv[4 * g] = lv.x; lsv[4 * g + 1] = lv.y; lsv[4 * g + 2] = lv.z; lsv[4 * g + 3] = lv.w;
          if constexpr (MODE == MODE_S) {
            const float4 dd = __ldg(d4 + g);
            dlv[4 * g] = dd.x; dlv[4 * g + 1] = dd.y; dlv[4 * g + 2] = dd.z; dlv[4 * g + 3] = dd.w;
          }
        }
      }
      // dropout multipliers of this tile from the keep bits the forward wrote (bit k of row q of the [Tp, Tp/8]-byte mask
      // of this head): independent of the scores, so they are fetched before the wait below.  Re-drawing the Philox
      // decisions here cost ~14 instructions per element in the key-major passes (a call covers 8 adjacent KEYS of one
      // query = 8 adjacent lanes, so the calls went through a shared-memory transpose) on warps that are instruction-bound.
      const int c0 = j * N + slot * NH;
      float mult[NH];
#pragma unroll
      for (int c = 0; c < NH; ++c) mult[c] = 1.f;
      if (a.drop.p > 0.f) {
        const uint8_t* bits = a.drop_bits + (int64_t)bh * a.Tp * (a.Tp >> 3);
        if constexpr (kv_roles) {
          // thread = key row, columns = queries: the 32 keys of this warp are one aligned 32-bit word of a query's mask row;
          // lane c fetches the word of query c0 + c, a shuffle hands it to everybody, lane l tests bit l
          const int kw = (r0 + quarter * 32) >> 5;
          const uint32_t mine = lane < NH ? __ldg(reinterpret_cast<const uint32_t*>(bits + (int64_t)(c0 + lane) * (a.Tp >> 3)) + kw) : 0u;
#pragma unroll
          for (int c = 0; c < NH; ++c) mult[c] = (__shfl_sync(0xffffffffu, mine, c) >> lane) & 1u ? inv_keep : 0.f;
        } else {  // thread = query row, columns = keys c0 .. c0 + NH - 1: NH / 8 bytes of this row's mask
#pragma unroll
          for (int g = 0; g < NH / 8; ++g) {
            const uint32_t kb = __ldg(bits + (int64_t)ri * (a.Tp >> 3) + ((c0 + 8 * g) >> 3));
#pragma unroll
            for (int e = 0; e < 8; ++e) mult[8 * g + e] = (kb >> e) & 1u ? inv_keep : 0.f;
          }
        }
      }
      mbar_wait(&sp_full[buf], MODE == MODE_V ? ((j >> 1) & 1) : (j & 1));
      tc_fence_after();
      float sv[NH], dv[NH];
      {
        uint32_t v1[NH];
        if constexpr (NH == 16) tmem_ld16(tmem_base + lane_off + cS + buf * N + slot * NH, v1);
        else tmem_ld8(tmem_base + lane_off + cS + buf * N + slot * NH, v1);
        if constexpr (MODE == MODE_S) {
          uint32_t v2[NH];
          if constexpr (NH == 16) tmem_ld16(tmem_base + lane_off + cDP + slot * NH, v2);
          else tmem_ld8(tmem_base + lane_off + cDP + slot * NH, v2);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < NH; ++c) dv[c] = __uint_as_float(v2[c]);
        } else {
          tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < NH; ++c) sv[c] = __uint_as_float(v1[c]);
      }
      tc_fence_before();
      mbar_arrive(&sp_empty[buf]);

      // No bounds checks: the forward stores lse = +inf for padded queries (so P' = exp2(-inf) = 0 there), delta
      // is zero-padded, and padded KEY rows only pollute accumulator rows that are never stored.
      if constexpr (kv_roles) {
#pragma unroll
        for (int c = 0; c < NH; ++c) {
          const float p = fast_exp2(fmaf(sv[c], a.scale_log2, -lsv[c]));
          if constexpr (MODE == MODE_S) sv[c] = p * (dv[c] * mult[c] - dlv[c]);  // dS'
          else sv[c] = p * mult[c];                                               // Pd'
        }
      } else {
#pragma unroll
        for (int c = 0; c < NH; ++c) {
          const float p = (c0 + c < a.T) ? fast_exp2(fmaf(sv[c], a.scale_log2, -lse_r)) : 0.f;
          sv[c] = p * (dv[c] * mult[c] - delta_r);  // dS'
        }
      }
      // A operand of the output MMA -> TMEM (two bf16 per column), hi and lo planes
      mbar_wait(ps_empty, (j & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < NH / 8; ++ch) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = sv[ch * 8 + e];
        uint32_t hw[4], lw[4];
        split8_words(x, hw, lw);
        const uint32_t col = (slot * NH + ch * 8) / 2;
        tmem_st4(tmem_base + lane_off + cPS_hi + col, hw[0], hw[1], hw[2], hw[3]);
        if (out_lo) tmem_st4(tmem_base + lane_off + cPS_lo + col, lw[0], lw[1], lw[2], lw[3]);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ps_full);
    }
    // ---- epilogue: accumulator -> d_qkv (fp32, packed [B,T,3*H*E]) and / or operand planes; each slot: AD*8 columns
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int I = a.H * a.E;
    // MODE_V: dV (col block 2I);  MODE_S/KV: dK (col block I, scaled);  MODE_S/!KV: dQ (col block 0, scaled)
    constexpr int kSec = MODE == MODE_V ? 2 : (KV ? 1 : 0);
    float* dst = a.dqkv ? a.dqkv + ((int64_t)b * a.T + ri) * (3 * I) + h * a.E + kSec * I : nullptr;
    const float sc = MODE == MODE_V ? 1.f : a.scale;
    const int64_t prow = (int64_t)b * a.T + ri;            // row of the [B*T, 3*H*Dp] gradient matrix
    const int atom0 = (kSec * a.H + h) * AD;                // first column atom of this head's slice
    const uint32_t stage = smem_u32(smem + L::x_ring);
#pragma unroll
    for (int cc = 0; cc < AD; ++cc) {
      const int d0 = slot * (AD * 8) + cc * 8;
      uint32_t v[8];
      tmem_ld8(tmem_base + lane_off + cOut + d0, v);
      tmem_ld_wait();
      if (ri < a.T) {
        if (dst) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (d0 + c < a.E) dst[d0 + c] = __uint_as_float(v[c]) * sc;
        }
      }
      if (a.dq_pl.hi) {
        // operand planes for the Wqkv weight-gradient and input-gradient GEMMs (pad columns are 0): staged in the idle
        // x ring in plane layout, then one bulk store per (head-dim atom, plane) writes 128 consecutive plane rows
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[e]) * sc;
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t so = (uint32_t)(d0 >> 5) * (128 * 64) + row * 64 + ((((d0 & 31) >> 3) ^ (int)((prow >> 1) & 3)) << 4);
        sts128(stage + so, hi);
        if (a.dq_pl.lo) sts128(stage + AD * 128 * 64 + so, lo);
      }
    }
    if (a.dq_pl.hi) {
      fence_proxy_async();
      named_bar_sync(1, kSmThreads);
      const int rows_valid = min(128, a.T - r0);  // rows past T belong to the next sample of the flat matrix
      if (threadIdx.x < AD * 2 && rows_valid > 0) {
        const int at_i = threadIdx.x >> 1, pln = threadIdx.x & 1;
        uint8_t* dstp = pln ? a.dq_pl.lo : a.dq_pl.hi;
        if (dstp) {
          const int64_t off = ((int64_t)(atom0 + at_i) * a.dq_pl.rows_p + ((int64_t)b * a.T + r0)) * 64;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstp + off),
                       "r"(stage + (uint32_t)(pln * AD + at_i) * (128 * 64)), "r"(rows_valid * 64)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

// =====================================================================================================================
// dV + dK from ONE recomputation of the scores: a CLUSTER OF TWO CTAs per 128-key tile (V1T_ATTN_BWD=pair).
//
// dV and dK of a key tile both need P' of every (key, query) pair, but the tensor memory of one SM cannot hold both
// accumulators (2 x 160 columns) next to the resident operands, the scores and the A operand of the accumulating MMA.
// Two CTAs of a cluster split the work instead and exchange P' through distributed shared memory:
//
//   rank 0 ("P side")   K in TMEM; per 64-query tile j:  S'^T = K Q_j^T,  P' = exp2(S' c - lse)  --P' (32 KB)-->  rank 1
//                       Pd' = P' * dropout  ->  dV += Pd'^T dO_j
//   rank 1 ("dS side")  V in TMEM; per tile j:  dP'^T = V dO_j^T,  dS' = P' (dP' * dropout - delta)  ->  dK += dS'^T Q_j
//
// Both ranks stream the same (Q_j, dO_j) tiles with 64-column MMAs (the fused dK pass of the three-pass kernel only has
// TMEM for 32-column tiles) and run the same MMA / bulk-copy code: ring 0 feeds the score MMA (rank 0: Q_j, rank 1:
// dO_j), ring 1 the accumulating MMA (rank 0: dO_j, rank 1: Q_j).  P' crosses as fp32 with the dropout decision in the
// SIGN bit (P' >= 0), so rank 1 neither draws nor transposes the mask.  MMA units per (key tile, query tile):
// 3 + 3 (rank 0) and 3 + 3 (rank 1) instead of 6 (dV pass) + 9 (dK pass); dQ stays the query-stationary pass above.
// TMEM columns (both ranks): R_hi | R_lo | scores[2][64] | A_hi[32] | A_lo[32] | out[Dp] = 2 HC + 192 + Dp (512 @ Dp = 160).
// The clusters are PERSISTENT (one per SM pair): each loops over (head, key tile) items with tensor memory, barriers and the
// rings kept alive; ring 0 prefetches the next item's first tiles during the current item's tail, ring 1 carries the next
// resident operand as soon as its last accumulating MMA has completed (hand-over measured by scripts/pair_trace_model.py).
// =====================================================================================================================
// optional per-tile timeline of one cluster (diagnostics: include/v1t_b200_diag.h, scripts/pair_trace.py); null = off
__device__ long long* g_pair_trace = nullptr;
// [item 2][rank 2][tile][event] cycles since the cluster's start barrier, for the first two items of cluster 0; row
// kTraceTiles - 1 holds the hand-over between items: 0 accumulator complete, 1 epilogue done, 2 resident operand stored,
// 3 (MMA warp) resident operand visible, 4 first accumulator chunk read
constexpr int kTraceTiles = 32, kTraceEvents = 8;
#define PAIR_TRACE(ev, j)                                                                                       \
  do {                                                                                                          \
    if (trace && (j) < kTraceTiles - 1) trace[((int)rank * kTraceTiles + (j)) * kTraceEvents + (ev)] = clock64() - t_start; \
  } while (0)
// hand-over events of an item (row kTraceTiles - 1)
#define PAIR_TRACE_X(ev)                                                                                         \
  do {                                                                                                           \
    if (trace) trace[((int)rank * kTraceTiles + kTraceTiles - 1) * kTraceEvents + (ev)] = clock64() - t_start; \
  } while (0)

template <int AD>
struct SmemPair {
  static constexpr int N = 64;
  static constexpr uint32_t kTile = AD * N * 64;   // one bf16 plane of a streamed 64-row tile
  static constexpr uint32_t kSlot = 2 * kTile;     // hi + lo
  static constexpr uint32_t ring0 = 0;             // 2 slots: operand of the score MMA
  static constexpr uint32_t ring1 = 2 * kSlot;     // 2 slots: operand of the accumulating MMA
  static constexpr uint32_t xch = 4 * kSlot;       // rank 1: P' from rank 0, [2 buffers][4 slots][4 chunks][128 rows][16 B]
  static constexpr uint32_t kXch = 2 * 4 * 4 * 128 * 16;
  static constexpr uint32_t bars = xch + kXch;
  static constexpr uint32_t total = bars + 512 + 1024;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// asynchronous 16-byte store into a peer CTA's shared memory that completes 16 transaction bytes on an mbarrier of
// THAT CTA (STAS): the receiver's wait on the barrier's phase orders the data, no fence on the sender
__device__ __forceinline__ void st_async_v4(uint32_t addr, float a, float b, float c, float d, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)), "r"(mbar)
               : "memory");
}
// "buffer consumed" signal to the peer: no data travels with it, so no release fence (a release.cluster arrive by
// every thread cost a MEMBAR each: stall_membar was the top stall reason of the first version, 870 us)
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// one 256-bit global store (sm_100: STG.256), 32-byte aligned; L2 evict-first: the dS' planes are written once and read
// by the next launch, they should not push the Q / dO planes that 13 clusters per head re-read out of L2
__device__ __forceinline__ void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
template <int AD>
__global__ void __launch_bounds__(kThreadsAttn, 1) attn_bwd_pair_kernel(const AttnBwdArgs a, const int tiles) {
  constexpr int N = 64, Dp = AD * 32, HC = AD * 16, NH = N / kSlots, OPC = N / 2;
  using L = SmemPair<AD>;
  extern __shared__ uint8_t smem_raw[];
  // the dynamic shared window starts at the same offset in both CTAs of the cluster, so this rounding agrees too
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::bars);
  uint64_t* res_full = bars + 0;   // resident operand rows stored to TMEM by the softmax threads
  uint64_t* ps_full = bars + 1;    // A operand of the accumulating MMA written
  uint64_t* ps_empty = bars + 2;
  uint64_t* o_full = bars + 3;
  uint64_t* r0_full = bars + 4;    // [2] ring 0
  uint64_t* r0_empty = bars + 6;
  uint64_t* r1_full = bars + 8;    // [2] ring 1
  uint64_t* r1_empty = bars + 10;
  uint64_t* sp_full = bars + 12;   // [2] score buffers
  uint64_t* sp_empty = bars + 14;
  uint64_t* pe_full = bars + 16;   // [2] rank 1: P' of a tile has landed (arming arrive + 32 KB of st.async bytes)
  uint64_t* pe_empty = bars + 18;  // [2] rank 0: rank 1 has consumed the buffer (one remote arrive per warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  uint64_t* stage_full = bars + 22;  // the resident operand's 128-row tile has landed in ring 1 (bulk copy)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0: P side (dV), 1: dS side (dK)
  const int nt = (a.T + N - 1) / N;
  // PERSISTENT clusters: cluster c works on items c, c + #clusters, ... (item = head * tiles + 128-key tile, so the clusters
  // running at one time share a few heads' planes in L2).  Tensor memory, barriers and the bulk-copy rings live across
  // items: every ring / score / exchange barrier is indexed by the running tile count `it` = k * nt + j, the per-item
  // barriers (stage_full, res_full, o_full) by the item count k.
  const int ncl = (int)gridDim.x >> 1, cid = (int)blockIdx.x >> 1, total = tiles * a.B * a.H;
  const int my_items = cid < total ? (total - cid + ncl - 1) / ncl : 0;
  const int total_its = my_items * nt;
  const uint8_t* R_hi = rank == 0 ? a.k_hi : a.v_hi;  const uint8_t* R_lo = rank == 0 ? a.k_lo : a.v_lo;   // resident
  const uint8_t* s0_hi = rank == 0 ? a.q_hi : a.do_hi; const uint8_t* s0_lo = rank == 0 ? a.q_lo : a.do_lo; // ring 0
  const uint8_t* s1_hi = rank == 0 ? a.do_hi : a.q_hi; const uint8_t* s1_lo = rank == 0 ? a.do_lo : a.q_lo; // ring 1

  if (threadIdx.x == 0) {
    mbar_init(res_full, kSmThreads);
    mbar_init(stage_full, 1);
    mbar_init(ps_full, kSmThreads);
    mbar_init(ps_empty, 1);
    mbar_init(o_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&r0_full[i], 1);
      mbar_init(&r0_empty[i], 1);
      mbar_init(&r1_full[i], 1);
      mbar_init(&r1_empty[i], 1);
      mbar_init(&sp_full[i], 1);
      mbar_init(&sp_empty[i], kSmThreads);
      mbar_init(&pe_full[i], 1);          // one arming arrive (expect_tx) + 32 KB of st.async transaction bytes
      mbar_init(&pe_empty[i], kSmWarps);  // one remote arrive per softmax warp of rank 1
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  const long long t_start = clock64();
  long long* const trace0 = (g_pair_trace && blockIdx.x < 2 && (threadIdx.x & 31) == 0) ? g_pair_trace : nullptr;
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t cR_hi = 0, cR_lo = HC, cS = 2 * HC, cPS_hi = cS + 2 * N, cPS_lo = cPS_hi + OPC, cOut = cPS_lo + OPC;
  static_assert(cOut + Dp <= 512, "TMEM budget exceeded");

  if (warp == kLoadWarpX || warp == kLoadWarpY) {
    // ============================== BULK-COPY PRODUCERS (one warp per ring) ==============================
    if (lane == 0) {
      const bool is0 = warp == kLoadWarpX;
      uint64_t* fullb = is0 ? r0_full : r1_full;
      uint64_t* emptyb = is0 ? r0_empty : r1_empty;
      const uint8_t* src_hi = is0 ? s0_hi : s1_hi;
      const uint8_t* src_lo = is0 ? s0_lo : s1_lo;
      uint8_t* ring = smem + (is0 ? L::ring0 : L::ring1);
      for (int k = 0; k < my_items; ++k) {
        const int item = cid + k * ncl, bh = item / tiles, r0 = (item % tiles) * 128, itb = k * nt;
        if (!is0) {
          // ring 1 is idle between the last accumulating MMA of an item and the first one of the next: it first carries
          // the resident operand's 128 rows (two consecutive 64-row plane tiles = ONE contiguous copy per plane) to the
          // softmax warps, which move them to TMEM
          if (k > 0) {  // both slots drained: the producer waits of tiles itb and itb + 1, without producing
            mbar_wait(&emptyb[itb & 1], ((itb >> 1) & 1) ^ 1);
            mbar_wait(&emptyb[(itb + 1) & 1], (((itb + 1) >> 1) & 1) ^ 1);
          }
          const int64_t roff = attn_plane_off(bh, 0, r0, a.Tp, AD);
          mbar_expect_tx(stage_full, (a.x3 ? 2 : 1) * 2 * L::kTile);
          bulk_g2s(ring, R_hi + roff, 2 * L::kTile, stage_full);
          if (a.x3) bulk_g2s(ring + 2 * L::kTile, R_lo + roff, 2 * L::kTile, stage_full);
          mbar_wait(res_full, k & 1);  // every row has been read out of the ring
        }
        for (int j = 0; j < nt; ++j) {
          const int it = itb + j, s = it & 1;
          mbar_wait(&emptyb[s], ((it >> 1) & 1) ^ 1);
          uint8_t* base = ring + s * L::kSlot;
          mbar_expect_tx(&fullb[s], (a.x3 ? 2 : 1) * L::kTile);
          const int64_t src = attn_plane_off(bh, 0, j * N, a.Tp, AD);  // a 64-row plane tile is contiguous: one copy per plane
          bulk_g2s(base, src_hi + src, L::kTile, &fullb[s]);
          if (a.x3) bulk_g2s(base + L::kTile, src_lo + src, L::kTile, &fullb[s]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================== MMA ISSUER (identical for both ranks) ==============================
    const bool leader = elect_one();
    const uint32_t idesc_s = idesc_bf16(128, N, 0, 0);
    const uint32_t idesc_o = idesc_bf16(128, Dp, 0, 1);  // B = streamed tile viewed MN-major (head dim contiguous)
    constexpr uint32_t tb = N * 64;
    // descriptor start addresses are 14-bit fields: in a cluster launch the shared-window address of a CTA with a
    // non-zero rank carries the rank in its upper bits, which would otherwise spill into the LBO field of the MN-major
    // descriptors below (rank 1 then read every head-dim atom but the first from the wrong place)
    const uint32_t g0 = (smem_u32(smem + L::ring0) >> 4) & 0x3FFFu, g1 = (smem_u32(smem + L::ring1) >> 4) & 0x3FFFu;
    const uint64_t mn_base = desc_mn_sw64_base(tb);
    const uint32_t tR_hi = tmem_base + cR_hi, tR_lo = tmem_base + cR_lo;
    const uint32_t tPS_hi = tmem_base + cPS_hi, tPS_lo = tmem_base + cPS_lo, tOut = tmem_base + cOut;

    long long* trace = nullptr;
    auto issue_scores = [&](int it, int j) {
      const int s = it & 1;
      const uint32_t xb = g0 + s * (L::kSlot >> 4);
      PAIR_TRACE(0, j);
      mbar_wait(&r0_full[s], (it >> 1) & 1);
      mbar_wait(&sp_empty[s], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      PAIR_TRACE(1, j);
      const uint64_t xh = kDescK64 | (uint64_t)xb, xl = kDescK64 | (uint64_t)(xb + (L::kTile >> 4));
      const uint32_t dS = tmem_base + cS + s * N;
#pragma unroll
      for (int ks = 0; ks < 2 * AD; ++ks) {
        const uint32_t bo = (ks >> 1) * (tb / 16) + (ks & 1) * 2;
        if (leader) {
          umma_bf16_ts(dS, tR_hi + ks * 8, xh + bo, idesc_s, ks > 0 ? 1u : 0u);
          if (a.x3) {
            umma_bf16_ts(dS, tR_lo + ks * 8, xh + bo, idesc_s, 1u);
            umma_bf16_ts(dS, tR_hi + ks * 8, xl + bo, idesc_s, 1u);
          }
        }
      }
      if (leader) {
        umma_commit(&r0_empty[s]);
        umma_commit(&sp_full[s]);
      }
      __syncwarp();
    };
    auto issue_out = [&](int it, int j, bool last) {
      const int s = it & 1;
      PAIR_TRACE(2, j);
      mbar_wait(&r1_full[s], (it >> 1) & 1);
      mbar_wait(ps_full, it & 1);
      tc_fence_after();
      PAIR_TRACE(3, j);
      const uint32_t sb = g1 + s * (L::kSlot >> 4);
      const uint64_t bh_ = mn_base | (uint64_t)sb, bl_ = mn_base | (uint64_t)(sb + (L::kTile >> 4));
#pragma unroll
      for (int ks = 0; ks < N / 16; ++ks) {
        if (leader) {
          umma_bf16_ts(tOut, tPS_hi + ks * 8, bh_ + ks * 64, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          if (a.x3) {
            umma_bf16_ts(tOut, tPS_lo + ks * 8, bh_ + ks * 64, idesc_o, 1u);
            umma_bf16_ts(tOut, tPS_hi + ks * 8, bl_ + ks * 64, idesc_o, 1u);
          }
        }
      }
      if (leader) {
        umma_commit(ps_empty);
        umma_commit(&r1_empty[s]);
        if (last) umma_commit(o_full);
      }
      __syncwarp();
    };

    for (int k = 0; k < my_items; ++k) {
      const int itb = k * nt;
      trace = (trace0 && k < 2) ? trace0 + k * (2 * kTraceTiles * kTraceEvents) : nullptr;
      // the resident operand of item k is in TMEM -- stored by the softmax threads after they drained `out` of item k - 1
      mbar_wait(res_full, k & 1);
      tc_fence_after();
      PAIR_TRACE_X(3);
      issue_scores(itb, 0);
      for (int j = 0; j + 1 < nt; ++j) {
        issue_scores(itb + j + 1, j + 1);
        issue_out(itb + j, j, false);
      }
      issue_out(itb + nt - 1, nt - 1, true);
    }
  } else {
    // ============================== SOFTMAX-BACKWARD / EPILOGUE ==============================
    const int quarter = warp & 3, slot = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    long long* trace = nullptr;

    // resident operand of item k: staged rows (ring 1) -> TMEM; slot s takes plane (s & 1) and the head-dim atoms of parity
    // (s >> 1).  Runs after the accumulator of item k - 1 has been read out, so res_full also tells the MMA warp that
    // `out` may be overwritten.
    auto load_resident = [&](int k) {
      const bool lo = (slot & 1) != 0;
      mbar_wait(stage_full, k & 1);
      if (!lo || a.x3)
        smem_row_to_tmem<AD>(smem_u32(smem + L::ring1) + (lo ? 2 * L::kTile : 0), row, tmem_base + lane_off + (lo ? cR_lo : cR_hi),
                             slot >> 1, 2);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(res_full);
      if (warp == 0) PAIR_TRACE_X(2);
    };

    const float inv_keep = a.drop.p > 0.f ? 1.f / (1.f - a.drop.p) : 1.f;
    const bool staged = a.dq_pl.hi != nullptr;  // the gradients leave as operand planes of the head-padded dqkv
    // exchange buffer addressing: chunk c (4 floats) of this thread's 16 columns sits at
    //   xch + ((buf * 4 + slot) * 4 + c) * 2048 + row * 16      (8 adjacent lanes = 128 contiguous bytes)
    const uint32_t xch_local = smem_u32(smem + L::xch) + (uint32_t)slot * 8192u + (uint32_t)row * 16u;
    const uint32_t xch_peer = mapa_u32(xch_local, 1u);                          // used by rank 0
    const uint32_t pe_full_peer = mapa_u32(smem_u32(pe_full), 1u);              // rank 0 -> rank 1
    const uint32_t pe_empty_peer = mapa_u32(smem_u32(pe_empty), 0u);            // rank 1 -> rank 0
    constexpr uint32_t kTileBytes = 128 * N * 4;                                // P' of one tile
    if (rank == 1 && threadIdx.x == 0) {  // arm the two exchange buffers for tiles 0 and 1
      if (total_its > 0) mbar_expect_tx(&pe_full[0], kTileBytes);
      if (total_its > 1) mbar_expect_tx(&pe_full[1], kTileBytes);
    }

    for (int k = 0; k < my_items; ++k) {
    const int item = cid + k * ncl, bh = item / tiles, r0 = (item % tiles) * 128, itb = k * nt;
    const int ri = r0 + row;  // key index
    const int b = bh / a.H, h = bh % a.H;
    const float* lse = a.lse + (int64_t)bh * a.Tp;
    const float* delta = a.delta + (int64_t)bh * a.Tp;
    trace = (trace0 && k < 2) ? trace0 + k * (2 * kTraceTiles * kTraceEvents) : nullptr;
    load_resident(k);
    for (int j = 0; j < nt; ++j) {
      const int it = itb + j, buf = it & 1;
      const int c0 = j * N + slot * NH;
      float stat[NH];  // rank 0: lse of this warp's query columns; rank 1: their delta
      {
        const float4* s4 = reinterpret_cast<const float4*>((rank == 0 ? lse : delta) + c0);
#pragma unroll
        for (int g = 0; g < NH / 4; ++g) {
          const float4 v = __ldg(s4 + g);
          stat[4 * g] = v.x; stat[4 * g + 1] = v.y; stat[4 * g + 2] = v.z; stat[4 * g + 3] = v.w;
        }
      }
      float sv[NH];
      if (rank == 0) {
        // dropout multipliers from the forward's keep bits (see attn_bwd2_body): lane c fetches the 32-key word of query
        // c0 + c, lane l tests bit l
        float mult[NH];
#pragma unroll
        for (int c = 0; c < NH; ++c) mult[c] = 1.f;
        if (a.drop.p > 0.f) {
          const uint8_t* bits = a.drop_bits + (int64_t)bh * a.Tp * (a.Tp >> 3);
          const int kw = (r0 + quarter * 32) >> 5;
          const uint32_t mine = lane < NH ? __ldg(reinterpret_cast<const uint32_t*>(bits + (int64_t)(c0 + lane) * (a.Tp >> 3)) + kw) : 0u;
#pragma unroll
          for (int c = 0; c < NH; ++c) mult[c] = (__shfl_sync(0xffffffffu, mine, c) >> lane) & 1u ? inv_keep : 0.f;
        }
        mbar_wait(&sp_full[buf], (it >> 1) & 1);
        tc_fence_after();
        if (warp == 0) PAIR_TRACE(4, j);
        {
          uint32_t v1[NH];
          tmem_ld16(tmem_base + lane_off + cS + buf * N + slot * NH, v1);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < NH; ++c) sv[c] = __uint_as_float(v1[c]);
        }
        tc_fence_before();
        mbar_arrive(&sp_empty[buf]);
#pragma unroll
        for (int c = 0; c < NH; ++c) sv[c] = fast_exp2(fmaf(sv[c], a.scale_log2, -stat[c]));  // P' (0 for padded queries)
        // ---- ship P' to rank 1: fp32, dropped elements carry the sign bit
        mbar_wait(&pe_empty[buf], ((it >> 1) & 1) ^ 1);  // rank 1 has consumed tile it - 2 (two tiles of slack)
        if (warp == 0) PAIR_TRACE(5, j);
#pragma unroll
        for (int c = 0; c < NH / 4; ++c) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = mult[4 * c + e] == 0.f ? __uint_as_float(__float_as_uint(sv[4 * c + e]) | 0x80000000u) : sv[4 * c + e];
          st_async_v4(xch_peer + (uint32_t)buf * 32768u + (uint32_t)c * 2048u, o[0], o[1], o[2], o[3],
                      pe_full_peer + (uint32_t)buf * 8u);
        }
#pragma unroll
        for (int c = 0; c < NH; ++c) sv[c] *= mult[c];  // Pd'
      } else {
        mbar_wait(&sp_full[buf], (it >> 1) & 1);
        tc_fence_after();
        if (warp == 0) PAIR_TRACE(4, j);
        float dv[NH];
        {
          uint32_t v1[NH];
          tmem_ld16(tmem_base + lane_off + cS + buf * N + slot * NH, v1);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < NH; ++c) dv[c] = __uint_as_float(v1[c]);
        }
        tc_fence_before();
        mbar_arrive(&sp_empty[buf]);
        // ---- P' from rank 0
        mbar_wait(&pe_full[buf], (it >> 1) & 1);  // all 32 KB of tile it have landed (st.async transaction bytes)
        if (warp == 0) PAIR_TRACE(5, j);
#pragma unroll
        for (int c = 0; c < NH / 4; ++c) {
          const uint4 v = lds128(xch_local + (uint32_t)buf * 32768u + (uint32_t)c * 2048u);
          sv[4 * c] = __uint_as_float(v.x); sv[4 * c + 1] = __uint_as_float(v.y);
          sv[4 * c + 2] = __uint_as_float(v.z); sv[4 * c + 3] = __uint_as_float(v.w);
        }
        // re-arm this buffer for tile j + 2: its bytes cannot start to arrive before every warp of this CTA has signalled
        // pe_empty below, i.e. not before all of them have passed the wait above
        if (threadIdx.x == 0 && it + 2 < total_its) mbar_expect_tx(&pe_full[buf], kTileBytes);
#pragma unroll
        for (int c = 0; c < NH; ++c) {
          const uint32_t u = __float_as_uint(sv[c]);
          const float p = __uint_as_float(u & 0x7fffffffu);
          const float m = (u & 0x80000000u) ? 0.f : inv_keep;
          sv[c] = p * (dv[c] * m - stat[c]);  // dS'
        }
      }
      // rank 1 also leaves dS' as GEMM operand planes for the dQ GEMM (bwd2_all): this thread's 16 queries of key row `ri`
      // are half a 64-byte plane row, i.e. one full 32-byte sector per plane
      uint32_t ds_h[NH / 8][4], ds_l[NH / 8][4];
      // layout [bh][128-query tile][64-key tile][4 query atoms][64 key rows][64 B]: a (128 queries x 64 keys) block -- one
      // k-block of the GEMM's A operand -- is 16 contiguous KB
      const int qa = j * 2 + (slot >> 1);  // 32-query atom of this thread's columns
      uint8_t* const ds_row =
          (rank == 1 && a.ds_hi)
              ? a.ds_hi + ((((((int64_t)bh * (a.Tp >> 7) + (qa >> 2)) * (a.Tp >> 6) + (ri >> 6)) * 4 + (qa & 3)) * 64 + (ri & 63)) << 6)
              : nullptr;
      // A operand of the accumulating MMA -> TMEM (two bf16 per column), hi and lo planes
      mbar_wait(ps_empty, (it & 1) ^ 1);
      tc_fence_after();
      if (warp == 0) PAIR_TRACE(6, j);
#pragma unroll
      for (int ch = 0; ch < NH / 8; ++ch) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = sv[ch * 8 + e];
        uint32_t hw[4], lw[4];
        split8_words(x, hw, lw);
        const uint32_t col = (slot * NH + ch * 8) / 2;
        tmem_st4(tmem_base + lane_off + cPS_hi + col, hw[0], hw[1], hw[2], hw[3]);
        if (a.x3) tmem_st4(tmem_base + lane_off + cPS_lo + col, lw[0], lw[1], lw[2], lw[3]);
        if (ds_row) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ds_h[ch][i] = hw[i];
            ds_l[ch][i] = lw[i];
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ps_full);
      if (warp == 0) PAIR_TRACE(7, j);
      if (rank == 1) {  // P' of tile j is consumed (its values went through the dS' -> TMEM chain above): one signal per warp
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_relaxed(pe_empty_peer + (uint32_t)buf * 8u);
      }
      // (after the arrivals above: the accumulating MMA and rank 0 do not wait for these stores)
      if (ds_row) {
        // the two 16-byte chunks of this thread are the two halves of one aligned 32-byte sector of the swizzled plane row
        // (their order swaps with the row's swizzle bit 0): ONE 256-bit store per plane -- 16-byte stores at a 64-byte
        // stride cost an LSU pass per half sector and slowed the kernel by a third
        const int sw = (ri >> 1) & 3;
        const int64_t co = (int64_t)(((((slot & 1) * 2) ^ sw) & ~1) << 4);
        const bool swap = (sw & 1) != 0;
        uint32_t f[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f[i] = swap ? ds_h[1][i] : ds_h[0][i];
          f[4 + i] = swap ? ds_h[0][i] : ds_h[1][i];
        }
        stg256(ds_row + co, f);
        if (a.ds_lo) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f[i] = swap ? ds_l[1][i] : ds_l[0][i];
            f[4 + i] = swap ? ds_l[0][i] : ds_l[1][i];
          }
          stg256(a.ds_lo + (ds_row - a.ds_hi) + co, f);
        }
      }
    }
    // ---- epilogue: rank 0 -> dV (column block 2I), rank 1 -> dK (column block I, scaled); see attn_bwd2_body
    mbar_wait(o_full, k & 1);
    tc_fence_after();
    if (warp == 0) PAIR_TRACE_X(0);
    const int I = a.H * a.E;
    const int kSec = rank == 0 ? 2 : 1;
    float* dst = a.dqkv ? a.dqkv + ((int64_t)b * a.T + ri) * (3 * I) + h * a.E + kSec * I : nullptr;
    const float sc = rank == 0 ? 1.f : a.scale;
    const int64_t prow = (int64_t)b * a.T + ri;
    const int atom0 = (kSec * a.H + h) * AD;
    // Operand planes straight from the registers: one 16-byte chunk per (row, 8 columns) and plane, merged into full
    // 64-byte plane rows in L2.  The epilogue is bound by getting 80 KB per item out of the SM while every other SM streams
    // its tiles from L2, not by how the stores are formed -- measured per item with every softmax warp held (cycles from
    // "accumulator complete" to the end of the epilogue, scripts/pair_trace_model.py): shared-memory stage + bulk stores
    // with their fence / wait rounds 5.6 k, stage + 512-byte coalesced copy-out by all warps 5.2 k, these stores 5.8 k
    // (but no barrier, no stage, and the shortest kernel of the three).
    uint8_t* const phi = staged && ri < a.T ? a.dq_pl.hi : nullptr;  // rows past T belong to the next sample
    uint8_t* const plo = staged && ri < a.T ? a.dq_pl.lo : nullptr;
#pragma unroll
    for (int cc = 0; cc < AD; ++cc) {
      const int d0 = slot * (AD * 8) + cc * 8;
      uint32_t v[8];
      tmem_ld8(tmem_base + lane_off + cOut + d0, v);
      tmem_ld_wait();
      if (cc == 0 && warp == 0) PAIR_TRACE_X(4);
      if (ri < a.T && dst) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (d0 + c < a.E) dst[d0 + c] = __uint_as_float(v[c]) * sc;
      }
      if (phi) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[e]) * sc;
        uint4 hi, lo;
        split8(x, hi, lo);
        const int64_t off = ((int64_t)(atom0 + (d0 >> 5)) * a.dq_pl.rows_p + prow) * 64 + ((((d0 & 31) >> 3) ^ (int)((prow >> 1) & 3)) << 4);
        *reinterpret_cast<uint4*>(phi + off) = hi;
        if (plo) *reinterpret_cast<uint4*>(plo + off) = lo;
      }
    }
    tc_fence_before();
    if (warp == 0) PAIR_TRACE_X(1);
    }  // items
  }

  __syncthreads();
  cluster_sync_all();  // neither CTA leaves while its peer may still store into its shared memory or barriers
  if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

// The three passes (dK, dQ, dV) are independent: ONE 1-D grid runs them, so that the 5.6-wave tails of three separate
// launches (832 CTAs on 148 SMs each) become one 16.9-wave launch.  Block order: heads are taken in GROUPS of `group`
// (b, h) pairs; within a group all dK CTAs come first, then all dQ, then all dV.
//   * group = B*H (pass outermost) re-streams every head's Q / K / V / dO planes from HBM once per pass: 753 MB of
//     DRAM reads per launch at B = 16 against 271 MB algorithmic;
//   * group = 1 ((b, h) outermost) reads 275 MB but measured 3.6 % SLOWER (1209 vs 1167 us under ncu): the SMs then run
//     a mix of the three template bodies (8.2 k SASS instructions together) instead of one at a time;
//   * a group of 16 heads keeps the group's planes (69 MB) in the 126 MB L2 while ~all resident CTAs run the same body.
// The dV CTAs (half as long as the others) come last within a group, so the launch also ends on short CTAs.
template <int AD>
__global__ void __launch_bounds__(kThreadsAttn, 1) attn_bwd2_kernel(const AttnBwdArgs a, const int only_dq, const int tiles,
                                                                    const int group) {
  if (only_dq) {  // grid (tiles, 1, B*H): the query-stationary pass alone (the pair kernel has done dV and dK)
    attn_bwd2_body<MODE_S, false, 32, AD>(a, blockIdx.x, blockIdx.z);
    return;
  }
  const int BH = a.B * a.H;
  const int per_full = 3 * tiles * group;
  const int g = blockIdx.x / per_full;
  const int gsz = min(group, BH - g * group);
  int rem = blockIdx.x - g * per_full;
  const int pass = rem / (tiles * gsz);
  rem -= pass * tiles * gsz;
  const int bh = g * group + rem / tiles, tile = rem % tiles;
  if (pass == 0) attn_bwd2_body<MODE_S, true, 32, AD>(a, tile, bh);        // dK
  else if (pass == 1) attn_bwd2_body<MODE_S, false, 32, AD>(a, tile, bh);  // dQ
  else attn_bwd2_body<MODE_V, true, 64, AD>(a, tile, bh);                  // dV
}

template <int AD>
int bwd2_all(const AttnBwdArgs& a_in, cudaStream_t st) {
  AttnBwdArgs a = a_in;
  const bool pair = attn_bwd_pair_env() != 0;
  if (!pair || !attn_dq_gemm_env()) a.ds_hi = a.ds_lo = nullptr;  // no dS' planes: the dQ pass recomputes what it needs
  constexpr uint32_t smem = std::max({Smem2<MODE_V, 64, AD>::total, Smem2<MODE_S, 32, AD>::total});
  static_assert(smem <= 232448, "shared memory budget exceeded");
  V1T_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel<AD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (pair) {
    // dV + dK by clusters of two CTAs that share one recomputation of P' (see attn_bwd_pair_kernel), then dQ alone
    using LP = SmemPair<AD>;
    static_assert(LP::total <= 232448, "shared memory budget exceeded");
    V1T_CUDA(cudaFuncSetAttribute(attn_bwd_pair_kernel<AD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LP::total));
    cudaLaunchConfig_t cfg = {};
    const int tiles_p = cdiv(a.T, 128);
    const int64_t items = (int64_t)tiles_p * a.B * a.H;
    V1T_CHECK_ARG(items <= (1 << 30), "attn_bwd2_tc: too many (head, key tile) items");
    // persistent clusters, one per SM pair (V1T_ATTN_PAIR_PERSIST=0: one cluster per item)
    static const bool persist = [] { const char* e = getenv("V1T_ATTN_PAIR_PERSIST"); return !(e && e[0] == '0'); }();
    cfg.gridDim = dim3(2 * (unsigned)(persist ? std::min<int64_t>(items, kNumSMs / 2) : items), 1, 1);
    cfg.blockDim = dim3(kThreadsAttn, 1, 1);
    cfg.dynamicSmemBytes = LP::total;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    {
      ProfScope prof(V1T_PHASE_ATTN_BWD_PAIR, st);
      V1T_CUDA(cudaLaunchKernelEx(&cfg, attn_bwd_pair_kernel<AD>, a, tiles_p));
      V1T_LAUNCH_CHECK();
    }
    ProfScope prof(V1T_PHASE_ATTN_BWD_DQ, st);
    if (a.ds_hi) {
      // dQ = scale * dS K as ONE batched tcgen05 GEMM over the dS' planes the pair kernel has just written (A, MN-major:
      // the contraction runs over the plane rows = keys) and the K attention planes (B, MN-major, tile-major layout):
      // 3 MMA units per (query tile, key tile) instead of the 9 of the query-stationary pass, which recomputes S and dP
      v1t_gemm_desc g{};
      g.m = a.T; g.k = a.T; g.n = a.dq_pl.hi ? AD * 32 : a.E;
      g.batch1 = a.B; g.batch2 = a.H; g.alpha = a.scale;
      g.a_m = 1; g.a_k = a.Tp; g.b_n = 1; g.b_k = a.Tp;
      const PlaneOp pa{a.ds_hi, a.x3 ? a.ds_lo : nullptr, a.Tp, a.Tp / 32, (int64_t)(a.Tp / 32) * a.Tp * 64, 2};
      const PlaneOp pb{a.k_hi, a.x3 ? a.k_lo : nullptr, a.Tp, AD, (int64_t)a.Tp * AD * 64, 1};
      EpiOp epi = no_epi();
      float* C = nullptr;
      if (a.dq_pl.hi) {  // dQ of (sample b, head h): rows b * T + t, column atoms h * AD + d / 32 of the gradient planes
        epi.kind = kEpiPlanesOut;
        epi.pl = PlaneOut{a.dq_pl.hi, a.x3 ? a.dq_pl.lo : nullptr, a.dq_pl.rows_p, a.T, AD};
      } else {
        const int64_t I3 = 3ll * a.H * a.E;
        C = a.dqkv;
        g.c_m = I3; g.c_b1 = (int64_t)a.T * I3; g.c_b2 = a.E;
      }
      return gemm_tc(g, nullptr, nullptr, C, nullptr, nullptr, st, no_drop(), a.x3, epi, pa, pb);
    }
    dim3 gq(cdiv(a.T, 128), 1, a.B * a.H);
    attn_bwd2_kernel<AD><<<gq, kThreadsAttn, smem, st>>>(a, 1, cdiv(a.T, 128), 1);
    V1T_LAUNCH_CHECK();
    return V1T_OK;
  }
  const int tiles = cdiv(a.T, 128), BH = a.B * a.H;
  const int group = std::max(1, std::min(attn_bwd_group_env(), BH));
  V1T_CHECK_ARG((int64_t)3 * tiles * BH <= 2147483647ll, "attn_bwd2_tc: grid too large");
  attn_bwd2_kernel<AD><<<dim3(3 * tiles * BH), kThreadsAttn, smem, st>>>(a, 0, tiles, group);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

}  // namespace

// diagnostics hook (include/v1t_b200_diag.h): per-tile timeline buffer of the pair kernel, 2 * 2 * 32 * 8 int64, or null
int attn_pair_trace_set(long long* buf) {
  V1T_CUDA(cudaMemcpyToSymbol(g_pair_trace, &buf, sizeof(buf)));
  return V1T_OK;
}

int attn_bwd2_tc(const AttnBwdArgs& a, cudaStream_t st) {
  V1T_CHECK_ARG(a.Dp % 32 == 0 && a.Dp >= 32 && a.Dp <= 160 && a.Tp % 128 == 0 && a.Tp >= a.T && a.E <= a.Dp,
                "attn_bwd2_tc: unsupported dims (Dp %d, Tp %d)", a.Dp, a.Tp);
  V1T_CHECK_ARG(a.B * a.H <= 65535, "attn_bwd2_tc: too many (batch, head) pairs");
  V1T_CHECK_ARG(a.drop.p <= 0.f || a.drop_bits, "attn_bwd2_tc: dropout needs the keep bits written by the forward");
  switch (a.Dp / 32) {
    case 1: return bwd2_all<1>(a, st);
    case 2: return bwd2_all<2>(a, st);
    case 3: return bwd2_all<3>(a, st);
    case 4: return bwd2_all<4>(a, st);
    default: return bwd2_all<5>(a, st);
  }
}

}  // namespace v1t

extern "C" int v1t_diag_attn_pair_trace(long long* buf) { return v1t::attn_pair_trace_set(buf); }
