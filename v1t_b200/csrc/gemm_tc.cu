// tcgen05 tensor-core GEMM with the same strided/batched/split-K contract as gemm_fp32.cu:
//     C[b][m,n] = alpha * sum_k A[b][m,k] * B[b][k,n] (+ bias[n]) (dropout) (+ R[b][m,n])
// Operands are fp32 in global memory with arbitrary leading dims and either orientation; PRODUCER warps load
// them (128-bit where aligned), split each value into bf16 hi + bf16 lo (x ~= hi + lo), and store both planes
// into shared memory in the K-major 64-byte-swizzled UMMA layout (a transposing store handles MN-contiguous
// sources, so only K-major descriptors are needed).  ONE thread issues tcgen05.mma (M=128, N<=256, K=16):
// hi*hi + lo*hi + hi*lo in "bf16x3" mode (fp32-class accuracy, fp32 accumulate in TMEM) or hi*hi only in "bf16"
// mode.  Accumulators are double-buffered in TMEM (2 x 256 columns) so the EPILOGUE warps (tcgen05.ld -> bias /
// dropout / residual -> global) overlap the next tile's main loop.  Persistent CTAs, one per SM.
//
// Warp roles (576 threads): warps 0-7 epilogue (TMEM lane quarter = warp & 3, column half = warp >> 2), warp 8
// TMEM alloc + MMA issue, warps 9-16 converting producers, warp 17 bulk-copy loader for operands that arrive as
// pre-swizzled bf16 planes (PlaneOp: weights converted once per step, see planes.cu).  Pipelines: full/empty mbarriers per smem stage (4 stages x 48 KB), tmem_full/tmem_empty
// per accumulator buffer.
//
// Plane operands may be batched (PlaneOp::batch_bytes: the T x T problems of the materialised attention, core.cu) and the
// MN-major B operand may come straight from the attention planes (PlaneOp::tile_major).  Epilogue kinds (EpiOp): bias /
// dropout / residual, GELU forward (+ planes), GELU gradient (+ planes + column sums), head planes (q | k | v or dO as
// attention planes), and planes-only output with per-batch row / atom offsets (kEpiPlanesOut: dQ = dS K of attn_bwd2.cu).
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace v1t {
extern int g_use_mn_major;
namespace {

using namespace tc;

#ifndef EPI_UNROLL
#define EPI_UNROLL 1
#endif
constexpr int kEpiUnroll = EPI_UNROLL;  // rows of a 32 x 32 group in flight per epilogue warp (ILP vs code size)
constexpr int BM = 128, BK = 32, STAGES = 2, RAW = 2, BN_MAX = 256;  // UMMA stages, raw fp32 staging slots
constexpr int kEpiWarps = 8, kProdWarps = 8;
constexpr int kLoaderWarp = kEpiWarps + 1 + kProdWarps;       // bulk-copies plane operands (PlaneOp) into the stages
constexpr int kThreads = (kLoaderWarp + 1) * 32;              // 576
constexpr int kProdThreads = kProdWarps * 32;
constexpr int A_PLANE = BM * 64;       // bytes of one bf16 plane of the A stage (128 rows x 32 k)
constexpr int B_PLANE = BN_MAX * 64;
constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;  // hi + lo planes of A and B
// "wide" stages of all-plane launches with N <= 160: BK = 64 (bulk copies cost ~90 cycles each regardless of size, and
// MN-major operands arrive in per-atom pieces of BK * 64 bytes: 2 KB pieces capped an SM at 45 GB/s, see
// scripts/bulk_microbench.py).  Two stages fill exactly the three 48 KB slots.
constexpr int WBK = 64, W_BN = 160, W_A_PLANE = BM * WBK * 2, W_B_PLANE = W_BN * WBK * 2;
constexpr int W_STAGE = 2 * W_A_PLANE + 2 * W_B_PLANE;
static_assert(2 * W_STAGE <= 3 * (2 * (BM * 64) + 2 * (BN_MAX * 64)), "wide stages must fit the first three slots");
constexpr int kEpiStage = 8 * 4096;  // per-warp 32 x 32 fp32 transpose tiles of the epilogue
constexpr int SMEM_BYTES = (STAGES + RAW) * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + kEpiStage;

struct TcArgs {
  GemmArgs g;
  int bn;        // N tile (multiple of 16, <= 256)
  int tiles_m, tiles_n;
  int x3;        // 1: hi*hi + lo*hi + hi*lo, 0: hi*hi
  int mn_a, mn_b;  // operand is M/N-contiguous in memory and staged un-transposed (MN-major UMMA descriptor)
  int a_vec, b_vec, c_vec, r_vec;  // 128-bit access allowed
  PlaneOp pa, pb;  // operands supplied as pre-swizzled bf16 planes (bulk-copied, not converted)
  int a_pl, b_pl;
  int wide;        // all-plane launch with bn <= 160: 64-deep k-blocks in two 72 KB stages (half as many bulk copies)
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct Tile {
  int m0, n0, k_begin, k_end;
  int64_t a_off, b_off, c_off, r_off;
  int64_t pa_off, pb_off;  // byte offsets of this batch entry's plane operands
  int pl_row0, pl_atom0;   // kEpiPlanesOut: first row / column atom of this batch entry in the output planes
};

__device__ __forceinline__ Tile decode_tile(const TcArgs& a, int t) {
  const GemmArgs& g = a.g;
  Tile tl;
  const int nt = t % a.tiles_n; t /= a.tiles_n;
  const int mt = t % a.tiles_m; t /= a.tiles_m;
  const int split = t % g.splits; t /= g.splits;
  const int b2 = t % g.d.batch2, b1 = t / g.d.batch2;
  tl.m0 = mt * BM;
  tl.n0 = nt * a.bn;
  tl.k_begin = split * g.k_chunk;
  tl.k_end = min(g.d.k, tl.k_begin + g.k_chunk);
  tl.a_off = b1 * g.d.a_b1 + b2 * g.d.a_b2;
  tl.b_off = b1 * g.d.b_b1 + b2 * g.d.b_b2;
  tl.c_off = b1 * g.d.c_b1 + b2 * g.d.c_b2 + (int64_t)split * g.c_split;
  tl.r_off = b1 * g.d.r_b1 + b2 * g.d.r_b2;
  tl.pa_off = (int64_t)(b1 * g.d.batch2 + b2) * a.pa.batch_bytes;
  tl.pb_off = (int64_t)(b1 * g.d.batch2 + b2) * a.pb.batch_bytes;
  tl.pl_row0 = b1 * g.epi.pl.b1_rows;
  tl.pl_atom0 = b2 * g.epi.pl.b2_atoms;
  return tl;
}

// Template parameters strip everything a launch does not need (the three warp roles share one instruction cache
// and the kernel is fetch-sensitive): EPI bit 0 = dropout on the main output, bits 1.. = EpiOp kind; ACONV / BCONV =
// the operand is converted from fp32 by the producer warps (false: it arrives as planes through the loader warp).
template <int EPI, bool ACONV, bool BCONV>
__global__ void __launch_bounds__(kThreads, 1) tc_gemm_kernel(const TcArgs a) {
  constexpr bool kDrop = (EPI & 1) != 0;
  constexpr int kKind = EPI >> 1;
  constexpr bool kConv = ACONV || BCONV, kPlanes = !ACONV || !BCONV;
  // with no operand to convert, the raw staging slots become two more UMMA stages (bulk copies need the depth)
  // (three of the four slots; the last one holds the transpose tiles of the extra epilogue warps, see below)
  constexpr int kNS = kConv ? STAGES : STAGES + RAW - 1;
  // ... and the idle conversion warps join the epilogue: 16 instead of 8 warps drain the accumulators
  constexpr int kEpiW = kConv ? kEpiWarps : kEpiWarps + kProdWarps;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (STAGES + RAW) * STAGE_BYTES);
  uint64_t* empty = full + (STAGES + RAW);
  uint64_t* pfull = empty + (STAGES + RAW);  // plane-operand bulk copies landed (expect_tx, loader warp)
  uint64_t* tfull = pfull + (STAGES + RAW);
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* stage_base = smem + (STAGES + RAW) * STAGE_BYTES + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmArgs& g = a.g;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kNS; ++s) {
      mbar_init(&full[s], kProdThreads);
      mbar_init(&pfull[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], kEpiW * 32);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = g.d.batch1 * g.d.batch2 * g.splits * a.tiles_m * a.tiles_n;

  if (warp == kLoaderWarp) {
    // ============================== PLANE LOADER ==============================
    // Per k-block: wait for the stage to drain, then bulk-copy the plane operand tiles (already in the swizzled
    // UMMA layout) from global memory; the copies complete on the stage's pfull barrier.  K-major operand: `rows`
    // consecutive plane rows of column atom k0/32 (one copy per plane).  MN-major: 32 plane rows (k) of each of
    // `atoms` consecutive column atoms (2 KB each).  One copy per lane, so a stage's copies are issued in parallel.
    if (!kConv && a.wide) {
      // ---- wide stages (BK = 64): per k-block the K-major operand is two atoms (two copies per plane), the
      // MN-major operand `atoms` pieces of klen * 64 bytes at a 4 KB pitch; klen = 32 for a ragged last block
      uint32_t it = 0;
      const int np = a.x3 ? 2 : 1;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const Tile tl = decode_tile(a, t);
        const int kend_p = (tl.k_end + 31) & ~31;
        const int num_kb = (tl.k_end - tl.k_begin + WBK - 1) / WBK;
        const int a_rows = min(BM, a.pa.rows_p - tl.m0), a_atoms = max(0, min(BM / 32, a.pa.catoms - tl.m0 / 32));
        const int b_rows = min(a.bn, a.pb.rows_p - tl.n0), b_atoms = max(0, min(a.bn / 32, a.pb.catoms - tl.n0 / 32));
        const int a_items = (a.mn_a ? a_atoms : 2) * np, b_items = (a.mn_b ? b_atoms : 2) * np;
        const bool mine = lane < a_items + b_items;
        const bool is_b = lane >= a_items;
        const int idx = is_b ? lane - a_items : lane;
        const int plane = idx % np, atom = idx / np;  // atom: MN atom (MN-major) or k atom 0/1 (K-major)
        const bool mn = is_b ? a.mn_b : a.mn_a;
        const PlaneOp& po = is_b ? a.pb : a.pa;
        const uint8_t* src_base = (plane ? po.lo : po.hi) + (is_b ? tl.pb_off : tl.pa_off);
        const int row0 = is_b ? tl.n0 : tl.m0;
        const int rows = is_b ? b_rows : a_rows;
        const int pitch_k = (is_b ? a.bn : BM) * 64;  // K-major: bytes between the two k atoms of a stage
        const uint32_t dst_off = (is_b ? 2 * W_A_PLANE : 0) + plane * (is_b ? W_B_PLANE : W_A_PLANE) +
                                 (mn ? atom * (WBK * 64) : atom * pitch_k);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it & 1;
          const uint32_t ph = (it >> 1) & 1;
          mbar_wait(&empty[stage], ph ^ 1);
          const int k0 = tl.k_begin + kb * WBK;
          const int klen = min(WBK, kend_p - k0);  // 32 or 64
          if (lane == 0) {
            const uint32_t ab = a.mn_a ? a_atoms * klen * 64 : (klen >> 5) * a_rows * 64;
            const uint32_t bb = a.mn_b ? b_atoms * klen * 64 : (klen >> 5) * b_rows * 64;
            mbar_expect_tx(&pfull[stage], (ab + bb) * np);
          }
          __syncwarp();
          if (mine && (mn || atom * 32 < klen)) {
            const int64_t src =
                !mn                  ? ((int64_t)((k0 >> 5) + atom) * po.rows_p + row0) * 64
                : po.tile_major == 1 ? ((int64_t)((k0 >> 6) * po.catoms + row0 / 32 + atom) * 64 + (k0 & 63)) * 64
                : po.tile_major == 2 ? ((((int64_t)(row0 >> 7) * (po.rows_p >> 6) + (k0 >> 6)) * 4 + atom) * 64 + (k0 & 63)) * 64
                                     : ((int64_t)(row0 / 32 + atom) * po.rows_p + k0) * 64;
            // a full 64-row k-block of a tile-major operand is contiguous over its atoms, in global memory as in the
            // stage: ONE copy per plane instead of one per atom (a bulk copy costs ~90 cycles on top of its bytes)
            const bool whole = mn && po.tile_major && klen == WBK;
            if (!whole) bulk_g2s(smem + stage * W_STAGE + dst_off, src_base + src, mn ? klen * 64 : rows * 64, &pfull[stage]);
            else if (atom == 0)
              bulk_g2s(smem + stage * W_STAGE + dst_off, src_base + src, (is_b ? b_atoms : a_atoms) * (WBK * 64), &pfull[stage]);
          }
        }
      }
    } else if (kPlanes) {
      uint32_t it = 0;
      const int np = a.x3 ? 2 : 1;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const Tile tl = decode_tile(a, t);
        const int num_kb = (tl.k_end - tl.k_begin + BK - 1) / BK;
        const int a_rows = min(BM, a.pa.rows_p - tl.m0), a_atoms = max(0, min(BM / 32, a.pa.catoms - tl.m0 / 32));
        const int b_rows = min(a.bn, a.pb.rows_p - tl.n0), b_atoms = max(0, min(a.bn / 32, a.pb.catoms - tl.n0 / 32));
        const int a_items = ACONV ? 0 : (a.mn_a ? a_atoms : 1) * np;
        const int b_items = BCONV ? 0 : (a.mn_b ? b_atoms : 1) * np;
        uint32_t bytes = 0;
        if (!ACONV) bytes += (a.mn_a ? a_atoms * 2048 : a_rows * 64) * np;
        if (!BCONV) bytes += (a.mn_b ? b_atoms * 2048 : b_rows * 64) * np;
        // this lane's copy (constant over the k loop except for the k offset)
        const bool mine = lane < a_items + b_items;
        const bool is_b = lane >= a_items;
        const int idx = is_b ? lane - a_items : lane;
        const int plane = idx % np, atom = idx / np;
        const bool mn = is_b ? a.mn_b : a.mn_a;
        const PlaneOp& po = is_b ? a.pb : a.pa;
        const uint8_t* src_base = (plane ? po.lo : po.hi) + (is_b ? tl.pb_off : tl.pa_off);
        const int row0 = is_b ? tl.n0 : tl.m0;
        const uint32_t dst_off = (is_b ? 2 * A_PLANE : 0) + plane * (is_b ? B_PLANE : A_PLANE) + (mn ? atom * 2048 : 0);
        const uint32_t cbytes = mn ? 2048u : (uint32_t)((is_b ? b_rows : a_rows) * 64);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % kNS;
          const uint32_t ph = (it / kNS) & 1;
          mbar_wait(&empty[stage], ph ^ 1);
          const int k0 = tl.k_begin + kb * BK;
          if (lane == 0) mbar_expect_tx(&pfull[stage], bytes);
          __syncwarp();
          if (mine) {
            const int64_t src = mn ? ((int64_t)(row0 / 32 + atom) * po.rows_p + k0) * 64
                                   : ((int64_t)(k0 >> 5) * po.rows_p + row0) * 64;
            bulk_g2s(smem + stage * STAGE_BYTES + dst_off, src_base + src, cbytes, &pfull[stage]);
          }
        }
      }
    }
  } else if (warp > kEpiWarps && kConv) {
    // ============================== PRODUCERS ==============================
    // Two-level pipeline.  (1) cp.async copies the raw fp32 operand chunks of k-block i+RAW into a per-thread
    // staging slot (each thread later reads only what it copied itself, so no cross-thread synchronisation is
    // needed and the global-memory latency of RAW k-blocks is in flight without holding registers);
    // (2) the same thread reads its chunks back, splits them into bf16 hi/lo and stores them into the swizzled
    // UMMA stage.  Thread -> chunk assignment is fixed per launch.
    const int ptid = threadIdx.x - (kEpiWarps + 1) * 32;
    const bool a_kc = (g.d.a_k == 1), b_kc = (g.d.b_k == 1);
    const int b_chunks = BCONV ? a.bn * 4 : 0;  // plane operands are not converted here
    uint8_t* raw = smem + STAGES * STAGE_BYTES;

    // ---- per-thread chunk geometry (constant for the whole launch: no divisions in the k loop) ----
    // A chunk = 8 consecutive elements along the memory-contiguous direction ("c" coordinate) of one line ("l"
    // coordinate).  K-contiguous operand: l = row (m or n), c = k.  MN-major staging: l = k, c = row.
    struct Chunk { int lc; uint32_t soff; int goff; };  // lc = (l << 12) | c
    Chunk ca[2], cb[4];
    const int64_t a_sl = a.mn_a ? g.d.a_k : g.d.a_m, a_sc = a.mn_a ? g.d.a_m : g.d.a_k;
    const int64_t b_sl = a.mn_b ? g.d.b_k : g.d.b_n, b_sc = a.mn_b ? g.d.b_n : g.d.b_k;
    const bool a_vec = (a_sc == 1) && a.a_vec, b_vec = (b_sc == 1) && a.b_vec;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int q = ptid + i * kProdThreads;
      int l, c;
      if (a.mn_a) { l = q >> 4; c = (q & 15) * 8; ca[i].soff = ((q & 15) >> 2) * (BK * 64) + sw64_offset(q >> 4, q & 3); }
      else {
        const int r = a_kc ? (q >> 2) : (q % BM), cc = a_kc ? (q & 3) : (q / BM);
        l = r; c = cc * 8; ca[i].soff = sw64_offset(r, cc);
      }
      ca[i].lc = (l << 12) | c;
      ca[i].goff = (int)(l * a_sl + c * a_sc);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = ptid + i * kProdThreads;
      int l, c;
      if (a.mn_b) {
        const int ncw = a.bn >> 3;
        const int kr = q / ncw, nc = q % ncw;
        l = kr; c = nc * 8; cb[i].soff = (nc >> 2) * (BK * 64) + sw64_offset(kr, nc & 3);
      } else {
        const int r = b_kc ? (q >> 2) : (q % a.bn), cc = b_kc ? (q & 3) : (q / a.bn);
        l = r; c = cc * 8; cb[i].soff = sw64_offset(r, cc);
      }
      cb[i].lc = (l << 12) | c;
      cb[i].goff = (int)(l * b_sl + c * b_sc);
    }

    // prefetch cursor over (tile, k-block); the consumer side only needs the iteration count
    int pf_t = blockIdx.x, pf_kb = 0, pf_nkb = 0;
    Tile pf_tl{};
    int total_iters = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const Tile tl = decode_tile(a, t);
      total_iters += (tl.k_end - tl.k_begin + BK - 1) / BK;
    }
    auto pf_load = [&]() {
      if (pf_t < total_tiles) {
        pf_tl = decode_tile(a, pf_t);
        pf_nkb = (pf_tl.k_end - pf_tl.k_begin + BK - 1) / BK;
      }
    };
    auto pf_next = [&]() {
      if (++pf_kb >= pf_nkb) {
        pf_kb = 0;
        pf_t += gridDim.x;
        pf_load();
      }
    };
    // copy 8 elements starting at p (element stride sc) into the 32-byte staging slot; zero-fill beyond limits
    auto copy_chunk = [&](uint8_t* dst, const float* p, bool in_line, int nv, int64_t sc, bool vec) {
      const uint32_t d = smem_u32(dst);
      if (!in_line || nv <= 0) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dst + 16) = make_uint4(0, 0, 0, 0);
      } else if (vec && nv >= 8) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(p) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16), "l"(p + 4) : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < nv) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4 * j), "l"(p + (int64_t)j * sc) : "memory");
          else *reinterpret_cast<float*>(dst + 4 * j) = 0.f;
        }
      }
    };
    auto issue_stage = [&](int rs) {
      const int k0 = pf_tl.k_begin + pf_kb * BK;
      uint8_t* rbase = raw + rs * STAGE_BYTES + ptid * 32;
      {
        const int l0 = a.mn_a ? k0 : pf_tl.m0, c0 = a.mn_a ? pf_tl.m0 : k0;
        const int l_lim = a.mn_a ? pf_tl.k_end : g.d.m, c_lim = a.mn_a ? g.d.m : pf_tl.k_end;
        const float* base = g.A + pf_tl.a_off + (int64_t)l0 * a_sl + (int64_t)c0 * a_sc;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (ACONV) copy_chunk(rbase + i * (kProdThreads * 32), base + ca[i].goff, l0 + (ca[i].lc >> 12) < l_lim,
                     c_lim - (c0 + (ca[i].lc & 4095)), a_sc, a_vec);
      }
      {
        const int l0 = a.mn_b ? k0 : pf_tl.n0, c0 = a.mn_b ? pf_tl.n0 : k0;
        const int l_lim = a.mn_b ? pf_tl.k_end : g.d.n, c_lim = a.mn_b ? g.d.n : pf_tl.k_end;
        const float* base = g.B + pf_tl.b_off + (int64_t)l0 * b_sl + (int64_t)c0 * b_sc;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (BCONV && ptid + i * kProdThreads < b_chunks)
            copy_chunk(rbase + (2 + i) * (kProdThreads * 32), base + cb[i].goff, l0 + (cb[i].lc >> 12) < l_lim,
                       c_lim - (c0 + (cb[i].lc & 4095)), b_sc, b_vec);
      }
    };

    pf_load();
    // iterations -RAW..-1 only prefetch (fill the raw slots); iteration `it` converts k-block it and refills its slot
    for (int it = -RAW; it < total_iters; ++it) {
      const int rs = (it + RAW) % RAW;
      if (it >= 0) {
        asm volatile("cp.async.wait_group %0;" ::"n"(RAW - 1) : "memory");
        const int stage = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const uint8_t* rbase = raw + rs * STAGE_BYTES + ptid * 32;
        mbar_wait(&empty[stage], ph ^ 1);
        const uint32_t sa_hi = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sb_hi = sa_hi + 2 * A_PLANE;
        const uint32_t rb = smem_u32(rbase);
        // chunk by chunk: staging slot -> registers -> bf16 hi/lo -> swizzled UMMA stage
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          if (i < 2 ? ACONV : (BCONV && ptid + (i - 2) * kProdThreads < b_chunks)) {
            const uint4 x = lds128(rb + i * (kProdThreads * 32));
            const uint4 y = lds128(rb + i * (kProdThreads * 32) + 16);
            const float v[8] = {__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w),
                                __uint_as_float(y.x), __uint_as_float(y.y), __uint_as_float(y.z), __uint_as_float(y.w)};
            const uint32_t dst = (i < 2 ? sa_hi + ca[i].soff : sb_hi + cb[i - 2].soff);
            if (a.x3) {
              uint4 hi, lo;
              split8(v, hi, lo);
              sts128(dst, hi);
              sts128(dst + (i < 2 ? A_PLANE : B_PLANE), lo);
            } else {  // plain-bf16 mode: four packs instead of the ~30-instruction hi/lo split (the converter warps, not the
                      // single-pass MMAs, bound these launches: the materialised attention of the scaled core)
              uint4 hi;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
              hi.x = *reinterpret_cast<uint32_t*>(&h0); hi.y = *reinterpret_cast<uint32_t*>(&h1);
              hi.z = *reinterpret_cast<uint32_t*>(&h2); hi.w = *reinterpret_cast<uint32_t*>(&h3);
              sts128(dst, hi);
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&full[stage]);
      }
      // refill this raw slot with k-block it + RAW (after its contents were consumed above)
      if (pf_t < total_tiles) {
        issue_stage(rs);
        pf_next();
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (warp == kEpiWarps) {
    // ============================== MMA ISSUER ==============================
    const uint32_t idesc = idesc_bf16(BM, a.bn, a.mn_a, a.mn_b);
    const bool leader = elect_one();  // one lane issues every tcgen05.mma / commit; address math stays warp-uniform
    uint32_t it = 0, tl_i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl_i) {
      const Tile tl = decode_tile(a, t);
      const int num_kb = (tl.k_end - tl.k_begin + BK - 1) / BK;
      const uint32_t buf = tl_i & 1, tph = (tl_i >> 1) & 1;
      mbar_wait(&tempty[buf], tph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * BN_MAX;
      if (!kConv && a.wide) {  // ---- wide stages: up to four 16-deep MMA steps per k-block
        const int kend_p = (tl.k_end + 31) & ~31;
        const int nkb = (tl.k_end - tl.k_begin + WBK - 1) / WBK;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int stage = it & 1;
          mbar_wait(&pfull[stage], (it >> 1) & 1);
          tc_fence_after();
          const int klen = min(WBK, kend_p - (tl.k_begin + kb * WBK));
          const uint32_t sa_hi = smem_u32(smem + stage * W_STAGE);
          const uint32_t sa_lo = sa_hi + W_A_PLANE, sb_hi = sa_hi + 2 * W_A_PLANE, sb_lo = sb_hi + W_B_PLANE;
          const uint32_t pitch_b = a.bn * 64;
#pragma unroll
          for (int kk = 0; kk < WBK / 16; ++kk) {
            if (kk * 16 < klen) {
              const uint32_t ka = (kk >> 1) * (BM * 64) + (kk & 1) * 32, kbo = (kk >> 1) * pitch_b + (kk & 1) * 32;
              const uint64_t ah = a.mn_a ? desc_mn_sw64(sa_hi + kk * 1024, WBK * 64) : desc_k_sw64(sa_hi + ka);
              const uint64_t bh = a.mn_b ? desc_mn_sw64(sb_hi + kk * 1024, WBK * 64) : desc_k_sw64(sb_hi + kbo);
              const uint64_t al = a.mn_a ? desc_mn_sw64(sa_lo + kk * 1024, WBK * 64) : desc_k_sw64(sa_lo + ka);
              const uint64_t bl = a.mn_b ? desc_mn_sw64(sb_lo + kk * 1024, WBK * 64) : desc_k_sw64(sb_lo + kbo);
              if (leader) {
                umma_bf16(d_tmem, ah, bh, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
                if (a.x3) {
                  umma_bf16(d_tmem, al, bh, idesc, 1u);
                  umma_bf16(d_tmem, ah, bl, idesc, 1u);
                }
              }
            }
          }
          if (leader) {
            umma_commit(&empty[stage]);
            if (kb == nkb - 1) umma_commit(&tfull[buf]);
          }
          __syncwarp();
        }
        continue;
      }
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int stage = it % kNS;
        const uint32_t ph = (it / kNS) & 1;
        if (kConv) mbar_wait(&full[stage], ph);     // converted operands stored
        if (kPlanes) mbar_wait(&pfull[stage], ph);  // plane operands landed
        tc_fence_after();
        {
          const uint32_t sa_hi = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sa_lo = sa_hi + A_PLANE, sb_hi = sa_hi + 2 * A_PLANE, sb_lo = sb_hi + B_PLANE;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            // K-major: advance 16 k = 32 B inside the 64 B atom row; MN-major: advance 16 k rows = 1024 B,
            // atoms along M/N are BK*64 = 2048 B apart (LBO)
            const uint64_t ah = a.mn_a ? desc_mn_sw64(sa_hi + kk * 1024, BK * 64) : desc_k_sw64(sa_hi + kk * 32);
            const uint64_t bh = a.mn_b ? desc_mn_sw64(sb_hi + kk * 1024, BK * 64) : desc_k_sw64(sb_hi + kk * 32);
            const uint64_t al = a.mn_a ? desc_mn_sw64(sa_lo + kk * 1024, BK * 64) : desc_k_sw64(sa_lo + kk * 32);
            const uint64_t bl = a.mn_b ? desc_mn_sw64(sb_lo + kk * 1024, BK * 64) : desc_k_sw64(sb_lo + kk * 32);
            if (leader) {
              umma_bf16(d_tmem, ah, bh, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
              if (a.x3) {
                umma_bf16(d_tmem, al, bh, idesc, 1u);
                umma_bf16(d_tmem, ah, bl, idesc, 1u);
              }
            }
          }
          if (leader) {
            umma_commit(&empty[stage]);                      // smem slot free once these MMAs retire
            if (kb == num_kb - 1) umma_commit(&tfull[buf]);  // accumulator complete
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ============================== EPILOGUE ==============================
    // 8 warps (16 when no operand needs converting: warps 9-16 then drain accumulators too): TMEM lane quarter =
    // warp & 3 (row block, fixed by the hardware); the warps sharing a quarter take alternate 32-column groups.
    // Every 32 x 32 group is transposed through a warp-private smem tile, so that a lane owns 4 adjacent columns
    // of one row and each 128-bit store instruction writes four full 128-byte row segments.  Ragged edges and
    // unaligned operands use the same path with element-wise memory accesses.  The row loop is deliberately NOT
    // unrolled: with dropout (Philox) and GELU inlined, an unrolled body overflowed the instruction cache and the
    // epilogue became fetch-bound ("no_instructions" stalls, profiles/r1_gemm_epilogue_icache.txt).
    uint32_t tl_i = 0;
    const int quarter = warp & 3;
    // column slot among the warps of this quarter: warps 0-7 -> 0, 1; warps 9-16 -> 2, 3
    const int slot = warp < kEpiWarps ? (warp >> 2) : 2 + ((warp - kEpiWarps - 1) >> 2);
    constexpr int kSlots = kEpiW / 4;
    const float inv_keep = (kDrop && g.drop.p > 0.f) ? 1.f / (1.f - g.drop.p) : 1.f;
    const int64_t drop_ld = drop_stride(g.d.n);
    const float inv_keep2 = g.epi.drop.p > 0.f ? 1.f / (1.f - g.epi.drop.p) : 1.f;
    const bool bias_vec = g.bias && ((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
    const uint32_t stg = warp < kEpiWarps ? smem_u32(stage_base + warp * 4096)
                                          : smem_u32(smem + (STAGES + RAW - 1) * STAGE_BYTES + (warp - kEpiWarps - 1) * 4096);
    const int cq = lane & 7;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tl_i) {
      const Tile tl = decode_tile(a, t);
      const uint32_t buf = tl_i & 1, tph = (tl_i >> 1) & 1;
      mbar_wait(&tfull[buf], tph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * BN_MAX + ((uint32_t)(quarter * 32) << 16);
      const int n_lim = min(g.d.n, tl.n0 + a.bn);
      int hp_b0 = 0, hp_t0 = 0;
      if constexpr (kKind == kEpiHeadPlanes) {
        const int r0 = tl.m0 + quarter * 32 + (lane >> 3);
        hp_b0 = r0 / g.epi.hp.T;
        hp_t0 = r0 - hp_b0 * g.epi.hp.T;
      }
      // 32-column groups alternate between the two column halves (columns past bn are never stored)
      for (int c0 = slot * 32; c0 < a.bn; c0 += 32 * kSlots) {
        const int nb = tl.n0 + c0;
        {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sts128(stg + lane * 128 + ((c ^ (lane & 7)) << 4), make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
        }
        __syncwarp();
        const int n4 = nb + 4 * cq;
        const int nv = min(4, n_lim - n4);  // valid columns of this lane's quad
        // kEpiHeadPlanes: this 32-column group is one head-dim atom of q, k or v of one head
        uint8_t *hp_hi = nullptr, *hp_lo = nullptr;
        int hp_atom = 0, hp_head = 0;
        if constexpr (kKind == kEpiHeadPlanes) {
          const HeadPlanes& hp = g.epi.hp;
          const int sec = nb >> 5, s3 = min(sec / (hp.AD * hp.H), 2);
          hp_atom = sec % hp.AD;
          hp_head = (sec / hp.AD) % hp.H;
          hp_hi = hp.p[s3][0]; hp_lo = hp.p[s3][1];
        }
        float bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (g.bias && nv > 0) {
          if (bias_vec && nv == 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n4));
            bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (e < nv) bv[e] = __ldg(g.bias + n4 + e);
          }
        }
        // The pre-activation (GELU gradient) and residual quads of row-iteration i + 1 are fetched while iteration
        // i computes: with the loop rolled, an un-prefetched global load would stall every iteration for its full
        // latency (8 epilogue warps cannot hide it).
        const int row0 = tl.m0 + quarter * 32 + (lane >> 3);
        const bool pre_u = kKind == kEpiGeluGrad && nv > 0;
        const bool pre_r = g.R && a.r_vec && nv == 4;
        float4 u_nx = make_float4(0.f, 0.f, 0.f, 0.f), r_nx = u_nx;
        auto prefetch = [&](int i) {
          const int mm = row0 + 4 * i;
          if (mm < g.d.m) {
            if (pre_u) u_nx = __ldg(reinterpret_cast<const float4*>(g.epi.u + (int64_t)mm * g.epi.ld + n4));
            if (pre_r) r_nx = __ldg(reinterpret_cast<const float4*>(g.R + tl.r_off + (int64_t)mm * g.d.r_m + n4));
          }
        };
        prefetch(0);
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll kEpiUnroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + (lane >> 3);
          const int mm = tl.m0 + quarter * 32 + r;
          const float4 uv = u_nx, rv = r_nx;
          if (i + 1 < 8) prefetch(i + 1);
          if (mm >= g.d.m || nv <= 0) {  // past the last row / column: only the zero padding of the emitted planes
            if (kKind != kEpiNone && kKind != kEpiPlanesOut && g.epi.pl.hi && mm < g.epi.pl.rows_p &&
                n4 < 32 * ((g.d.n + 31) / 32)) {
              const int64_t off = plane_chunk_off(n4 >> 5, g.epi.pl.rows_p, mm, cq >> 1) + (cq & 1) * 8;
              *reinterpret_cast<uint2*>(g.epi.pl.hi + off) = make_uint2(0u, 0u);
              if (g.epi.pl.lo) *reinterpret_cast<uint2*>(g.epi.pl.lo + off) = make_uint2(0u, 0u);
            }
            continue;
          }
          const uint4 xr = lds128(stg + r * 128 + ((cq ^ (r & 7)) << 4));
          float o[4] = {fmaf(g.d.alpha, __uint_as_float(xr.x), bv[0]), fmaf(g.d.alpha, __uint_as_float(xr.y), bv[1]),
                        fmaf(g.d.alpha, __uint_as_float(xr.z), bv[2]), fmaf(g.d.alpha, __uint_as_float(xr.w), bv[3])};
          if constexpr (kKind == kEpiHeadPlanes) {
            // this 32-column group is one head-dim atom of q, k or v of one head: row (b, t) of its plane slab
            // (sample, token) of this row: one division per tile (hp_b0 / hp_t0), then a wrap at sample boundaries
            int t = hp_t0 + 4 * i, bb = hp_b0;
            while (t >= g.epi.hp.T) { t -= g.epi.hp.T; ++bb; }
            const int64_t off = attn_plane_off((int64_t)bb * g.epi.hp.H + hp_head, hp_atom, t, g.epi.hp.Tq, g.epi.hp.AD) +
                                (((cq >> 1) ^ ((t >> 1) & 3)) << 4) + (cq & 1) * 8;
            uint2 ph, plo;
            split4(o, ph, plo);
            *reinterpret_cast<uint2*>(hp_hi + off) = ph;
            if (hp_lo) *reinterpret_cast<uint2*>(hp_lo + off) = plo;
            if (!g.C) continue;
          }
          const bool full4 = (nv == 4);
          if (kDrop && g.drop.p > 0.f) {
            float mk[4];
            dropout_mult4(g.drop.seed, g.drop.site, ((uint64_t)mm * drop_ld + n4) >> 2, g.drop.p, inv_keep, mk);
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] *= mk[e];
          }
          float mk2[4] = {1.f, 1.f, 1.f, 1.f};
          if (kKind != kEpiNone && g.epi.drop.p > 0.f)
            dropout_mult4(g.epi.drop.seed, g.epi.drop.site, ((uint64_t)mm * drop_ld + n4) >> 2, g.epi.drop.p, inv_keep2, mk2);
          if constexpr (kKind == kEpiGeluGrad) {
            // epi.ld % 4 == 0 and 16-byte aligned rows are checked on the host: the quad is always readable
            o[0] *= gelu_fast_df(uv.x) * mk2[0]; o[1] *= gelu_fast_df(uv.y) * mk2[1];
            o[2] *= gelu_fast_df(uv.z) * mk2[2]; o[3] *= gelu_fast_df(uv.w) * mk2[3];
          }
          if (g.R) {
            if (pre_r) {
              o[0] += rv.x; o[1] += rv.y; o[2] += rv.z; o[3] += rv.w;
            } else {
              const float* rp = g.R + tl.r_off + (int64_t)mm * g.d.r_m + n4;
#pragma unroll
              for (int e = 0; e < 4; ++e) if (e < nv) o[e] += __ldg(rp + e);
            }
          }
          if constexpr (kKind == kEpiGeluGrad) {  // column sums of the result (bias gradient), registers per lane
#pragma unroll
            for (int e = 0; e < 4; ++e) cs[e] += (e < nv) ? o[e] : 0.f;
          }
          if (g.C) {
            float* cp = g.C + tl.c_off + (int64_t)mm * g.d.c_m + n4;
            if (a.c_vec && full4) {
              if (g.d.accumulate) {
                const float4 cv = *reinterpret_cast<const float4*>(cp);
                o[0] += cv.x; o[1] += cv.y; o[2] += cv.z; o[3] += cv.w;
              }
              *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (e < nv) {
                  if (g.d.accumulate) o[e] += cp[e];
                  cp[e] = o[e];
                }
            }
          }
          if constexpr (kKind == kEpiGeluOut) {
            const float ge[4] = {gelu_fast_f(o[0]) * mk2[0], gelu_fast_f(o[1]) * mk2[1], gelu_fast_f(o[2]) * mk2[2], gelu_fast_f(o[3]) * mk2[3]};
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = ge[e];  // o now holds the activation (what the planes carry)
            if (g.epi.aux) {
              float* ap = g.epi.aux + (int64_t)mm * g.epi.ld + n4;
              if (full4) *reinterpret_cast<float4*>(ap) = make_float4(ge[0], ge[1], ge[2], ge[3]);
              else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (e < nv) ap[e] = ge[e];
              }
            }
          }
          if (kKind != kEpiNone && g.epi.pl.hi) {  // operand planes of the activation-side result
            const float pv[4] = {o[0], nv > 1 ? o[1] : 0.f, nv > 2 ? o[2] : 0.f, nv > 3 ? o[3] : 0.f};
            uint2 ph, plo;
            split4(pv, ph, plo);
            const int64_t off = plane_chunk_off(tl.pl_atom0 + (n4 >> 5), g.epi.pl.rows_p, tl.pl_row0 + mm, cq >> 1) + (cq & 1) * 8;
            *reinterpret_cast<uint2*>(g.epi.pl.hi + off) = ph;
            if (g.epi.pl.lo) *reinterpret_cast<uint2*>(g.epi.pl.lo + off) = plo;
          }
        }
        if constexpr (kKind == kEpiGeluGrad) {
          if (g.epi.cs_part) {  // 32-row partial column sums of this group -> [tile_m * 4 + quarter][n]
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);
              cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
            }
            if (lane < 8 && nv > 0) {
              float* pp = g.epi.cs_part + (int64_t)((tl.m0 / BM) * 4 + quarter) * g.d.n + n4;
#pragma unroll
              for (int e = 0; e < 4; ++e) if (e < nv) pp[e] = cs[e];
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&tempty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc<512>(tmem_base);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
int g_use_mn_major = 1;  // stage M/N-contiguous operands un-transposed (vector loads) with MN-major descriptors
namespace {

// N tile: an MN-major plane B operand is copied in whole 32-column atoms
int pick_bn(const v1t_gemm_desc& d, const PlaneOp& pb, const EpiOp& epi = no_epi()) {
  const int nt = cdiv(d.n, BN_MAX);
  const bool whole_atoms = (pb.hi && d.b_k != 1) || epi.pl.hi || epi.kind == kEpiHeadPlanes;
  // plane-emitting epilogues are the bottleneck of their launches and run on 4 column slots per TMEM quarter: keep
  // the 32-column groups of every tile a multiple of 4 (full 256-wide tiles plus one narrower tail tile)
  if (nt > 1 && (epi.pl.hi || epi.kind == kEpiHeadPlanes)) return BN_MAX;
  return (int)round_up(cdiv(d.n, nt), whole_atoms ? 32 : 16);
}

int launch_tc(const v1t_gemm_desc& d, const float* A, const float* B, float* C, const float* bias, const float* R,
              DropSpec drop, int x3, int splits, int k_chunk, int64_t c_split, cudaStream_t st, EpiOp epi = no_epi(),
              PlaneOp pa = no_plane(), PlaneOp pb = no_plane()) {
  TcArgs a;
  a.g.d = d;
  a.g.A = A; a.g.B = B; a.g.C = C; a.g.bias = bias; a.g.R = R;
  a.g.drop = drop;
  a.g.epi = epi;
  a.g.splits = splits; a.g.k_chunk = k_chunk; a.g.c_split = c_split;
  a.pa = pa; a.pb = pb;
  a.a_pl = pa.hi != nullptr; a.b_pl = pb.hi != nullptr;
  if (a.a_pl || a.b_pl)
    V1T_CHECK_ARG(((d.batch1 == 1 && d.batch2 == 1) || ((!a.a_pl || pa.batch_bytes > 0) && (!a.b_pl || pb.batch_bytes > 0))) &&
                      (!x3 || ((!a.a_pl || pa.lo) && (!a.b_pl || pb.lo))),
                  "tc gemm: plane operands need an unbatched problem (or a batch stride) and, in bf16x3 mode, the lo plane");
  a.bn = pick_bn(d, pb, epi);
  a.tiles_n = cdiv(d.n, a.bn);
  a.tiles_m = cdiv(d.m, BM);
  a.x3 = x3;
  a.wide = (a.a_pl && a.b_pl && a.bn <= W_BN) ? 1 : 0;
  V1T_CHECK_ARG((!pa.tile_major || (pa.tile_major == 2 && a.wide && d.a_k != 1 && pa.rows_p % 64 == 0 && pa.catoms % 4 == 0)) &&
                    (!pb.tile_major || (pb.tile_major == 1 && a.wide && d.b_k != 1)),
                "tc gemm: tile-major operands are supported as MN-major operands of all-plane launches (A: layout 2, B: layout 1)");
  // plane operands cannot be transposed while staging: their orientation follows the problem
  a.mn_a = a.a_pl ? (d.a_k != 1) : (g_use_mn_major && d.a_k != 1 && d.a_m == 1);
  a.mn_b = a.b_pl ? (d.b_k != 1) : (g_use_mn_major && d.b_k != 1 && d.b_n == 1 && a.bn % 32 == 0);
  // row stride of the staged rows: a_m / b_n for K-contiguous sources, a_k / b_k for MN-major staging
  a.a_vec = aligned16(A) && (a.mn_a ? d.a_k : d.a_m) % 4 == 0 && d.a_b1 % 4 == 0 && d.a_b2 % 4 == 0;
  a.b_vec = aligned16(B) && (a.mn_b ? d.b_k : d.b_n) % 4 == 0 && d.b_b1 % 4 == 0 && d.b_b2 % 4 == 0;
  a.c_vec = aligned16(C) && d.c_m % 4 == 0 && d.c_b1 % 4 == 0 && d.c_b2 % 4 == 0 && c_split % 4 == 0;
  a.r_vec = R && aligned16(R) && d.r_m % 4 == 0 && d.r_b1 % 4 == 0 && d.r_b2 % 4 == 0;
  const int64_t total = (int64_t)d.batch1 * d.batch2 * splits * a.tiles_m * a.tiles_n;
  V1T_CHECK_ARG(total < (1ll << 31), "tc gemm: too many tiles");
  const int grid = (int)std::min<int64_t>(total, kNumSMs);
  // instantiation table: [epilogue variant][A converted][B converted]
  using Kern = void (*)(const TcArgs);
#define V1T_TC_ROW(E) {{tc_gemm_kernel<E, false, false>, tc_gemm_kernel<E, false, true>}, \
                       {tc_gemm_kernel<E, true, false>, tc_gemm_kernel<E, true, true>}}
  static const Kern table[6][2][2] = {V1T_TC_ROW(0), V1T_TC_ROW(1), V1T_TC_ROW(kEpiGeluOut << 1),
                                      V1T_TC_ROW(kEpiGeluGrad << 1), V1T_TC_ROW(kEpiHeadPlanes << 1),
                                      V1T_TC_ROW(kEpiPlanesOut << 1)};
#undef V1T_TC_ROW
  static bool attr_set[6][2][2] = {};
  V1T_CHECK_ARG(epi.kind == kEpiNone || drop.p <= 0.f, "tc gemm: a fused activation excludes dropout on the main output");
  const int ev = epi.kind == kEpiGeluOut ? 2 : epi.kind == kEpiGeluGrad ? 3 : epi.kind == kEpiHeadPlanes ? 4 :
                 epi.kind == kEpiPlanesOut ? 5 : (drop.p > 0.f ? 1 : 0);
  const int ac = a.a_pl ? 0 : 1, bc = a.b_pl ? 0 : 1;
  Kern kern = table[ev][ac][bc];
  if (!attr_set[ev][ac][bc]) {
    V1T_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set[ev][ac][bc] = true;
  }
  kern<<<grid, kThreads, SMEM_BYTES, st>>>(a);
  V1T_LAUNCH_CHECK();
  return V1T_OK;
}

bool tc_supported(const v1t_gemm_desc& d) {
  return d.k > 0 && (d.a_k == 1 || d.a_m == 1) && (d.b_k == 1 || d.b_n == 1);
}

}  // namespace

int gemm_tc(const v1t_gemm_desc& d, const float* A, const float* B, float* C, const float* bias, const float* R,
            cudaStream_t st, DropSpec drop, int x3, EpiOp epi, PlaneOp pa, PlaneOp pb) {
  V1T_CHECK_ARG(d.m >= 0 && d.n >= 0 && d.k >= 0 && d.batch1 >= 1 && d.batch2 >= 1, "gemm: bad sizes");
  if (d.m == 0 || d.n == 0) return V1T_OK;
  if (!tc_supported(d)) return gemm_fp32(d, A, B, C, bias, R, st, drop);
  if (epi.kind == kEpiHeadPlanes) {
    const HeadPlanes& hp = epi.hp;
    const int per = hp.H * hp.AD * 32;
    V1T_CHECK_ARG(d.batch1 == 1 && d.batch2 == 1 && !bias && !R && !d.accumulate && hp.p[0][0] && per > 0 &&
                      (d.n == per || (d.n == 3 * per && hp.p[1][0] && hp.p[2][0])) && hp.T > 0 && d.m % hp.T == 0 &&
                      hp.Tq >= hp.T,
                  "gemm_tc: head-plane output does not match the problem");
  } else if (epi.kind == kEpiPlanesOut) {
    V1T_CHECK_ARG(epi.pl.hi && !C && !bias && !R && !d.accumulate && d.n % 32 == 0 && d.n <= BN_MAX &&
                      ((d.batch1 == 1 && d.batch2 == 1) || (epi.pl.b1_rows >= d.m && epi.pl.b2_atoms >= d.n / 32)),
                  "gemm_tc: plane-only output needs whole 32-column atoms in one N tile and per-batch row / atom offsets");
  } else if (epi.kind != kEpiNone) {
    V1T_CHECK_ARG(d.batch1 == 1 && d.batch2 == 1 && epi.ld % 4 == 0 && aligned16(epi.kind == kEpiGeluOut ? (const void*)epi.aux : (const void*)epi.u),
                  "gemm_tc: fused activation needs an unbatched problem and 16-byte aligned rows");
    V1T_CHECK_ARG(epi.kind != kEpiGeluOut || epi.aux || epi.pl.hi, "gemm_tc: GELU output has no destination");
    // emitted planes must be covered by the N tiling (whole 32-column atoms) and by the rows of the problem
    V1T_CHECK_ARG(!epi.pl.hi || (cdiv(d.n, pick_bn(d, pb, epi)) * pick_bn(d, pb, epi) >= 32 * cdiv(d.n, 32) && epi.pl.rows_p >= d.m &&
                                 d.n % 4 == 0),
                  "gemm_tc: plane output does not fit the tiling");
  }
  return launch_tc(d, A, B, C, bias, R, drop, x3, 1, (int)round_up(d.k, BK), 0, st, epi, pa, pb);
}

int gemm_tc_splitk(const v1t_gemm_desc& d, const float* A, const float* B, float* C, float* partials,
                   size_t partial_bytes, cudaStream_t st, int x3, PlaneOp pa, PlaneOp pb, GroupMap gm) {
  V1T_CHECK_ARG(d.batch1 == 1 && d.batch2 == 1, "splitk gemm: no batch");
  if (d.m == 0 || d.n == 0) return V1T_OK;
  if (!tc_supported(d)) return gemm_fp32_splitk(d, A, B, C, partials, partial_bytes, st);
  const int bn = pick_bn(d, pb);
  const int tiles = cdiv(d.n, bn) * cdiv(d.m, BM);
  const int64_t n_ld = round_up(d.n, 4);
  const int64_t per = (int64_t)d.m * n_ld * (int64_t)sizeof(float);
  // one wave of the persistent grid: as many K chunks as fit the SMs (tiles * splits <= 148), so that no CTA gets a
  // second work item while others idle, and the partials stay small
  int splits = std::max(1, kNumSMs / tiles);
  splits = std::min(splits, cdiv(d.k, 4 * BK));
  splits = (int)std::min<int64_t>(splits, (int64_t)partial_bytes / per);
  const bool mapped = gm.row_gout > 0 || gm.col_gout > 0;  // a remapped product always goes through the partials
  if (splits <= 1 && !mapped) return gemm_tc(d, A, B, C, nullptr, nullptr, st, no_drop(), x3, no_epi(), pa, pb);
  splits = std::max(splits, 1);
  const int k_chunk = (int)round_up(cdiv(d.k, splits), (pa.hi && pb.hi) ? WBK : BK);  // whole (wide) stages per split
  splits = cdiv(d.k, k_chunk);
  v1t_gemm_desc p = d;
  p.accumulate = 0;
  p.c_m = n_ld;
  V1T_TRY(launch_tc(p, A, B, partials, nullptr, nullptr, no_drop(), x3, splits, k_chunk, (int64_t)d.m * n_ld, st, no_epi(),
                    pa, pb));
  return reduce_partials_ld(partials, C, splits, d.m, d.n, n_ld, d.c_m, d.accumulate, st, gm);
}

}  // namespace v1t

extern "C" int v1t_gemm_tc_set_mn_major(int on) {
  v1t::g_use_mn_major = on != 0;
  return V1T_OK;
}

extern "C" int v1t_gemm_tc(const v1t_gemm_desc* d, const float* A, const float* B, float* C, const float* bias,
                           const float* R, int impl, void* stream) {
  V1T_CHECK_ARG(d && A && B && C, "v1t_gemm_tc: null argument");
  V1T_CHECK_ARG(impl == V1T_IMPL_BF16X3 || impl == V1T_IMPL_BF16, "v1t_gemm_tc: impl must be BF16X3 or BF16");
  return v1t::gemm_tc(*d, A, B, C, bias, R, (cudaStream_t)stream, v1t::no_drop(), impl == V1T_IMPL_BF16X3);
}

extern "C" size_t v1t_matrix_plane_bytes(int64_t rows, int64_t cols) {
  return (rows > 0 && cols > 0) ? v1t::matrix_plane_bytes(rows, cols) : 0;
}

extern "C" int v1t_matrix_planes(const float* X, int64_t ld, int64_t rows, int64_t cols, void* hi, void* lo,
                                 void* stream) {
  V1T_CHECK_ARG(X && hi && rows > 0 && cols > 0 && ld >= cols, "v1t_matrix_planes: bad argument");
  v1t::PlaneOp out;
  return v1t::matrix_planes(X, ld, rows, cols, hi, lo, &out, (cudaStream_t)stream);
}

extern "C" int v1t_gemm_tc_planes(const v1t_gemm_desc* d, const float* A, const float* B, float* C, const float* bias,
                                  const float* R, int impl, const void* a_hi, const void* a_lo, int64_t a_rows,
                                  int64_t a_cols, const void* b_hi, const void* b_lo, int64_t b_rows, int64_t b_cols,
                                  void* stream) {
  using namespace v1t;
  V1T_CHECK_ARG(d && C && (A || a_hi) && (B || b_hi), "v1t_gemm_tc_planes: null argument");
  V1T_CHECK_ARG(impl == V1T_IMPL_BF16X3 || impl == V1T_IMPL_BF16, "v1t_gemm_tc_planes: impl must be BF16X3 or BF16");
  V1T_CHECK_ARG(d->k > 0 && (d->a_k == 1 || d->a_m == 1) && (d->b_k == 1 || d->b_n == 1),
                "v1t_gemm_tc_planes: operands must be contiguous along k or along m/n");
  PlaneOp pa = no_plane(), pb = no_plane();
  if (a_hi) {  // the planes' matrix is A itself (k = columns) or A^T (k = rows)
    V1T_CHECK_ARG(d->a_k == 1 ? (a_rows == d->m && a_cols == d->k) : (a_rows == d->k && a_cols == d->m),
                  "v1t_gemm_tc_planes: A planes do not match the problem");
    pa = PlaneOp{(const uint8_t*)a_hi, (const uint8_t*)a_lo, (int)round_up(a_rows, 32), cdiv(a_cols, 32)};
  }
  if (b_hi) {
    V1T_CHECK_ARG(d->b_k == 1 ? (b_rows == d->n && b_cols == d->k) : (b_rows == d->k && b_cols == d->n),
                  "v1t_gemm_tc_planes: B planes do not match the problem");
    pb = PlaneOp{(const uint8_t*)b_hi, (const uint8_t*)b_lo, (int)round_up(b_rows, 32), cdiv(b_cols, 32)};
  }
  return gemm_tc(*d, A, B, C, bias, R, (cudaStream_t)stream, no_drop(), impl == V1T_IMPL_BF16X3, no_epi(), pa, pb);
}
