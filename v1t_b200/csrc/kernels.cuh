// Internal (non-ABI) declarations shared between the translation units of libv1t_b200.
#pragma once
#include "common.cuh"

namespace v1t {

struct DropSpec {  // inverted dropout; p == 0 -> disabled
  uint64_t seed;
  uint32_t site;
  float p;
};
static inline DropSpec no_drop() { return DropSpec{0, 0, 0.f}; }

// optional fused element-wise stage of the tensor-core GEMM epilogue (MLP of the ViT block, vit.py:144-148)
enum { kEpiNone = 0, kEpiGeluOut = 1, kEpiGeluGrad = 2, kEpiHeadPlanes = 3, kEpiPlanesOut = 4 };
// Destination for operand planes (see PlaneOp below) emitted by the kernel that PRODUCES an activation, so that the
// GEMMs consuming it never convert: hi/lo plane base pointers and the padded row count of the [rows, cols] matrix.
struct PlaneOut {
  uint8_t* hi;
  uint8_t* lo;
  int rows_p;
  // kEpiPlanesOut of a batched problem: batch entry (b1, b2) writes rows b1 * b1_rows + m of column atoms
  // b2 * b2_atoms + n / 32 (e.g. dQ of sample b1, head b2 inside the planes of the head-padded [B*T, 3*H*Dp] gradient)
  int b1_rows, b2_atoms;
};
static inline PlaneOut no_plane_out() { return PlaneOut{nullptr, nullptr, 0, 0, 0}; }

// Destination of kEpiHeadPlanes: the GEMM output [B*T, S * H * AD*32] (S = 3: q | k | v, S = 1: dO; every head padded to AD*32
// columns) leaves the epilogue as the per-(sample, head) attention operand planes of planes.cu
// ([b*H + h][AD atoms][Tq rows][64 B]) instead of fp32 -- no separate conversion pass over qkv.
struct HeadPlanes {
  uint8_t* p[3][2];   // [q, k, v][hi, lo]
  int T, Tq, H, AD;
};

struct EpiOp {
  int kind;           // kEpiGeluOut: aux = gelu(C) * dropout ;  kEpiGeluGrad: C *= gelu'(u) * dropout
                      // kEpiHeadPlanes: alpha * acc -> attention planes `hp` (C may be null)
                      // kEpiPlanesOut: alpha * acc -> matrix planes `pl` only (C null; batched problems allowed)
  float* aux;         // [m, ld] second output (kEpiGeluOut)
  const float* u;     // [m, ld] pre-activation (kEpiGeluGrad)
  int64_t ld;
  DropSpec drop;      // dropout of the activation (element index m * roundup(n,4) + n)
  PlaneOut pl;        // optional planes of the activation-side result (aux for kEpiGeluOut, C for kEpiGeluGrad)
  HeadPlanes hp;
  float* cs_part;     // kEpiGeluGrad: optional [4 * ceil(m/128)][n] partial column sums of C (32-row groups, 0 past m)
};
static inline EpiOp no_epi() { return EpiOp{kEpiNone, nullptr, nullptr, 0, DropSpec{0, 0, 0.f}, no_plane_out(), HeadPlanes{}, nullptr}; }

// Pre-swizzled bf16 hi/lo planes of a row-major matrix X[rows, cols] (planes.cu: matrix_planes): 32-column atoms,
// [catoms][rows_p][64 B], rows 64 B apart, 16-byte chunks XOR-swizzled with ((row >> 1) & 3), zero padded.  A GEMM
// operand given as planes is bulk-copied straight into the UMMA stage (no conversion work in the GEMM): K-major
// when the contraction runs over X's columns, MN-major when it runs over X's rows.
struct PlaneOp {
  const uint8_t* hi;
  const uint8_t* lo;
  int rows_p, catoms;
  int64_t batch_bytes;  // batched problems: bytes between the planes of consecutive (batch1, batch2) matrices (0: unbatched)
  int tile_major;       // MN-major operand whose 64-row k-blocks are contiguous over the atoms (ONE bulk copy per k-block):
                        // 1 = attention planes [64-row tile][catoms atoms][64 rows][64 B] (planes.cu), B operand;
                        // 2 = [128-column M tile][64-row tile][4 atoms][64 rows][64 B], A operand (dS' of attn_bwd2.cu)
};
static inline PlaneOp no_plane() { return PlaneOp{nullptr, nullptr, 0, 0, 0, 0}; }

struct GemmArgs {
  v1t_gemm_desc d;
  DropSpec drop;    // applied to alpha*acc + bias (element index m*roundup(n_cols,4) + n), before the residual add
  EpiOp epi;
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  const float* R;
  int splits;       // split-K factor (blockIdx.z = (b1*batch2 + b2)*splits + split)
  int k_chunk;      // K range per split
  int64_t c_split;  // C element stride between splits (partials)
};

// gemm_fp32.cu
int gemm_fp32(const v1t_gemm_desc& d, const float* A, const float* B, float* C, const float* bias, const float* R,
              cudaStream_t st, DropSpec drop = no_drop());
int gemm_fp32_splitk(const v1t_gemm_desc& d, const float* A, const float* B, float* C, float* partials,
                     size_t partial_bytes, cudaStream_t st);
int reduce_partials(const float* partials, float* out, int parts, int64_t rows, int64_t cols, int64_t ld_out,
                    int accumulate, cudaStream_t st);

// GroupMap: the product was computed with head-padded rows / columns (groups of *_gout holding *_gin real
// entries); the reduction writes only the real ones, compacted, into `out`
struct GroupMap { int row_gin, row_gout, col_gin, col_gout; };
static inline GroupMap no_group_map() { return GroupMap{0, 0, 0, 0}; }
int reduce_partials_ld(const float* partials, float* out, int parts, int64_t rows, int64_t cols, int64_t in_ld,
                       int64_t out_ld, int accumulate, cudaStream_t st, GroupMap gm = no_group_map());

// gemm_tc.cu (tcgen05): same contract; x3 = 1 -> bf16 hi/lo split (3 MMAs), 0 -> plain bf16 operands
int gemm_tc(const v1t_gemm_desc& d, const float* A, const float* B, float* C, const float* bias, const float* R,
            cudaStream_t st, DropSpec drop, int x3, EpiOp epi = no_epi(), PlaneOp pa = no_plane(),
            PlaneOp pb = no_plane());
int gemm_tc_splitk(const v1t_gemm_desc& d, const float* A, const float* B, float* C, float* partials,
                   size_t partial_bytes, cudaStream_t st, int x3, PlaneOp pa = no_plane(), PlaneOp pb = no_plane(),
                   GroupMap gm = no_group_map());

// impl dispatch used by the orchestrator
inline bool gemm_uses_tc(int impl, const v1t_gemm_desc& d) {
  const int64_t work = (int64_t)d.m * d.n * (d.k > 0 ? d.k : 1) * d.batch1 * d.batch2;
  return impl != V1T_IMPL_FP32 && work >= (1ll << 22) && d.k > 0 && (d.a_k == 1 || d.a_m == 1) &&
         (d.b_k == 1 || d.b_n == 1);
}
inline int gemm_any(int impl, const v1t_gemm_desc& d, const float* A, const float* B, float* C, const float* bias,
                    const float* R, cudaStream_t st, DropSpec drop = no_drop(), EpiOp epi = no_epi(),
                    PlaneOp pa = no_plane(), PlaneOp pb = no_plane()) {
  if (!gemm_uses_tc(impl, d)) return gemm_fp32(d, A, B, C, bias, R, st, drop);  // caller applies `epi` separately
  return gemm_tc(d, A, B, C, bias, R, st, drop, impl == V1T_IMPL_BF16X3, epi, pa, pb);
}
inline int gemm_any_splitk(int impl, const v1t_gemm_desc& d, const float* A, const float* B, float* C,
                           float* partials, size_t partial_bytes, cudaStream_t st, PlaneOp pa = no_plane(),
                           PlaneOp pb = no_plane(), GroupMap gm = no_group_map()) {
  const int64_t work = (int64_t)d.m * d.n * (d.k > 0 ? d.k : 1);
  if (impl == V1T_IMPL_FP32 || work < (1ll << 22)) return gemm_fp32_splitk(d, A, B, C, partials, partial_bytes, st);
  return gemm_tc_splitk(d, A, B, C, partials, partial_bytes, st, impl == V1T_IMPL_BF16X3, pa, pb, gm);
}

// planes.cu / attn_tc.cu (fused attention on tcgen05)
size_t plane_bytes(int B, int H, int Tp, int Dp);
// planes of a plain row-major matrix (GEMM operands, see PlaneOp): bytes of ONE plane, and the converter
size_t matrix_plane_bytes(int64_t rows, int64_t cols);
// row_gin / row_gout: every group of row_gin source rows becomes row_gout plane rows (zero padded), e.g. the
// per-head blocks of Wqkv (155 rows) padded to 160 so that the QKV GEMM output is head-aligned.
int matrix_planes(const float* X, int64_t ld, int64_t rows, int64_t cols, void* hi, void* lo, PlaneOp* out,
                  cudaStream_t st, int row_gin = 0, int row_gout = 0, int col_gin = 0, int col_gout = 0);
// ---- materialised attention on tensor cores with plane operands (head dim > 160) -------------------
// per-(sample, head) matrix planes [b*H + h][Dp/32 atoms][Tq rows][64 B] of X[(b*T + t), col0 + h*E + d]
size_t bh_plane_bytes(int B, int H, int Tq, int Dp);
int bh_planes(const float* X, int64_t ld, int col0, int B, int H, int T, int Tq, int E, int Dp, void* hi, void* lo,
              cudaStream_t st);
// bytes of one plane of `mats` T x T probability matrices: [mat][Tq/32 atoms][Tq rows][64 B]
size_t prob_plane_bytes(int64_t mats, int Tq);
// rows of S [mats*T, ld] (scores * scale already applied by the GEMM) -> P = softmax(row) (dropout) as planes; keeps
// nothing in fp32.  row_offset = global row index of S's first row (dropout indexing as softmax_rows)
int softmax_rows_planes(const float* S, int64_t mats, int T, int Tq, int64_t ld, DropSpec dr, int64_t row_offset, void* p_hi,
                        void* p_lo, cudaStream_t st);
// backward: S (recomputed scores) and dPd [mats*T, ld] -> planes of Pd = P * mask and of dS = P * (dPd * mask - delta)
int softmax_bwd_rows_planes(const float* S, const float* dPd, int64_t mats, int T, int Tq, int64_t ld, DropSpec dr,
                            int64_t row_offset, void* pd_hi, void* pd_lo, void* ds_hi, void* ds_lo, cudaStream_t st);
int attn_delta_planes(const PlaneOp& o, const void* do_hi, const void* do_lo, float* delta, int B, int H, int T, int Tp,
                      int AD, cudaStream_t st);
// the same for several matrices in ONE launch (all weights of a block)
struct PlaneJob {
  const float* X;
  int64_t ld, rows, cols, rows_p;  // rows_p is filled in by matrix_planes_batch
  uint8_t *hi, *lo;
  int row_gin, row_gout, col_gin, col_gout;
};
struct PlaneJobs {
  static constexpr int kMax = 8;
  PlaneJob job[kMax];
  int64_t first[kMax + 1];
  int n;
};
int matrix_planes_batch(PlaneJobs& jobs, cudaStream_t st);
// zero rows [T, Tq) of every (sample, head, atom) slab of up to 6 attention planes
// attn_ad > 0: the planes are attention planes with attn_ad atoms per head (slabs = B*H*attn_ad), else matrix planes
int zero_plane_pad_rows(uint8_t* const* planes, int n_planes, int64_t slabs, int Tq, int T, cudaStream_t st,
                        int attn_ad = 0);
// fp32 qkv [B*T, 3*H*E] back from the attention planes (hi + lo), for the attention-map hooks
int planes_to_qkv(const HeadPlanes& hp, int B, int E, float* qkv, cudaStream_t st);
int make_planes(const float* X, int64_t ld, int col0, int B, int H, int T, int Tp, int E, int Dp, void* rm_hi,
                void* rm_lo, void* tr_hi, void* tr_lo, cudaStream_t st);
int attn_delta(const float* O, const float* dO, float* delta, int B, int H, int T, int Tp, int E, int64_t ld,
               cudaStream_t st);
struct AttnFwdArgs {
  const uint8_t *q_hi, *q_lo, *k_hi, *k_lo, *v_hi, *v_lo;  // pre-swizzled bf16 RM planes (rows = tokens; planes.cu)
  float* O;          // [B, T, o_ld]: O[(b*T+t)*o_ld + h*E + d] (may be null when o_pl is given)
  PlaneOut o_pl;     // optional: GEMM-operand planes of the head-padded [B*T, H*Dp] output
  int64_t o_ld;
  float* lse;        // [B*H, Tp] log2-domain log-sum-exp (may be null)
  uint8_t* drop_bits;  // [B*H, Tp, Tp/8] keep bits of the attention dropout (bit k of row q: P[q,k] kept), written
                       // when drop.p > 0 and non-null; the backward reads them instead of re-drawing (attn_drop_bits_bytes)
  int B, H, T, Tp, E, Dp;
  float scale_log2;  // E^-0.5 * log2(e)
  int x3;
  int prec;          // precision-budget experiment (attn_prec_env): bit 0 = P enters P V as its bf16 hi plane only
  // EMIT mode (attn_emit_probs_tc): only the normalised probabilities are produced, from the lse a forward saved
  float* probs;          // [B*H, T, T] fp32, softmax(Q K^T E^-0.5) before dropout
  const float* lse_in;   // [B*H, Tp] base-2 log-sum-exp of that forward
  DropSpec drop;
};
// V1T_ATTN_PREC (default 0 = every contraction with all three bf16x3 terms).  Bits drop ONE cross term of a
// contraction: 1 = the Pd'/dS'/P operand of the accumulating MMAs is hi-only, 2 = dP' without the resident-lo term
// (the only SS-form MMAs of the backward), 4 = dP' from the hi planes alone.  Measured in DESIGN.md 4.2.
int attn_prec_env();
// dV + dK from one recomputation of P' by two-CTA clusters (attn_bwd2.cu) -- the default; V1T_ATTN_BWD=three selects
// the three atomic-free passes (dK | dQ | dV) instead
int attn_bwd_pair_env();
// V1T_ATTN_BWD_GROUP: (b, h) pairs per block-order group of the three-pass backward (default 16; see attn_bwd2.cu)
int attn_bwd_group_env();
// bytes of the keep-bit mask of one attention call (0 rows are never read for padded queries)
static inline size_t attn_drop_bits_bytes(int B, int H, int Tp) { return (size_t)B * H * Tp * (size_t)(Tp / 8); }
int attn_fwd2_tc(const AttnFwdArgs& a, cudaStream_t st);  // Q and P in tensor memory (attn_fwd2.cu)
// softmax(QK^T) [B*H,T,T] alone, for the attention-map hooks: the forward kernel's score pipeline with P = exp2(S c - lse)
// written out row-coalesced (no P V, no second pass): one launch instead of qkv rebuild + GEMM + softmax pass
int attn_emit_probs_tc(const AttnFwdArgs& a, cudaStream_t st);
int attn_fwd_dispatch(const AttnFwdArgs& a, cudaStream_t st);
struct AttnBwdArgs {
  const uint8_t *q_hi, *q_lo, *k_hi, *k_lo, *v_hi, *v_lo, *do_hi, *do_lo;  // RM planes (rows = tokens, K = head dim)
  const float* lse;    // [B*H, Tp] base-2 log-sum-exp saved by the forward
  const float* delta;  // [B*H, Tp] rowsum(dO * O)
  const uint8_t* drop_bits;  // keep bits written by the forward (required when drop.p > 0)
  float* dqkv;         // [B, T, 3*H*E] packed like to_qkv's output: dQ | dK | dV (may be null when dq_pl is given)
  uint8_t *ds_hi, *ds_lo;  // optional scratch: dS' as GEMM operand planes [b*H + h][Tp/128 query tiles][Tp/64 key tiles][4 atoms][64][64 B],
                           // written by the pair kernel; dQ = scale * dS K is then ONE batched plane GEMM (attn_bwd2.cu)
  PlaneOut dq_pl;      // optional: GEMM-operand planes of the head-padded [B*T, 3*H*Dp] gradient (dQ | dK | dV)
  int B, H, T, Tp, E, Dp;
  float scale_log2, scale;
  int x3;
  int prec;            // see attn_prec_env()
  DropSpec drop;
};
int attn_bwd2_tc(const AttnBwdArgs& a, cudaStream_t st);  // resident operands in tensor memory (attn_bwd2.cu)
int attn_dq_gemm_env();
int attn_bwd_dispatch(const AttnBwdArgs& a, cudaStream_t st);
struct AttnPlanes {  // [0] = hi, [1] = lo
  uint8_t *q[2], *k[2], *v[2];    // RM planes (rows = tokens, K = head dim)
  uint8_t *dO[2];                 // backward
  uint8_t *dS[2];                 // backward: dS' planes for the dQ GEMM, [B*H][Tp/32][Tp][64 B] each
  float *lse, *delta;                                 // [B*H, Tp]
  uint8_t* drop_bits;                                 // [B*H, Tp, Tp/8] (standalone entry points)
  size_t total;
};
AttnPlanes carve_attn_planes(void* base, int B, int H, int Tp, int Dp, bool with_backward);

// elementwise.cu
int im2col(const float* img, float* patches, int B, int C, int H, int W, int p, int s, int gh, int gw,
           cudaStream_t st);
int col2im(const float* dpatches, float* dimg, int B, int C, int H, int W, int p, int s, int gh, int gw,
           cudaStream_t st);
int cls_rows(const float* cls, const float* pos, float* x, int B, int T, int E, int ld, cudaStream_t st);
// colsum_out (optional): also out[c] = sum_r dst[r, c] (fixed-order reduction through `partials`)
int dropout_rows(const float* src, float* dst, int64_t rows, int cols, int64_t ld, DropSpec dr, cudaStream_t st,
                 PlaneOut pl = no_plane_out(), float* colsum_out = nullptr, float* partials = nullptr,
                 size_t partial_bytes = 0);
int bmlp_forward(const float* beh, const float* w0, const float* b0, const float* w3, const float* b3, float* hid,
                 float* lat, int B, int bdim, int H, int E, cudaStream_t st);
int tanh_grad(const float* dy, const float* y, float* dz, int64_t n, cudaStream_t st);  // dz = dy*(1-y^2)
size_t bmlp_backward_smem(int B, int H, int E);
int bmlp_backward(const float* dlat, const float* lat, const float* hid, const float* beh, const float* w3, float* dw0,
                  float* db0, float* dw3, float* db3, int B, int bdim, int H, int E, cudaStream_t st);
int ln_forward(const float* x_in, const float* add, int rows_per_batch, float* x_out, const float* gamma,
               const float* beta, float* h, float* stats, int64_t rows, int E, int ld, cudaStream_t st,
               PlaneOut pl = no_plane_out());
int ln_backward(const float* dh, const float* x, const float* stats, const float* gamma, float* dx_accum,
                float* dgamma, float* dbeta, float* partials, size_t partial_bytes, int64_t rows, int E, int ld,
                cudaStream_t st);
int softmax_rows(float* S, int64_t rows, int cols, int64_t ld, DropSpec dr, int64_t row_offset, cudaStream_t st);
int softmax_bwd_rows(float* P, float* dP, int64_t rows, int cols, int64_t ld, DropSpec dr, int64_t row_offset,
                     cudaStream_t st);
int gelu_forward(const float* u, float* g, int64_t rows, int cols, int64_t ld, DropSpec dr, cudaStream_t st);
int gelu_backward(float* dg_inout, const float* u, int64_t rows, int cols, int64_t ld, DropSpec dr,
                  cudaStream_t st);
// out[b, c] = sum_r X[b, r, c]   (X element strides: batch xb, row ld, col 1)
// out[c] = sum over `strips` rows of partials[strip * cols + c] (the finishing pass of colsum, for fused producers)
int colsum_finish(const float* partials, float* out, int cols, int strips, cudaStream_t st);
int colsum(const float* X, float* out, int batch, int64_t rows, int cols, int64_t xb, int64_t ld,
           int64_t out_ld, float* partials, size_t partial_bytes, cudaStream_t st);
// out[t, c] = sum_b X[b, t, c]
int batchsum(const float* X, float* out, int B, int64_t rows, int cols, int64_t xb, int64_t ld, int64_t out_ld,
             cudaStream_t st);
int dropout_mask(float* out, int64_t n, DropSpec dr, cudaStream_t st);

}  // namespace v1t
