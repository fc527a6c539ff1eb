// Host-side orchestration of the ViT core forward / backward (the launch sequence lives in native code, the
// Python boundary only hands over pointers).  Reference: ViTCore.forward vit.py:423-436 and everything it
// calls (Image2Patches :122-129, Transformer.forward :348-362, Attention.mha :267-275, MLP :153-154), plus
// the autograd of that graph.  This file implements V1T_IMPL_FP32 (CUDA-core GEMMs, materialised attention
// probabilities per batch chunk); the tcgen05 paths plug into the same save/scratch layout.
#include <algorithm>
#include <memory>

#include "common.cuh"
#include "kernels.cuh"

namespace v1t {
namespace {

enum Site { kSiteTokens = 0, kSiteAttn = 1, kSiteProj = 2, kSiteMlp1 = 3, kSiteMlp2 = 4 };
inline DropSpec site_drop(const v1t_core_shape& s, int block, Site site) {
  const float p = site == kSiteTokens ? s.p_drop_tokens : s.p_drop_block;
  return DropSpec{s.seed, (uint32_t)(block * 8 + site), p};
}

struct Dims {
  int B, C, H, W, p, s, gh, gw, L, T, Tp, Tq, E, Ep, heads, I, M, Mp, pd, hid, bdim, blocks, impl;
  bool fused;  // fused tcgen05 attention (tensor-core impls, head dim <= 160)
  bool drop;   // block dropout active (the fused attention then keeps its keep-bit mask for the backward)
  bool matpl;  // materialised attention (head dim > 160) on tensor cores: batched GEMMs over bf16 operand planes
  int64_t R;  // B*T rows
};

int make_dims(const v1t_core_shape* sh, Dims& d) {
  V1T_CHECK_ARG(sh, "core: null shape");
  V1T_CHECK_ARG(sh->batch > 0 && sh->in_ch > 0 && sh->in_h > 0 && sh->in_w > 0, "core: bad image shape");
  V1T_CHECK_ARG(sh->patch > 0 && sh->stride >= 1 && sh->stride <= sh->patch, "core: need 1 <= stride <= patch");
  V1T_CHECK_ARG(sh->in_h >= sh->patch && sh->in_w >= sh->patch, "core: image smaller than a patch");
  V1T_CHECK_ARG(sh->emb > 0 && sh->emb <= 512 && sh->heads > 0 && sh->mlp > 0, "core: bad widths (emb <= 512)");
  V1T_CHECK_ARG(sh->blocks > 0 && sh->blocks <= V1T_MAX_BLOCKS, "core: 1..%d blocks", V1T_MAX_BLOCKS);
  V1T_CHECK_ARG(sh->bdim == 0 || sh->bdim == 3 || sh->bdim == 5, "core: bdim must be 0, 3 or 5");
  V1T_CHECK_ARG(sh->p_drop_tokens >= 0.f && sh->p_drop_tokens < 1.f && sh->p_drop_block >= 0.f &&
                    sh->p_drop_block < 1.f, "core: dropout p must be in [0,1)");
  V1T_CHECK_ARG(sh->impl == V1T_IMPL_FP32 || sh->impl == V1T_IMPL_BF16X3 || sh->impl == V1T_IMPL_BF16,
                "core: unknown impl %d", sh->impl);
  d.B = sh->batch; d.C = sh->in_ch; d.H = sh->in_h; d.W = sh->in_w; d.p = sh->patch; d.s = sh->stride;
  d.gh = (d.H - d.p) / d.s + 1;
  d.gw = (d.W - d.p) / d.s + 1;
  d.L = d.gh * d.gw;
  d.T = d.L + 1;
  d.Tp = (int)round_up(d.T, 4);  // row stride of the materialised attention matrices
  d.impl = sh->impl;
  d.E = sh->emb;
  d.Ep = (int)round_up(d.E, 32);
  d.heads = sh->heads;
  d.I = d.heads * d.E;
  d.M = sh->mlp;
  d.Mp = (int)round_up(d.M, 32);
  d.pd = d.C * d.p * d.p;
  d.hid = d.E / 2;
  d.bdim = sh->bdim;
  d.blocks = sh->blocks;
  d.R = (int64_t)d.B * d.T;
  d.Tq = (int)round_up(d.T, 128);  // token padding of the attention operand planes
  d.fused = d.impl != V1T_IMPL_FP32 && d.Ep <= 160 && (int64_t)d.B * d.heads <= 65535;
  d.drop = sh->p_drop_block > 0.f;
  d.matpl = !d.fused && d.impl != V1T_IMPL_FP32 && d.Tq <= 2048;
  return V1T_OK;
}

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base((char*)b) {}
  float* take(int64_t floats) {
    float* p = base ? (float*)(base + off) : nullptr;
    off += (size_t)round_up(floats * (int64_t)sizeof(float), 256);
    return p;
  }
};

struct BlockSaved {
  float *x1, *st1, *qkv, *o, *x2, *st2, *u, *g, *bhid, *blat, *lse;
  uint8_t *qp[2], *kp[2], *vp[2];  // bf16 hi/lo operand planes of Q, K, V (fused attention), reused by the backward
  uint8_t* wpl[4][2];              // GEMM-operand planes of Wqkv, Wproj, W1, W2 (converted once per step)
  uint8_t* gpl[2];                 // operand planes of gelu(u) * dropout (emitted by the MLP GEMM epilogue)
  uint8_t *h1pl[2], *h2pl[2];      // operand planes of the two LayerNorm outputs (kept: no recompute in the backward)
  uint8_t* wqp[2];                 // Wqkv planes with every head's rows padded to Ep (head-aligned QKV GEMM output)
  uint8_t* wpp[2];                 // Wproj planes with every head's columns padded to Ep
  uint8_t* opl[2];                 // operand planes of the head-padded attention output [R, H*Ep]
  uint8_t* dbits;                  // keep bits of the attention dropout [B*H, Tq, Tq/8] (fused attention, p > 0)
};
struct Saved {
  BlockSaved blk[V1T_MAX_BLOCKS];
  size_t total;
};
void carve_block(Carver& c, const Dims& d, BlockSaved& b) {
  b.x1 = c.take(d.R * d.Ep);
  b.st1 = c.take(d.R * 2);
  b.qkv = c.take(d.R * 3 * d.I);
  b.o = c.take(d.R * d.I);
  b.x2 = c.take(d.R * d.Ep);
  b.st2 = c.take(d.R * 2);
  b.u = c.take(d.R * d.Mp);
  b.g = c.take(d.R * d.Mp);  // gelu(u) * dropout, kept for the MLP weight gradient
  b.bhid = c.take((int64_t)d.B * d.hid + 1);
  b.blat = c.take((int64_t)d.B * d.E);
  b.lse = c.take(d.fused ? (int64_t)d.B * d.heads * d.Tq : 1);
  b.dbits = (uint8_t*)c.take(d.fused && d.drop ? (int64_t)(attn_drop_bits_bytes(d.B, d.heads, d.Tq) / sizeof(float)) : 1);
  const int64_t pf = d.fused ? (int64_t)(plane_bytes(d.B, d.heads, d.Tq, d.Ep) / sizeof(float)) : 1;
  for (int i = 0; i < 2; ++i) {
    b.qp[i] = (uint8_t*)c.take(pf);
    b.kp[i] = (uint8_t*)c.take(pf);
    b.vp[i] = (uint8_t*)c.take(pf);
  }
  const int64_t wr[4] = {3 * d.I, d.E, d.M, d.E}, wc[4] = {d.E, d.I, d.E, d.M};
  for (int w = 0; w < 4; ++w)
    for (int i = 0; i < 2; ++i)
      b.wpl[w][i] = (uint8_t*)c.take(d.impl != V1T_IMPL_FP32 ? (int64_t)(matrix_plane_bytes(wr[w], wc[w]) / sizeof(float)) : 1);
  for (int i = 0; i < 2; ++i) {
    b.gpl[i] = (uint8_t*)c.take(d.impl != V1T_IMPL_FP32 ? (int64_t)(matrix_plane_bytes(d.R, d.M) / sizeof(float)) : 1);
    b.h1pl[i] = (uint8_t*)c.take(d.impl != V1T_IMPL_FP32 ? (int64_t)(matrix_plane_bytes(d.R, d.E) / sizeof(float)) : 1);
    b.h2pl[i] = (uint8_t*)c.take(d.impl != V1T_IMPL_FP32 ? (int64_t)(matrix_plane_bytes(d.R, d.E) / sizeof(float)) : 1);
    b.wqp[i] = (uint8_t*)c.take(d.fused ? (int64_t)(matrix_plane_bytes(3 * d.heads * d.Ep, d.E) / sizeof(float)) : 1);
    b.wpp[i] = (uint8_t*)c.take(d.fused ? (int64_t)(matrix_plane_bytes(d.E, d.heads * d.Ep) / sizeof(float)) : 1);
    b.opl[i] = (uint8_t*)c.take(d.fused ? (int64_t)(matrix_plane_bytes(d.R, d.heads * d.Ep) / sizeof(float)) : 1);
  }
}
// the QKV GEMM of the fused-attention path: output columns (q|k|v, head, Ep) written straight into attention planes
v1t_gemm_desc qkv_planes_desc(const Dims& d) {
  v1t_gemm_desc g{};
  g.m = (int)d.R; g.n = 3 * d.heads * d.Ep; g.k = d.E; g.batch1 = 1; g.batch2 = 1; g.alpha = 1.f;
  g.a_m = d.Ep; g.a_k = 1; g.b_k = 1; g.b_n = d.E;
  return g;
}
bool qkv_to_planes(const Dims& d) { return d.fused && gemm_uses_tc(d.impl, qkv_planes_desc(d)); }
HeadPlanes head_planes(const Dims& d, const BlockSaved& S) {
  HeadPlanes hp{};
  const bool x3 = d.impl == V1T_IMPL_BF16X3;
  hp.p[0][0] = S.qp[0]; hp.p[1][0] = S.kp[0]; hp.p[2][0] = S.vp[0];
  hp.p[0][1] = x3 ? S.qp[1] : nullptr; hp.p[1][1] = x3 ? S.kp[1] : nullptr; hp.p[2][1] = x3 ? S.vp[1] : nullptr;
  hp.T = d.T; hp.Tq = d.Tq; hp.H = d.heads; hp.AD = d.Ep / 32;
  return hp;
}
enum { kWqkv = 0, kWproj = 1, kW1 = 2, kW2 = 3 };
// activation planes: the [R, cols] matrix as GEMM operand / as destination of the kernel producing it
PlaneOp act_plane(const Dims& d, uint8_t* const (&pl)[2], int64_t cols) {
  if (d.impl == V1T_IMPL_FP32) return no_plane();
  return PlaneOp{pl[0], d.impl == V1T_IMPL_BF16X3 ? pl[1] : nullptr, (int)round_up(d.R, 32), cdiv(cols, 32)};
}
PlaneOut act_plane_out(const Dims& d, uint8_t* const (&pl)[2]) {
  if (d.impl == V1T_IMPL_FP32) return no_plane_out();
  return PlaneOut{pl[0], d.impl == V1T_IMPL_BF16X3 ? pl[1] : nullptr, (int)round_up(d.R, 32)};
}
// plane operand of weight `w` ([rows, cols] = torch Linear [out, in]); none in fp32 mode
PlaneOp wplane(const Dims& d, const BlockSaved& S, int w) {
  if (d.impl == V1T_IMPL_FP32) return no_plane();
  const int64_t wr[4] = {3 * d.I, d.E, d.M, d.E}, wc[4] = {d.E, d.I, d.E, d.M};
  return PlaneOp{S.wpl[w][0], d.impl == V1T_IMPL_BF16X3 ? S.wpl[w][1] : nullptr, (int)round_up(wr[w], 32), cdiv(wc[w], 32)};
}
Saved carve_saved(const Dims& d, void* base) {
  Carver c(base);
  Saved s;
  for (int i = 0; i < d.blocks; ++i) carve_block(c, d, s.blk[i]);
  s.total = c.off;
  return s;
}

constexpr size_t kPartialBytes = 64u << 20;

struct Scratch {
  BlockSaved tmp;    // one block of "saved" space for inference (keep_for_backward == 0)
  float *h, *dh, *g, *dqkv, *dO, *patches, *P1, *P2, *partials, *dlat, *dz3, *dhid, *dpos;
  uint8_t *dpl[2], *dupl[2];           // operand planes of dropout(dx) and of du (backward)
  uint8_t* dqpl[2];                    // operand planes of the head-padded dqkv (from the attention backward)
  // materialised attention on plane operands (Dims::matpl), one chunk of samples at a time: per-(sample, head) planes
  // of Q, K, V, dO [chunk*H][Ep/32][Tq][64 B] and of the T x T matrices Pd, dS [chunk*H][Tq/32][Tq][64 B]
  uint8_t *bhq[2], *bhk[2], *bhv[2], *bhdo[2], *pdpl[2], *dspl[2];
  AttnPlanes planes;
  int chunk;         // attention batch chunk
  size_t total;
};
int attn_chunk(const Dims& d) {
  const int64_t per = (int64_t)d.heads * d.T * d.Tp * (int64_t)sizeof(float);
  int64_t c = (2ll << 30) / std::max<int64_t>(per, 1);  // two fp32 T x T buffers of <= 2 GiB each
  c = std::max<int64_t>(1, std::min<int64_t>(c, d.B));
  // keep grid.z = chunk*heads within limits
  while (c * d.heads > 65535) --c;
  return (int)c;
}
Scratch carve_scratch(const Dims& d, void* base) {
  Carver c(base);
  Scratch s;
  carve_block(c, d, s.tmp);
  s.chunk = attn_chunk(d);
  s.h = c.take(d.R * d.Ep);
  s.dh = c.take(d.R * d.Ep);
  s.g = c.take(d.R * d.Mp);
  s.dqkv = c.take(d.R * 3 * d.I);
  s.dO = c.take(d.R * d.I);
  s.patches = c.take((int64_t)d.B * d.L * d.pd);
  // materialised attention matrices only for the fp32 path; the fused path keeps bf16 operand planes instead
  s.P1 = c.take(d.fused ? 1 : (int64_t)s.chunk * d.heads * d.T * d.Tp);
  s.P2 = c.take(d.fused ? 1 : (int64_t)s.chunk * d.heads * d.T * d.Tp);
  if (d.fused) {
    const size_t pb = carve_attn_planes(nullptr, d.B, d.heads, d.Tq, d.Ep, true).total;
    float* pbase = c.take((int64_t)(pb / sizeof(float)) + 256);
    s.planes = carve_attn_planes(pbase, d.B, d.heads, d.Tq, d.Ep, true);
  } else {
    s.planes = AttnPlanes{};
  }
  for (int i = 0; i < 2; ++i) {
    const bool tc = d.impl != V1T_IMPL_FP32;
    s.dpl[i] = (uint8_t*)c.take(tc ? (int64_t)(matrix_plane_bytes(d.R, d.E) / sizeof(float)) : 1);
    s.dupl[i] = (uint8_t*)c.take(tc ? (int64_t)(matrix_plane_bytes(d.R, d.M) / sizeof(float)) : 1);
    s.dqpl[i] = (uint8_t*)c.take(d.fused ? (int64_t)(matrix_plane_bytes(d.R, 3 * d.heads * d.Ep) / sizeof(float)) : 1);
    const bool need = d.matpl && (i == 0 || d.impl == V1T_IMPL_BF16X3);
    const int64_t bhf = need ? (int64_t)(bh_plane_bytes(s.chunk, d.heads, d.Tq, d.Ep) / sizeof(float)) : 1;
    const int64_t ppf = need ? (int64_t)(prob_plane_bytes((int64_t)s.chunk * d.heads, d.Tq) / sizeof(float)) : 1;
    s.bhq[i] = (uint8_t*)c.take(bhf);
    s.bhk[i] = (uint8_t*)c.take(bhf);
    s.bhv[i] = (uint8_t*)c.take(bhf);
    s.bhdo[i] = (uint8_t*)c.take(bhf);
    s.pdpl[i] = (uint8_t*)c.take(ppf);
    s.dspl[i] = (uint8_t*)c.take(ppf);
  }
  s.partials = c.take(kPartialBytes / sizeof(float));
  s.dlat = c.take((int64_t)d.B * d.E);
  s.dz3 = c.take((int64_t)d.B * d.E);
  s.dhid = c.take((int64_t)d.B * d.hid + 1);
  s.dpos = c.take((int64_t)d.T * d.E);
  s.total = c.off;
  return s;
}

v1t_gemm_desc gd(int m, int n, int k) {
  v1t_gemm_desc g{};
  g.m = m; g.n = n; g.k = k; g.batch1 = 1; g.batch2 = 1; g.alpha = 1.f;
  return g;
}

// ---- materialised attention over plane operands (Dims::matpl) ----
inline PlaneOp bh_op(const Dims& d, uint8_t* const p[2]) {  // one (sample, head) matrix [T, E] per batch entry
  return PlaneOp{p[0], d.impl == V1T_IMPL_BF16X3 ? p[1] : nullptr, d.Tq, d.Ep / 32, (int64_t)(d.Ep / 32) * d.Tq * 64};
}
inline PlaneOp prob_op(const Dims& d, uint8_t* const p[2]) {  // one T x T matrix per batch entry
  return PlaneOp{p[0], d.impl == V1T_IMPL_BF16X3 ? p[1] : nullptr, d.Tq, d.Tq / 32, (int64_t)(d.Tq / 32) * d.Tq * 64};
}
// batched [bc, heads] problem with both operands from planes; a_mn / b_mn: operand contracted over its plane ROWS
inline v1t_gemm_desc bh_gemm(const Dims& d, int bc, int m, int n, int k, bool a_mn, bool b_mn) {
  v1t_gemm_desc g = gd(m, n, k);
  g.batch1 = bc; g.batch2 = d.heads;
  g.a_m = a_mn ? 1 : d.Tq; g.a_k = a_mn ? d.Tq : 1;
  g.b_n = b_mn ? 1 : d.Tq; g.b_k = b_mn ? d.Tq : 1;
  return g;
}
int bh_planes_of(const Dims& d, const float* X, int64_t ld, int col0, int bc, uint8_t* const p[2], cudaStream_t st) {
  return bh_planes(X, ld, col0, bc, d.heads, d.T, d.Tq, d.E, d.Ep, p[0], d.impl == V1T_IMPL_BF16X3 ? p[1] : nullptr, st);
}
// scores S[b,h] = scale * Q K^T for samples [b0, b0+bc) from the Q / K planes of the chunk
int scores_from_planes(const Dims& d, const Scratch& sc, int bc, float* S, cudaStream_t st) {
  v1t_gemm_desc g = bh_gemm(d, bc, d.T, d.T, d.E, false, false);
  g.c_m = d.Tp; g.c_b1 = (int64_t)d.heads * d.T * d.Tp; g.c_b2 = (int64_t)d.T * d.Tp;
  g.alpha = 1.0f / sqrtf((float)d.E);  // emb_dim ** -0.5 (vit.py:234)
  return gemm_any(d.impl, g, nullptr, nullptr, S, nullptr, nullptr, st, no_drop(), no_epi(), bh_op(d, sc.bhq), bh_op(d, sc.bhk));
}

// P[b,h] = softmax(scale * Q K^T) (optionally with dropout) for samples [b0, b0+bc) into P1
int attention_probs(const Dims& d, const float* qkv, int b0, int bc, float* P, int64_t pld, DropSpec dr,
                    cudaStream_t st) {
  const int64_t ld = 3 * d.I;
  v1t_gemm_desc g = gd(d.T, d.T, d.E);
  g.batch1 = bc; g.batch2 = d.heads;
  g.a_m = ld; g.a_k = 1; g.a_b1 = d.T * ld; g.a_b2 = d.E;
  g.b_k = 1; g.b_n = ld; g.b_b1 = d.T * ld; g.b_b2 = d.E;
  g.c_m = pld; g.c_b1 = (int64_t)d.heads * d.T * pld; g.c_b2 = (int64_t)d.T * pld;
  g.alpha = 1.0f / sqrtf((float)d.E);  // emb_dim ** -0.5 (vit.py:234)
  const float* q = qkv + (int64_t)b0 * d.T * ld;
  V1T_TRY(gemm_any(d.impl, g, q, q + d.I, P, nullptr, nullptr, st));
  return softmax_rows(P, (int64_t)bc * d.heads * d.T, d.T, pld, dr, (int64_t)b0 * d.heads * d.T, st);
}

}  // namespace
}  // namespace v1t

using namespace v1t;

extern "C" int v1t_core_dims_of(const v1t_core_shape* shape, v1t_core_dims* out) {
  Dims d;
  V1T_TRY(make_dims(shape, d));
  V1T_CHECK_ARG(out, "core_dims_of: null out");
  out->gh = d.gh; out->gw = d.gw; out->tokens = d.T; out->emb_ld = d.Ep; out->inner = d.I; out->mlp_ld = d.Mp;
  out->patch_dim = d.pd; out->hid = d.hid;
  out->attn_path = d.fused ? V1T_ATTN_FUSED : V1T_ATTN_MATERIALISED;
  return V1T_OK;
}

extern "C" size_t v1t_core_saved_bytes(const v1t_core_shape* shape) {
  Dims d;
  if (make_dims(shape, d) != V1T_OK) return 0;
  return carve_saved(d, nullptr).total;
}
extern "C" size_t v1t_core_scratch_bytes(const v1t_core_shape* shape) {
  Dims d;
  if (make_dims(shape, d) != V1T_OK) return 0;
  return carve_scratch(d, nullptr).total;
}

extern "C" int v1t_core_forward(const v1t_core_shape* shape, const v1t_core_ptrs* P, const float* images,
                                const float* behaviors, float* x, void* saved_mem, void* scratch_mem,
                                int keep, void* stream) {
  Dims d;
  V1T_TRY(make_dims(shape, d));
  V1T_CHECK_ARG(P && images && x && scratch_mem, "core_forward: null argument");
  V1T_CHECK_ARG(!keep || saved_mem, "core_forward: keep_for_backward needs the saved buffer");
  V1T_CHECK_ARG(d.bdim == 0 || behaviors, "core_forward: behaviors missing");
  V1T_CHECK_ARG(P->cls && P->pos && P->wpe && P->bpe, "core_forward: patch-embedding parameter missing");
  cudaStream_t st = (cudaStream_t)stream;
  Saved sv = carve_saved(d, keep ? saved_mem : nullptr);
  Scratch sc = carve_scratch(d, scratch_mem);

  // ---- Image2Patches (vit.py:122-129): unfold -> Linear -> [CLS; tokens] + pos -> dropout
  {
  ProfScope prof(V1T_PHASE_PATCH, st);
  V1T_CUDA(cudaMemsetAsync(x, 0, sizeof(float) * d.R * d.Ep, st));
  V1T_TRY(im2col(images, sc.patches, d.B, d.C, d.H, d.W, d.p, d.s, d.gh, d.gw, st));
  {
    v1t_gemm_desc g = gd(d.L, d.E, d.pd);
    g.batch1 = d.B;
    g.a_m = d.pd; g.a_k = 1; g.a_b1 = (int64_t)d.L * d.pd;
    g.b_k = 1; g.b_n = d.pd;                       // Wpe [E, pd]
    g.c_m = d.Ep; g.c_b1 = (int64_t)d.T * d.Ep;
    g.r_m = d.E; g.r_b1 = 0;                       // + pos[1 + l]
    V1T_TRY(gemm_any(d.impl, g, sc.patches, P->wpe, x + d.Ep, P->bpe, P->pos + d.E, st));
  }
  V1T_TRY(cls_rows(P->cls, P->pos, x, d.B, d.T, d.E, d.Ep, st));
  if (shape->p_drop_tokens > 0.f) V1T_TRY(dropout_rows(x, x, d.R, d.E, d.Ep, site_drop(*shape, 0, kSiteTokens), st));
  }

  for (int i = 0; i < d.blocks; ++i) {
    const v1t_block_ptrs& W = P->blk[i];
    const BlockSaved& S = keep ? sv.blk[i] : sc.tmp;
    V1T_CHECK_ARG(W.ln1_w && W.ln1_b && W.wqkv && W.wproj && W.ln2_w && W.ln2_b && W.w1 && W.w2,
                  "core_forward: block %d parameter missing", i);
    // ---- behaviour latent added to every token, persists in the residual stream (vit.py:355-359)
    const float* lat = nullptr;
    {
    ProfScope prof(V1T_PHASE_LN_QKV, st);
    if (d.bdim > 0) {
      V1T_CHECK_ARG(W.bw0 && W.bw3, "core_forward: block %d b-mlp weights missing", i);
      V1T_TRY(bmlp_forward(behaviors, W.bw0, W.bb0, W.bw3, W.bb3, S.bhid, S.blat, d.B, d.bdim, d.hid, d.E, st));
      lat = S.blat;
    }
    if (d.impl != V1T_IMPL_FP32) {  // weights -> bf16 hi/lo operand planes, shared by the forward and backward GEMMs
      const bool x3 = d.impl == V1T_IMPL_BF16X3;
      PlaneJobs jobs{};
      auto add = [&](const float* X, int64_t ld, int64_t rows, int64_t cols, uint8_t* const (&pl)[2], int rgi = 0,
                     int rgo = 0, int cgi = 0, int cgo = 0) {
        jobs.job[jobs.n++] = PlaneJob{X, ld, rows, cols, 0, pl[0], x3 ? pl[1] : nullptr, rgi, rgo, cgi, cgo};
      };
      add(W.wqkv, d.E, 3 * d.I, d.E, S.wpl[kWqkv]);
      add(W.wproj, d.I, d.E, d.I, S.wpl[kWproj]);
      add(W.w1, d.E, d.M, d.E, S.wpl[kW1]);
      add(W.w2, d.M, d.E, d.M, S.wpl[kW2]);
      if (qkv_to_planes(d)) {  // head-padded variants for the attention branch
        add(W.wqkv, d.E, 3 * d.I, d.E, S.wqp, d.E, d.Ep);
        add(W.wproj, d.I, d.E, d.I, S.wpp, 0, 0, d.E, d.Ep);
      }
      V1T_TRY(matrix_planes_batch(jobs, st));
    }
    // ---- Attention.mha (vit.py:267-275)
    // the normalised rows leave as operand planes; the fp32 copy is written only when a consumer reads it
    V1T_TRY(ln_forward(x, lat, d.T, S.x1, W.ln1_w, W.ln1_b, qkv_to_planes(d) ? nullptr : sc.h, S.st1, d.R, d.E, d.Ep, st,
                       act_plane_out(d, S.h1pl)));
    V1T_CHECK_ARG(d.R <= INT32_MAX, "core_forward: too many rows");
    if (qkv_to_planes(d)) {
      // q, k, v leave the GEMM epilogue as the attention kernels' operand planes (fp32 qkv is never materialised;
      // v1t_attention_probs rebuilds it from the planes for the attention-map hooks)
      const bool x3 = d.impl == V1T_IMPL_BF16X3;
      EpiOp epi = no_epi();
      epi.kind = kEpiHeadPlanes;
      epi.hp = head_planes(d, S);
      uint8_t* pads[6] = {S.qp[0], S.kp[0], S.vp[0], S.qp[1], S.kp[1], S.vp[1]};
      V1T_TRY(zero_plane_pad_rows(pads, x3 ? 6 : 3, (int64_t)d.B * d.heads * (d.Ep / 32), d.Tq, d.T, st, d.Ep / 32));
      const PlaneOp wq{S.wqp[0], x3 ? S.wqp[1] : nullptr, (int)round_up(3 * d.heads * d.Ep, 32), cdiv(d.E, 32)};
      V1T_TRY(gemm_any(d.impl, qkv_planes_desc(d), sc.h, nullptr, nullptr, nullptr, nullptr, st, no_drop(), epi,
                       act_plane(d, S.h1pl, d.E), wq));
    } else {
      v1t_gemm_desc g = gd((int)d.R, 3 * d.I, d.E);
      g.a_m = d.Ep; g.a_k = 1; g.b_k = 1; g.b_n = d.E; g.c_m = 3 * d.I;
      V1T_TRY(gemm_any(d.impl, g, sc.h, W.wqkv, S.qkv, nullptr, nullptr, st, no_drop(), no_epi(),
                       act_plane(d, S.h1pl, d.E), wplane(d, S, kWqkv)));
    }
    }
    if (d.fused) {  // tcgen05 fused attention: qkv -> bf16 operand planes -> O, lse (nothing T x T in HBM)
      ProfScope prof(V1T_PHASE_ATTN_FWD, st);
      const int x3 = d.impl == V1T_IMPL_BF16X3;
      const AttnPlanes& pl = sc.planes;
      // Q, K, V planes live in the per-block saved area (the backward reuses them)
      if (!qkv_to_planes(d)) {
        V1T_TRY(make_planes(S.qkv, 3 * d.I, 0, d.B, d.heads, d.T, d.Tq, d.E, d.Ep, S.qp[0], x3 ? S.qp[1] : nullptr,
                            nullptr, nullptr, st));
        V1T_TRY(make_planes(S.qkv, 3 * d.I, d.I, d.B, d.heads, d.T, d.Tq, d.E, d.Ep, S.kp[0], x3 ? S.kp[1] : nullptr,
                            nullptr, nullptr, st));
        V1T_TRY(make_planes(S.qkv, 3 * d.I, 2 * d.I, d.B, d.heads, d.T, d.Tq, d.E, d.Ep, S.vp[0], x3 ? S.vp[1] : nullptr,
                            nullptr, nullptr, st));
      }
      AttnFwdArgs fa{};
      fa.q_hi = S.qp[0]; fa.q_lo = S.qp[1]; fa.k_hi = S.kp[0]; fa.k_lo = S.kp[1]; fa.v_hi = S.vp[0]; fa.v_lo = S.vp[1];
      fa.O = S.o; fa.o_ld = d.I; fa.lse = S.lse;
      fa.drop_bits = d.drop ? S.dbits : nullptr;
      if (qkv_to_planes(d)) {  // O leaves the kernel only as operand planes of the head-padded [R, H*Ep] matrix
        fa.O = nullptr;
        fa.o_pl = act_plane_out(d, S.opl);
        uint8_t* pads[2] = {S.opl[0], S.opl[1]};  // rows past B*T (contracted by the Wproj weight gradient)
        V1T_TRY(zero_plane_pad_rows(pads, x3 ? 2 : 1, d.heads * (d.Ep / 32), (int)round_up(d.R, 32), (int)d.R, st));
      }
      fa.B = d.B; fa.H = d.heads; fa.T = d.T; fa.Tp = d.Tq; fa.E = d.E; fa.Dp = d.Ep;
      fa.scale_log2 = (1.0f / sqrtf((float)d.E)) * 1.4426950408889634f;
      fa.x3 = x3;
      fa.prec = attn_prec_env();
      fa.drop = site_drop(*shape, i, kSiteAttn);
      {
        ProfScope kprof(V1T_PHASE_ATTN_FWD_KERNEL, st);
        V1T_TRY(attn_fwd_dispatch(fa, st));
      }
    }
    for (int b0 = 0; d.matpl && b0 < d.B; b0 += sc.chunk) {
      // head dim > 160: S and P go through HBM, but every GEMM reads ready-made bf16 operand planes (no conversion
      // work in the GEMMs) and the softmax writes the probabilities as planes
      ProfScope prof(V1T_PHASE_ATTN_FWD, st);
      const int bc = std::min(sc.chunk, d.B - b0);
      const float* q = S.qkv + (int64_t)b0 * d.T * 3 * d.I;
      V1T_TRY(bh_planes_of(d, q, 3 * d.I, 0, bc, sc.bhq, st));
      V1T_TRY(bh_planes_of(d, q, 3 * d.I, d.I, bc, sc.bhk, st));
      V1T_TRY(bh_planes_of(d, q, 3 * d.I, 2 * d.I, bc, sc.bhv, st));
      V1T_TRY(scores_from_planes(d, sc, bc, sc.P1, st));
      V1T_TRY(softmax_rows_planes(sc.P1, (int64_t)bc * d.heads, d.T, d.Tq, d.Tp, site_drop(*shape, i, kSiteAttn),
                                  (int64_t)b0 * d.heads * d.T, sc.pdpl[0], d.impl == V1T_IMPL_BF16X3 ? sc.pdpl[1] : nullptr, st));
      v1t_gemm_desc g = bh_gemm(d, bc, d.T, d.E, d.T, false, true);  // O = Pd V
      g.c_m = d.I; g.c_b1 = (int64_t)d.T * d.I; g.c_b2 = d.E;
      V1T_TRY(gemm_any(d.impl, g, nullptr, nullptr, S.o + (int64_t)b0 * d.T * d.I, nullptr, nullptr, st, no_drop(), no_epi(),
                       prob_op(d, sc.pdpl), bh_op(d, sc.bhv)));
    }
    for (int b0 = 0; !d.fused && !d.matpl && b0 < d.B; b0 += sc.chunk) {
      ProfScope prof(V1T_PHASE_ATTN_FWD, st);
      const int bc = std::min(sc.chunk, d.B - b0);
      V1T_TRY(attention_probs(d, S.qkv, b0, bc, sc.P1, d.Tp, site_drop(*shape, i, kSiteAttn), st));
      v1t_gemm_desc g = gd(d.T, d.E, d.T);  // O = P V
      g.batch1 = bc; g.batch2 = d.heads;
      g.a_m = d.Tp; g.a_k = 1; g.a_b1 = (int64_t)d.heads * d.T * d.Tp; g.a_b2 = (int64_t)d.T * d.Tp;
      g.b_k = 3 * d.I; g.b_n = 1; g.b_b1 = (int64_t)d.T * 3 * d.I; g.b_b2 = d.E;
      g.c_m = d.I; g.c_b1 = (int64_t)d.T * d.I; g.c_b2 = d.E;
      V1T_TRY(gemm_any(d.impl, g, sc.P1, S.qkv + (int64_t)b0 * d.T * 3 * d.I + 2 * d.I, S.o + (int64_t)b0 * d.T * d.I,
                        nullptr, nullptr, st));
    }
    {  // x2 = x1 + dropout(o Wproj^T + b)
      ProfScope prof(V1T_PHASE_PROJ, st);
      if (qkv_to_planes(d)) {  // head-padded contraction (K = H*Ep), both operands from planes
        const int KP = d.heads * d.Ep;
        v1t_gemm_desc g = gd((int)d.R, d.E, KP);
        g.a_m = KP; g.a_k = 1; g.b_k = 1; g.b_n = KP; g.c_m = d.Ep; g.r_m = d.Ep;
        const PlaneOp wp{S.wpp[0], d.impl == V1T_IMPL_BF16X3 ? S.wpp[1] : nullptr, (int)round_up(d.E, 32), cdiv(KP, 32)};
        V1T_TRY(gemm_any(d.impl, g, nullptr, nullptr, S.x2, W.bproj, S.x1, st, site_drop(*shape, i, kSiteProj), no_epi(),
                         act_plane(d, S.opl, KP), wp));
      } else {
        v1t_gemm_desc g = gd((int)d.R, d.E, d.I);
        g.a_m = d.I; g.a_k = 1; g.b_k = 1; g.b_n = d.I; g.c_m = d.Ep; g.r_m = d.Ep;
        V1T_TRY(gemm_any(d.impl, g, S.o, W.wproj, S.x2, W.bproj, S.x1, st, site_drop(*shape, i, kSiteProj), no_epi(),
                         no_plane(), wplane(d, S, kWproj)));
      }
    }
    // ---- MLP (vit.py:143-150)
    ProfScope prof_mlp(V1T_PHASE_MLP, st);
    {
      v1t_gemm_desc g = gd((int)d.R, d.M, d.E);
      g.a_m = d.Ep; g.a_k = 1; g.b_k = 1; g.b_n = d.E; g.c_m = d.Mp;
      V1T_TRY(ln_forward(S.x2, nullptr, d.T, nullptr, W.ln2_w, W.ln2_b, gemm_uses_tc(d.impl, g) ? nullptr : sc.h, S.st2, d.R, d.E,
                         d.Ep, st, act_plane_out(d, S.h2pl)));
      if (gemm_uses_tc(d.impl, g)) {  // u and gelu(u)*dropout from one epilogue
        // gelu(u) * dropout leaves the epilogue as operand planes only (consumers: W2 GEMM and its weight gradient)
        const EpiOp act{kEpiGeluOut, nullptr, nullptr, d.Mp, site_drop(*shape, i, kSiteMlp1), act_plane_out(d, S.gpl)};
        V1T_TRY(gemm_any(d.impl, g, sc.h, W.w1, S.u, W.b1, nullptr, st, no_drop(), act, act_plane(d, S.h2pl, d.E),
                         wplane(d, S, kW1)));
      } else {
        V1T_TRY(gemm_any(d.impl, g, sc.h, W.w1, S.u, W.b1, nullptr, st));
        V1T_TRY(gelu_forward(S.u, S.g, d.R, d.M, d.Mp, site_drop(*shape, i, kSiteMlp1), st));
      }
    }
    {
      v1t_gemm_desc g = gd((int)d.R, d.E, d.M);
      g.a_m = d.Mp; g.a_k = 1; g.b_k = 1; g.b_n = d.M; g.c_m = d.Ep; g.r_m = d.Ep;
      const bool gp = gemm_uses_tc(d.impl, g);  // then the MLP GEMM above emitted planes instead of fp32 g
      V1T_TRY(gemm_any(d.impl, g, S.g, W.w2, x, W.b2, S.x2, st, site_drop(*shape, i, kSiteMlp2), no_epi(),
                       gp ? act_plane(d, S.gpl, d.M) : no_plane(), wplane(d, S, kW2)));
    }
  }
  return V1T_OK;
}

extern "C" int v1t_core_backward(const v1t_core_shape* shape, const v1t_core_ptrs* P, const float* images,
                                 const float* behaviors, float* dx, const void* saved_mem, void* scratch_mem,
                                 const v1t_core_ptrs* G, float* d_images, void* stream) {
  Dims d;
  V1T_TRY(make_dims(shape, d));
  V1T_CHECK_ARG(P && images && dx && saved_mem && scratch_mem && G, "core_backward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  Saved sv = carve_saved(d, const_cast<void*>(saved_mem));
  Scratch sc = carve_scratch(d, scratch_mem);
  const float scale = 1.0f / sqrtf((float)d.E);
  const int R = (int)d.R;
  const bool drop = shape->p_drop_block > 0.f;

  for (int i = d.blocks - 1; i >= 0; --i) {
    const v1t_block_ptrs& W = P->blk[i];
    const v1t_block_ptrs& GW = G->blk[i];
    const BlockSaved& S = sv.blk[i];
    // ================= MLP branch: x_out = x2 + drop(g W2^T + b2) =================
    std::unique_ptr<ProfScope> lin1(new ProfScope(V1T_PHASE_LINEAR_BWD, st));
    const float* dm = dx;
    PlaneOp dmp = no_plane();  // operand planes of dm (emitted by the dropout kernel)
    if (drop) {
      // dm = dropout(dx) as fp32 + operand planes, and db2 = colsum(dm) from the same pass
      V1T_TRY(dropout_rows(dx, sc.dh, d.R, d.E, d.Ep, site_drop(*shape, i, kSiteMlp2), st, act_plane_out(d, sc.dpl), GW.b2,
                           sc.partials, kPartialBytes));
      dm = sc.dh;
      dmp = act_plane(d, sc.dpl, d.E);
    }
    // whether the forward MLP GEMM ran on the tensor cores (then gelu(u)*dropout exists as planes only)
    const bool mlp_tc = [&] { v1t_gemm_desc g = gd(R, d.M, d.E); g.a_k = 1; g.b_k = 1; return gemm_uses_tc(d.impl, g); }();
    if (GW.w2) {  // dW2[e,m] = sum_r dm[r,e] g[r,m]
      v1t_gemm_desc g = gd(d.E, d.M, R);
      g.a_m = 1; g.a_k = d.Ep; g.b_k = d.Mp; g.b_n = 1; g.c_m = d.M;
      V1T_TRY(gemm_any_splitk(d.impl, g, dm, S.g, GW.w2, sc.partials, kPartialBytes, st, dmp,
                              mlp_tc ? act_plane(d, S.gpl, d.M) : no_plane()));
    }
    if (GW.b2 && !drop) V1T_TRY(colsum(dm, GW.b2, 1, d.R, d.E, 0, d.Ep, 0, sc.partials, kPartialBytes, st));
    {  // dg[r,m] = sum_e dm[r,e] W2[e,m]   -> sc.g (g no longer needed)
      v1t_gemm_desc g = gd(R, d.M, d.E);
      g.a_m = d.Ep; g.a_k = 1; g.b_k = d.M; g.b_n = 1; g.c_m = d.Mp;
      if (gemm_uses_tc(d.impl, g)) {  // du = (dm W2) * gelu'(u) * dropout in the epilogue
        // du leaves the epilogue as operand planes (for dW1 and dh2) plus 32-row partial column sums (db1): no fp32 du
        EpiOp act{kEpiGeluGrad, nullptr, S.u, d.Mp, site_drop(*shape, i, kSiteMlp1), act_plane_out(d, sc.dupl)};
        act.cs_part = GW.b1 ? sc.partials : nullptr;
        V1T_TRY(gemm_any(d.impl, g, dm, W.w2, nullptr, nullptr, nullptr, st, no_drop(), act, dmp, wplane(d, S, kW2)));
        if (GW.b1) V1T_TRY(colsum_finish(sc.partials, GW.b1, d.M, 4 * cdiv(R, 128), st));
      } else {
        V1T_TRY(gemm_any(d.impl, g, dm, W.w2, sc.g, nullptr, nullptr, st));
        V1T_TRY(gelu_backward(sc.g, S.u, d.R, d.M, d.Mp, site_drop(*shape, i, kSiteMlp1), st));  // du in sc.g
      }
    }
    // the weight gradients read the LayerNorm outputs from the planes the forward kept; only the CUDA-core path
    // (fp32 impl or tiny problems) recomputes them in fp32
    auto wgrad_tc = [&](int64_t m, int64_t n) { return d.impl != V1T_IMPL_FP32 && m * n * (int64_t)R >= (1ll << 22); };
    if (!wgrad_tc(d.M, d.E))
      V1T_TRY(ln_forward(S.x2, nullptr, d.T, nullptr, W.ln2_w, W.ln2_b, sc.h, nullptr, d.R, d.E, d.Ep, st));
    const PlaneOp dup = mlp_tc ? act_plane(d, sc.dupl, d.M) : no_plane();  // du planes from the epilogue above
    if (GW.w1) {  // dW1[m,e] = sum_r du[r,m] h2[r,e]
      v1t_gemm_desc g = gd(d.M, d.E, R);
      g.a_m = 1; g.a_k = d.Mp; g.b_k = d.Ep; g.b_n = 1; g.c_m = d.E;
      V1T_TRY(gemm_any_splitk(d.impl, g, sc.g, sc.h, GW.w1, sc.partials, kPartialBytes, st, dup,
                              act_plane(d, S.h2pl, d.E)));
    }
    if (GW.b1 && !mlp_tc) V1T_TRY(colsum(sc.g, GW.b1, 1, d.R, d.M, 0, d.Mp, 0, sc.partials, kPartialBytes, st));
    {  // dh2[r,e] = sum_m du[r,m] W1[m,e]
      v1t_gemm_desc g = gd(R, d.E, d.M);
      g.a_m = d.Mp; g.a_k = 1; g.b_k = d.E; g.b_n = 1; g.c_m = d.Ep;
      V1T_TRY(gemm_any(d.impl, g, sc.g, W.w1, sc.dh, nullptr, nullptr, st, no_drop(), no_epi(), dup, wplane(d, S, kW1)));
    }
    V1T_TRY(ln_backward(sc.dh, S.x2, S.st2, W.ln2_w, dx, GW.ln2_w, GW.ln2_b, sc.partials, kPartialBytes, d.R, d.E,
                        d.Ep, st));
    // ================= attention branch: x2 = x1 + drop(o Wproj^T + b) =================
    const float* da = dx;
    PlaneOp dap = no_plane();
    if (drop) {
      V1T_TRY(dropout_rows(dx, sc.dh, d.R, d.E, d.Ep, site_drop(*shape, i, kSiteProj), st, act_plane_out(d, sc.dpl),
                           GW.bproj, sc.partials, kPartialBytes));
      da = sc.dh;
      dap = act_plane(d, sc.dpl, d.E);
    }
    const bool hp_path = qkv_to_planes(d);  // head-padded plane dataflow through the attention branch
    const int KP = d.heads * d.Ep;
    const PlaneOp wpp{S.wpp[0], d.impl == V1T_IMPL_BF16X3 ? S.wpp[1] : nullptr, (int)round_up(d.E, 32), cdiv(KP, 32)};
    if (GW.wproj) {  // dWp[e,i] = sum_r da[r,e] o[r,i]
      if (hp_path) {  // o from the planes the attention forward wrote; padded columns dropped by the reduction
        v1t_gemm_desc g = gd(d.E, KP, R);
        g.a_m = 1; g.a_k = d.Ep; g.b_k = KP; g.b_n = 1; g.c_m = d.I;
        V1T_TRY(gemm_any_splitk(d.impl, g, da, nullptr, GW.wproj, sc.partials, kPartialBytes, st, dap,
                                act_plane(d, S.opl, KP), GroupMap{0, 0, d.E, d.Ep}));
      } else {
        v1t_gemm_desc g = gd(d.E, d.I, R);
        g.a_m = 1; g.a_k = d.Ep; g.b_k = d.I; g.b_n = 1; g.c_m = d.I;
        V1T_TRY(gemm_any_splitk(d.impl, g, da, S.o, GW.wproj, sc.partials, kPartialBytes, st, dap));
      }
    }
    if (GW.bproj && !drop) V1T_TRY(colsum(da, GW.bproj, 1, d.R, d.E, 0, d.Ep, 0, sc.partials, kPartialBytes, st));
    if (hp_path) {  // dO[r, (h, d)] = sum_e da[r,e] Wp[e, h*E + d]  -> straight into the attention kernels' dO planes
      v1t_gemm_desc g = gd(R, KP, d.E);
      g.a_m = d.Ep; g.a_k = 1; g.b_k = KP; g.b_n = 1;
      EpiOp epi = no_epi();
      epi.kind = kEpiHeadPlanes;
      epi.hp.p[0][0] = sc.planes.dO[0];
      epi.hp.p[0][1] = d.impl == V1T_IMPL_BF16X3 ? sc.planes.dO[1] : nullptr;
      epi.hp.T = d.T; epi.hp.Tq = d.Tq; epi.hp.H = d.heads; epi.hp.AD = d.Ep / 32;
      uint8_t* pads[2] = {sc.planes.dO[0], sc.planes.dO[1]};
      V1T_TRY(zero_plane_pad_rows(pads, d.impl == V1T_IMPL_BF16X3 ? 2 : 1, (int64_t)d.B * d.heads * (d.Ep / 32), d.Tq, d.T, st,
                                  d.Ep / 32));
      V1T_TRY(gemm_any(d.impl, g, da, nullptr, nullptr, nullptr, nullptr, st, no_drop(), epi, dap, wpp));
    } else {  // dO[r,i] = sum_e da[r,e] Wp[e,i]
      v1t_gemm_desc g = gd(R, d.I, d.E);
      g.a_m = d.Ep; g.a_k = 1; g.b_k = d.I; g.b_n = 1; g.c_m = d.I;
      V1T_TRY(gemm_any(d.impl, g, da, W.wproj, sc.dO, nullptr, nullptr, st, no_drop(), no_epi(), dap, wplane(d, S, kWproj)));
    }
    const int64_t ld = 3 * d.I;
    lin1.reset();
    if (d.fused) {
      ProfScope prof(V1T_PHASE_ATTN_BWD, st);
      const int x3 = d.impl == V1T_IMPL_BF16X3;
      const AttnPlanes& pl = sc.planes;
      if (hp_path) {  // dO planes came out of the GEMM above; delta = rowsum(O * dO) from the planes
        V1T_TRY(attn_delta_planes(act_plane(d, S.opl, KP), pl.dO[0], x3 ? pl.dO[1] : nullptr, pl.delta, d.B, d.heads, d.T,
                                  d.Tq, d.Ep / 32, st));
      } else {
        V1T_TRY(make_planes(sc.dO, d.I, 0, d.B, d.heads, d.T, d.Tq, d.E, d.Ep, pl.dO[0], x3 ? pl.dO[1] : nullptr,
                            nullptr, nullptr, st));
        V1T_TRY(attn_delta(S.o, sc.dO, pl.delta, d.B, d.heads, d.T, d.Tq, d.E, d.I, st));
      }
      AttnBwdArgs ba{};
      ba.q_hi = S.qp[0]; ba.q_lo = S.qp[1]; ba.k_hi = S.kp[0]; ba.k_lo = S.kp[1]; ba.v_hi = S.vp[0]; ba.v_lo = S.vp[1];
      ba.do_hi = pl.dO[0]; ba.do_lo = pl.dO[1];
      ba.lse = S.lse; ba.delta = pl.delta;
      ba.ds_hi = pl.dS[0]; ba.ds_lo = x3 ? pl.dS[1] : nullptr;
      ba.drop_bits = d.drop ? S.dbits : nullptr;
      if (qkv_to_planes(d)) {  // dQ | dK | dV leave the kernels as operand planes of the head-padded gradient only
        ba.dqkv = nullptr;
        ba.dq_pl = act_plane_out(d, sc.dqpl);
        uint8_t* pads[2] = {sc.dqpl[0], sc.dqpl[1]};  // rows past B*T of the planes (contracted by the weight gradient)
        V1T_TRY(zero_plane_pad_rows(pads, x3 ? 2 : 1, 3 * d.heads * (d.Ep / 32), (int)round_up(d.R, 32), (int)d.R, st));
      } else {
        ba.dqkv = sc.dqkv;
      }
      ba.B = d.B; ba.H = d.heads; ba.T = d.T; ba.Tp = d.Tq; ba.E = d.E; ba.Dp = d.Ep;
      ba.scale = scale; ba.scale_log2 = scale * 1.4426950408889634f;
      ba.x3 = x3;
      ba.prec = attn_prec_env();
      ba.drop = site_drop(*shape, i, kSiteAttn);
      {
        ProfScope kprof(V1T_PHASE_ATTN_BWD_KERNEL, st);
        V1T_TRY(attn_bwd_dispatch(ba, st));
      }
    }
    for (int b0 = 0; d.matpl && b0 < d.B; b0 += sc.chunk) {
      ProfScope prof(V1T_PHASE_ATTN_BWD, st);
      const int bc = std::min(sc.chunk, d.B - b0);
      const float* q = S.qkv + (int64_t)b0 * d.T * ld;
      float* dq = sc.dqkv + (int64_t)b0 * d.T * ld;
      const bool x3 = d.impl == V1T_IMPL_BF16X3;
      V1T_TRY(bh_planes_of(d, q, ld, 0, bc, sc.bhq, st));
      V1T_TRY(bh_planes_of(d, q, ld, d.I, bc, sc.bhk, st));
      V1T_TRY(bh_planes_of(d, q, ld, 2 * d.I, bc, sc.bhv, st));
      V1T_TRY(bh_planes_of(d, sc.dO + (int64_t)b0 * d.T * d.I, d.I, 0, bc, sc.bhdo, st));
      V1T_TRY(scores_from_planes(d, sc, bc, sc.P1, st));  // recomputed scores
      {  // dPd = dO V^T
        v1t_gemm_desc g = bh_gemm(d, bc, d.T, d.T, d.E, false, false);
        g.c_m = d.Tp; g.c_b1 = (int64_t)d.heads * d.T * d.Tp; g.c_b2 = (int64_t)d.T * d.Tp;
        V1T_TRY(gemm_any(d.impl, g, nullptr, nullptr, sc.P2, nullptr, nullptr, st, no_drop(), no_epi(), bh_op(d, sc.bhdo),
                         bh_op(d, sc.bhv)));
      }
      // softmax, dropout and the softmax backward in one pass over the rows: Pd and dS leave as operand planes
      V1T_TRY(softmax_bwd_rows_planes(sc.P1, sc.P2, (int64_t)bc * d.heads, d.T, d.Tq, d.Tp, site_drop(*shape, i, kSiteAttn),
                                      (int64_t)b0 * d.heads * d.T, sc.pdpl[0], x3 ? sc.pdpl[1] : nullptr, sc.dspl[0],
                                      x3 ? sc.dspl[1] : nullptr, st));
      v1t_gemm_desc g = bh_gemm(d, bc, d.T, d.E, d.T, true, true);
      g.c_m = ld; g.c_b1 = d.T * ld; g.c_b2 = d.E;
      // dV[j,d] = sum_i Pd[i,j] dO[i,d]
      V1T_TRY(gemm_any(d.impl, g, nullptr, nullptr, dq + 2 * d.I, nullptr, nullptr, st, no_drop(), no_epi(), prob_op(d, sc.pdpl),
                       bh_op(d, sc.bhdo)));
      g.alpha = scale;  // dK[j,d] = scale * sum_i dS[i,j] Q[i,d]
      V1T_TRY(gemm_any(d.impl, g, nullptr, nullptr, dq + d.I, nullptr, nullptr, st, no_drop(), no_epi(), prob_op(d, sc.dspl),
                       bh_op(d, sc.bhq)));
      g = bh_gemm(d, bc, d.T, d.E, d.T, false, true);  // dQ[i,d] = scale * sum_j dS[i,j] K[j,d]
      g.alpha = scale;
      g.c_m = ld; g.c_b1 = d.T * ld; g.c_b2 = d.E;
      V1T_TRY(gemm_any(d.impl, g, nullptr, nullptr, dq, nullptr, nullptr, st, no_drop(), no_epi(), prob_op(d, sc.dspl),
                       bh_op(d, sc.bhk)));
    }
    for (int b0 = 0; !d.fused && !d.matpl && b0 < d.B; b0 += sc.chunk) {
      ProfScope prof(V1T_PHASE_ATTN_BWD, st);
      const int bc = std::min(sc.chunk, d.B - b0);
      const float* q = S.qkv + (int64_t)b0 * d.T * ld;
      const float* dO = sc.dO + (int64_t)b0 * d.T * d.I;
      float* dq = sc.dqkv + (int64_t)b0 * d.T * ld;
      const int64_t pb1 = (int64_t)d.heads * d.T * d.Tp, pb2 = (int64_t)d.T * d.Tp;
      V1T_TRY(attention_probs(d, S.qkv, b0, bc, sc.P1, d.Tp, no_drop(), st));  // P (no dropout)
      {  // dPd = dO V^T
        v1t_gemm_desc g = gd(d.T, d.T, d.E);
        g.batch1 = bc; g.batch2 = d.heads;
        g.a_m = d.I; g.a_k = 1; g.a_b1 = (int64_t)d.T * d.I; g.a_b2 = d.E;
        g.b_k = 1; g.b_n = ld; g.b_b1 = d.T * ld; g.b_b2 = d.E;
        g.c_m = d.Tp; g.c_b1 = pb1; g.c_b2 = pb2;
        V1T_TRY(gemm_any(d.impl, g, dO, q + 2 * d.I, sc.P2, nullptr, nullptr, st));
      }
      V1T_TRY(softmax_bwd_rows(sc.P1, sc.P2, (int64_t)bc * d.heads * d.T, d.T, d.Tp, site_drop(*shape, i, kSiteAttn),
                               (int64_t)b0 * d.heads * d.T, st));  // P1 <- Pd, P2 <- dS
      {  // dV[j,d] = sum_i Pd[i,j] dO[i,d]
        v1t_gemm_desc g = gd(d.T, d.E, d.T);
        g.batch1 = bc; g.batch2 = d.heads;
        g.a_m = 1; g.a_k = d.Tp; g.a_b1 = pb1; g.a_b2 = pb2;
        g.b_k = d.I; g.b_n = 1; g.b_b1 = (int64_t)d.T * d.I; g.b_b2 = d.E;
        g.c_m = ld; g.c_b1 = d.T * ld; g.c_b2 = d.E;
        V1T_TRY(gemm_any(d.impl, g, sc.P1, dO, dq + 2 * d.I, nullptr, nullptr, st));
      }
      {  // dQ[i,d] = scale * sum_j dS[i,j] K[j,d]
        v1t_gemm_desc g = gd(d.T, d.E, d.T);
        g.batch1 = bc; g.batch2 = d.heads; g.alpha = scale;
        g.a_m = d.Tp; g.a_k = 1; g.a_b1 = pb1; g.a_b2 = pb2;
        g.b_k = ld; g.b_n = 1; g.b_b1 = d.T * ld; g.b_b2 = d.E;
        g.c_m = ld; g.c_b1 = d.T * ld; g.c_b2 = d.E;
        V1T_TRY(gemm_any(d.impl, g, sc.P2, q + d.I, dq, nullptr, nullptr, st));
      }
      {  // dK[j,d] = scale * sum_i dS[i,j] Q[i,d]
        v1t_gemm_desc g = gd(d.T, d.E, d.T);
        g.batch1 = bc; g.batch2 = d.heads; g.alpha = scale;
        g.a_m = 1; g.a_k = d.Tp; g.a_b1 = pb1; g.a_b2 = pb2;
        g.b_k = ld; g.b_n = 1; g.b_b1 = d.T * ld; g.b_b2 = d.E;
        g.c_m = ld; g.c_b1 = d.T * ld; g.c_b2 = d.E;
        V1T_TRY(gemm_any(d.impl, g, sc.P2, q, dq + d.I, nullptr, nullptr, st));
      }
    }
    ProfScope lin2(V1T_PHASE_LINEAR_BWD, st);
    if (!wgrad_tc(3 * d.I, d.E))
      V1T_TRY(ln_forward(S.x1, nullptr, d.T, nullptr, W.ln1_w, W.ln1_b, sc.h, nullptr, d.R, d.E, d.Ep, st));
    if (qkv_to_planes(d)) {
      // head-padded problem (every head's 155 columns sit in a 160-column group): both GEMMs read dqkv from the
      // planes written by the attention backward; the weight gradient drops the padded rows while reducing
      const int NP = 3 * d.heads * d.Ep;
      const PlaneOp dqp = act_plane(d, sc.dqpl, NP);
      const bool x3 = d.impl == V1T_IMPL_BF16X3;
      if (GW.wqkv) {  // dWqkv[n,e] = sum_r dqkv[r,n] h1[r,e]
        v1t_gemm_desc g = gd(NP, d.E, R);
        g.a_m = 1; g.a_k = NP; g.b_k = d.Ep; g.b_n = 1; g.c_m = d.E;
        V1T_TRY(gemm_any_splitk(d.impl, g, nullptr, sc.h, GW.wqkv, sc.partials, kPartialBytes, st, dqp,
                                act_plane(d, S.h1pl, d.E), GroupMap{d.E, d.Ep, 0, 0}));
      }
      {  // dh1[r,e] = sum_n dqkv[r,n] Wqkv[n,e]
        v1t_gemm_desc g = gd(R, d.E, NP);
        g.a_m = NP; g.a_k = 1; g.b_k = d.E; g.b_n = 1; g.c_m = d.Ep;
        const PlaneOp wq{S.wqp[0], x3 ? S.wqp[1] : nullptr, (int)round_up(NP, 32), cdiv(d.E, 32)};
        V1T_TRY(gemm_any(d.impl, g, nullptr, nullptr, sc.dh, nullptr, nullptr, st, no_drop(), no_epi(), dqp, wq));
      }
    } else {
      if (GW.wqkv) {  // dWqkv[n,e] = sum_r dqkv[r,n] h1[r,e]
        v1t_gemm_desc g = gd(3 * d.I, d.E, R);
        g.a_m = 1; g.a_k = ld; g.b_k = d.Ep; g.b_n = 1; g.c_m = d.E;
        V1T_TRY(gemm_any_splitk(d.impl, g, sc.dqkv, sc.h, GW.wqkv, sc.partials, kPartialBytes, st, no_plane(),
                                act_plane(d, S.h1pl, d.E)));
      }
      {  // dh1[r,e] = sum_n dqkv[r,n] Wqkv[n,e]
        v1t_gemm_desc g = gd(R, d.E, 3 * d.I);
        g.a_m = ld; g.a_k = 1; g.b_k = d.E; g.b_n = 1; g.c_m = d.Ep;
        V1T_TRY(gemm_any(d.impl, g, sc.dqkv, W.wqkv, sc.dh, nullptr, nullptr, st, no_drop(), no_epi(), no_plane(),
                         wplane(d, S, kWqkv)));
      }
    }
    V1T_TRY(ln_backward(sc.dh, S.x1, S.st1, W.ln1_w, dx, GW.ln1_w, GW.ln1_b, sc.partials, kPartialBytes, d.R, d.E,
                        d.Ep, st));
    // ================= behaviour MLP: x1 = x_in + lat[b] =================
    if (d.bdim > 0 && (GW.bw0 || GW.bb0 || GW.bw3 || GW.bb3)) {
      V1T_TRY(colsum(dx, sc.dlat, d.B, d.T, d.E, (int64_t)d.T * d.Ep, d.Ep, d.E, sc.partials, kPartialBytes, st));
      if (bmlp_backward_smem(d.B, d.hid, d.E) <= 200 * 1024) {  // the rest of this backward is one small CTA
        V1T_TRY(bmlp_backward(sc.dlat, S.blat, S.bhid, behaviors, W.bw3, GW.bw0, GW.bb0, GW.bw3, GW.bb3, d.B, d.bdim,
                              d.hid, d.E, st));
        continue;
      }
      V1T_TRY(tanh_grad(sc.dlat, S.blat, sc.dz3, (int64_t)d.B * d.E, st));
      if (GW.bw3) {  // dW3[e,j] = sum_b dz3[b,e] hid[b,j]
        v1t_gemm_desc g = gd(d.E, d.hid, d.B);
        g.a_m = 1; g.a_k = d.E; g.b_k = d.hid; g.b_n = 1; g.c_m = d.hid;
        V1T_TRY(gemm_any(d.impl, g, sc.dz3, S.bhid, GW.bw3, nullptr, nullptr, st));
      }
      if (GW.bb3) V1T_TRY(colsum(sc.dz3, GW.bb3, 1, d.B, d.E, 0, d.E, 0, sc.partials, kPartialBytes, st));
      {  // dhid[b,j] = sum_e dz3[b,e] W3[e,j]
        v1t_gemm_desc g = gd(d.B, d.hid, d.E);
        g.a_m = d.E; g.a_k = 1; g.b_k = d.hid; g.b_n = 1; g.c_m = d.hid;
        V1T_TRY(gemm_any(d.impl, g, sc.dz3, W.bw3, sc.dhid, nullptr, nullptr, st));
      }
      V1T_TRY(tanh_grad(sc.dhid, S.bhid, sc.dhid, (int64_t)d.B * d.hid, st));  // dz0
      if (GW.bw0) {  // dW0[j,i] = sum_b dz0[b,j] beh[b,i]
        v1t_gemm_desc g = gd(d.hid, d.bdim, d.B);
        g.a_m = 1; g.a_k = d.hid; g.b_k = d.bdim; g.b_n = 1; g.c_m = d.bdim;
        V1T_TRY(gemm_any(d.impl, g, sc.dhid, behaviors, GW.bw0, nullptr, nullptr, st));
      }
      if (GW.bb0) V1T_TRY(colsum(sc.dhid, GW.bb0, 1, d.B, d.hid, 0, d.hid, 0, sc.partials, kPartialBytes, st));
    }
  }

  // ================= patch embedding =================
  ProfScope prof_pe(V1T_PHASE_LINEAR_BWD, st);
  if (shape->p_drop_tokens > 0.f) V1T_TRY(dropout_rows(dx, dx, d.R, d.E, d.Ep, site_drop(*shape, 0, kSiteTokens), st));
  V1T_TRY(batchsum(dx, sc.dpos, d.B, d.T, d.E, (int64_t)d.T * d.Ep, d.Ep, d.E, st));
  if (G->pos) V1T_CUDA(cudaMemcpyAsync(G->pos, sc.dpos, sizeof(float) * d.T * d.E, cudaMemcpyDeviceToDevice, st));
  if (G->cls) V1T_CUDA(cudaMemcpyAsync(G->cls, sc.dpos, sizeof(float) * d.E, cudaMemcpyDeviceToDevice, st));
  if (G->bpe) V1T_TRY(colsum(sc.dpos + d.E, G->bpe, 1, d.L, d.E, 0, d.E, 0, sc.partials, kPartialBytes, st));
  if (G->wpe) {  // dWpe[e,k] = sum_{b,l} dx[b,1+l,e] patches[b,l,k]: per-sample partials, then a fixed-order sum
    V1T_TRY(im2col(images, sc.patches, d.B, d.C, d.H, d.W, d.p, d.s, d.gh, d.gw, st));
    const int64_t per = (int64_t)d.E * d.pd;
    int bstep = (int)std::max<int64_t>(1, std::min<int64_t>(d.B, (int64_t)(kPartialBytes / sizeof(float)) / per));
    for (int b0 = 0; b0 < d.B; b0 += bstep) {
      const int bc = std::min(bstep, d.B - b0);
      v1t_gemm_desc g = gd(d.E, d.pd, d.L);
      g.batch1 = bc;
      g.a_m = 1; g.a_k = d.Ep; g.a_b1 = (int64_t)d.T * d.Ep;
      g.b_k = d.pd; g.b_n = 1; g.b_b1 = (int64_t)d.L * d.pd;
      g.c_m = d.pd; g.c_b1 = per;
      V1T_TRY(gemm_any(d.impl, g, dx + (int64_t)b0 * d.T * d.Ep + d.Ep, sc.patches + (int64_t)b0 * d.L * d.pd, sc.partials,
                        nullptr, nullptr, st));
      V1T_TRY(reduce_partials(sc.partials, G->wpe, bc, d.E, d.pd, d.pd, b0 > 0, st));
    }
  }
  if (d_images) {  // d_patches = dtok Wpe -> fold
    v1t_gemm_desc g = gd(d.L, d.pd, d.E);
    g.batch1 = d.B;
    g.a_m = d.Ep; g.a_k = 1; g.a_b1 = (int64_t)d.T * d.Ep;
    g.b_k = d.pd; g.b_n = 1;
    g.c_m = d.pd; g.c_b1 = (int64_t)d.L * d.pd;
    V1T_TRY(gemm_any(d.impl, g, dx + d.Ep, P->wpe, sc.patches, nullptr, nullptr, st));
    V1T_TRY(col2im(sc.patches, d_images, d.B, d.C, d.H, d.W, d.p, d.s, d.gh, d.gw, st));
  }
  return V1T_OK;
}

extern "C" int v1t_attention_probs(const v1t_core_shape* shape, const void* saved_mem, int block, float* probs,
                                   void* stream) {
  Dims d;
  V1T_TRY(make_dims(shape, d));
  V1T_CHECK_ARG(saved_mem && probs && block >= 0 && block < d.blocks, "attention_probs: bad argument");
  Saved sv = carve_saved(d, const_cast<void*>(saved_mem));
  const int chunk = attn_chunk(d);
  if (d.fused) {
    // the forward kept Q, K as operand planes and the base-2 log-sum-exp of every row: P = exp2(S c - lse) comes out of
    // the forward kernel's score pipeline in ONE launch (no fp32 qkv rebuild, no T x T GEMM output + softmax passes)
    const BlockSaved& S = sv.blk[block];
    AttnFwdArgs fa{};
    fa.q_hi = S.qp[0]; fa.q_lo = S.qp[1]; fa.k_hi = S.kp[0]; fa.k_lo = S.kp[1];
    fa.B = d.B; fa.H = d.heads; fa.T = d.T; fa.Tp = d.Tq; fa.E = d.E; fa.Dp = d.Ep;
    fa.scale_log2 = (1.0f / sqrtf((float)d.E)) * 1.4426950408889634f;
    fa.x3 = d.impl == V1T_IMPL_BF16X3;
    fa.probs = probs; fa.lse_in = S.lse;
    return attn_emit_probs_tc(fa, (cudaStream_t)stream);
  }
  if (qkv_to_planes(d))  // the forward kept q, k, v only as bf16 hi/lo planes: rebuild fp32 qkv in its (unused) slot
    V1T_TRY(planes_to_qkv(head_planes(d, sv.blk[block]), d.B, d.E, sv.blk[block].qkv, (cudaStream_t)stream));
  for (int b0 = 0; b0 < d.B; b0 += chunk) {
    const int bc = std::min(chunk, d.B - b0);
    V1T_TRY(attention_probs(d, sv.blk[block].qkv, b0, bc, probs + (int64_t)b0 * d.heads * d.T * d.T, d.T, no_drop(),
                            (cudaStream_t)stream));
  }
  return V1T_OK;
}
