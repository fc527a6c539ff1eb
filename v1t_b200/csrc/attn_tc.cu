// Fused attention on tcgen05: plane carving, dispatch and the standalone C-ABI entry points (reference:
// Attention.scaled_dot_product_attention, vit.py:253-265, and its autograd).
//
//   P = softmax(Q K^T * E^-0.5) (dropout)   O = P V        per (sample, head), T = 1654 tokens, head dim E = 155
//
// The kernels live in attn_fwd2.cu (forward: Q and P in tensor memory, two-pass softmax, K / V streamed with
// cp.async.bulk from pre-swizzled bf16 planes, nothing of size T x T reaches HBM) and attn_bwd2.cu (backward: three
// atomic-free launches dV / dK / dQ with P recomputed from the saved base-2 log-sum-exp).  The first-generation
// kernels (operands resident in shared memory, SS-form MMAs) were retired once the TMEM-resident versions beat
// them by 1.6-2.3x; they remain in the git history.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

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
  for (int i = 0; i < 2; ++i) { p.q[i] = take(); p.k[i] = take(); p.v[i] = take(); }
  if (with_backward) {
    for (int i = 0; i < 2; ++i) p.dO[i] = take();
    const size_t sb = (size_t)round_up((int64_t)B * H * (Tp / 32) * (int64_t)Tp * 64, 1024);
    for (int i = 0; i < 2; ++i) {
      p.dS[i] = c ? (uint8_t*)(c + off) : nullptr;
      off += sb;
    }
  }
  p.lse = (float*)(c ? c + off : nullptr);
  off += (size_t)round_up((int64_t)B * H * Tp * 4, 1024);
  p.delta = (float*)(c ? c + off : nullptr);
  off += (size_t)round_up((int64_t)B * H * Tp * 4, 1024);
  p.drop_bits = (uint8_t*)(c ? c + off : nullptr);  // standalone entry points: forward writes, backward reads
  off += (size_t)round_up((int64_t)attn_drop_bits_bytes(B, H, Tp), 1024);
  p.total = off;
  return p;
}
}  // namespace v1t

namespace v1t {
int attn_prec_env() {
  const char* e = getenv("V1T_ATTN_PREC");  // read on every call: tests switch it between launches
  return e ? atoi(e) & 7 : 0;
}
int attn_bwd_pair_env() {  // default: pair; V1T_ATTN_BWD=three selects the three-pass organisation
  const char* e = getenv("V1T_ATTN_BWD");
  return !(e && e[0] == 't');
}
int attn_dq_gemm_env() {  // default: dQ as a batched GEMM over the dS' planes of the pair kernel; V1T_ATTN_DQ=pass: dQ pass
  const char* e = getenv("V1T_ATTN_DQ");
  return !(e && e[0] == 'p');
}
int attn_bwd_group_env() {
  const char* e = getenv("V1T_ATTN_BWD_GROUP");
  const int g = e ? atoi(e) : 16;
  return g > 0 ? g : 16;
}
int attn_fwd_dispatch(const AttnFwdArgs& a, cudaStream_t st) {
  return attn_fwd2_tc(a, st);
}
int attn_bwd_dispatch(const AttnBwdArgs& a, cudaStream_t st) {
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
  V1T_TRY(make_planes(qkv, 3 * I, 2 * I, B, H, T, Tp, E, Dp, p.v[0], x3 ? p.v[1] : nullptr, nullptr, nullptr, st));
  AttnFwdArgs a{};
  a.q_hi = p.q[0]; a.q_lo = p.q[1]; a.k_hi = p.k[0]; a.k_lo = p.k[1]; a.v_hi = p.v[0]; a.v_lo = p.v[1];
  a.O = out; a.o_ld = I; a.lse = lse_out ? lse_out : p.lse;
  a.drop_bits = p_drop > 0.f ? p.drop_bits : nullptr;
  a.B = B; a.H = H; a.T = T; a.Tp = Tp; a.E = E; a.Dp = Dp;
  a.scale_log2 = (1.0f / sqrtf((float)E)) * 1.4426950408889634f;
  a.x3 = x3;
  a.prec = attn_prec_env();
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
  a.ds_hi = p.dS[0]; a.ds_lo = x3 ? p.dS[1] : nullptr;
  a.drop_bits = p_drop > 0.f ? p.drop_bits : nullptr;  // written by v1t_attn_forward into the same scratch
  a.B = B; a.H = H; a.T = T; a.Tp = Tp; a.E = E; a.Dp = Dp;
  a.scale = 1.0f / sqrtf((float)E);
  a.scale_log2 = a.scale * 1.4426950408889634f;
  a.x3 = x3;
  a.prec = attn_prec_env();
  a.drop = DropSpec{seed, site, p_drop};
  return attn_bwd_dispatch(a, st);
}
