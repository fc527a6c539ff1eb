/* v1t_b200 — C-ABI of the B200-native V1T hot path (ViT core + Gaussian2d readout + Poisson loss).
 *
 * The reference (bryanlimy/V1T) is pure Python/PyTorch and has no FFI; these entry points are what a
 * binding for the hot path replaces (file:line relative to /root/reference/):
 *
 *   v1t_core_forward / v1t_core_backward      ViTCore.forward + its autograd   src/v1t/models/core/vit.py:423-436
 *        (Image2Patches vit.py:122-129, BehaviorMLP :200-202, Attention.mha :267-275,
 *         scaled_dot_product_attention :253-265, MLP :153-154, Transformer.forward :348-362)
 *   v1t_readout_forward / v1t_readout_backward  Gaussian2DReadout.forward + autograd
 *                                                              src/v1t/models/readout/gaussian2d.py:195-278
 *   v1t_elu1_forward/_backward, v1t_poisson_forward/_backward
 *                                             ELU1 (src/v1t/models/utils.py:109-118), PoissonLoss.forward
 *                                             (src/v1t/losses.py:114-119,153-166); also fused into the readout
 *   v1t_attention_probs                       what attention_rollout.Recorder's hook on Attention.attend
 *                                             observes (src/v1t/utils/attention_rollout.py:31-36)
 *
 * Conventions: every function returns 0 on success or a negative V1T_ERR_* code; v1t_last_error() gives the
 * message (thread-local).  All tensor pointers are DEVICE pointers owned by the caller (PyTorch), fp32 unless
 * stated, never allocated or freed here; workspaces are passed in and sized by the *_bytes queries.  Launches
 * go to the given cudaStream_t (passed as void*); nothing synchronises.  One host thread per device.
 */
#ifndef V1T_B200_H
#define V1T_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V1T_MAX_BLOCKS 16

#define V1T_OK 0
#define V1T_ERR_INVALID (-1)   /* bad shape / null pointer / unsupported flag */
#define V1T_ERR_CUDA (-2)      /* a CUDA runtime call or launch failed */
#define V1T_ERR_WORKSPACE (-3) /* workspace too small */

/* arithmetic used for the GEMM/attention contractions */
#define V1T_IMPL_FP32 0   /* fp32 CUDA-core kernels (also: materialised attention probabilities) */
#define V1T_IMPL_BF16X3 1 /* tcgen05 tensor cores, bf16 hi+lo split operands, fp32 accumulate ("exact") */
#define V1T_IMPL_BF16 2   /* tcgen05 tensor cores, plain bf16 operands, fp32 accumulate ("fast") */

typedef struct v1t_core_shape {
  int32_t batch;                 /* B */
  int32_t in_ch, in_h, in_w;     /* image C,H,W after the cropper (1|2, 36, 64) */
  int32_t patch, stride;         /* nn.Unfold kernel/stride (vit.py:68) */
  int32_t emb;                   /* E = emb_dim = head_dim (vit.py:218) */
  int32_t heads;                 /* H */
  int32_t mlp;                   /* M = mlp_dim */
  int32_t blocks;                /* num_blocks <= V1T_MAX_BLOCKS */
  int32_t bdim;                  /* BehaviorMLP input width: 0 (behavior_mode 0/1), 3 (mode 2), 5 (mode 3/4) */
  int32_t impl;                  /* V1T_IMPL_* */
  float p_drop_tokens;           /* Image2Patches dropout p (vit.py:106); 0 in eval */
  float p_drop_block;            /* Transformer dropout p: attention probs, proj out, MLP x2 (vit.py:148-150,227,231) */
  uint64_t seed;                 /* counter-based RNG seed for the dropout masks (replayed in backward) */
} v1t_core_shape;

typedef struct v1t_core_dims { /* derived sizes, see v1t_core_dims_of() */
  int32_t gh, gw;   /* patch grid (29,57) */
  int32_t tokens;   /* T = gh*gw + 1 */
  int32_t emb_ld;   /* row stride (floats) of the token buffers: E rounded up to 32 (155 -> 160) */
  int32_t inner;    /* I = H*E */
  int32_t mlp_ld;   /* row stride of the MLP hidden buffer (M rounded up to 32) */
  int32_t patch_dim;/* C*patch*patch */
  int32_t hid;      /* BehaviorMLP hidden = E/2 */
  int32_t attn_path;/* V1T_ATTN_FUSED or V1T_ATTN_MATERIALISED: which attention kernels this shape/impl runs */
} v1t_core_dims;
/* Attention dispatch (decided by shape and impl, reported here so that no caller is switched silently):
 * FUSED: tcgen05 flash-style kernels, operands resident in tensor memory; needs head dim (= emb) <= 160, because the
 * 128 x Dp fp32 accumulators plus the resident bf16 operands must fit the 512 TMEM columns.
 * MATERIALISED: softmax(QK^T) of a batch chunk is written out ([chunk,H,T,T] fp32) between batched GEMMs -- CUDA-core
 * GEMMs for V1T_IMPL_FP32, tcgen05 GEMMs otherwise (the scaled core, emb 512 = head dim 512). */
#define V1T_ATTN_MATERIALISED 0
#define V1T_ATTN_FUSED 1

/* per-block parameter pointers, reference state-dict layouts; NULL = absent bias */
typedef struct v1t_block_ptrs {
  float *ln1_w, *ln1_b;   /* mha.layer_norm.{weight,bias}        [E]            */
  float *wqkv;            /* mha.to_qkv.weight                   [3*H*E, E]     */
  float *wproj, *bproj;   /* mha.projection.0.{weight,bias}      [E, H*E], [E]  */
  float *ln2_w, *ln2_b;   /* mlp.model.0.{weight,bias}           [E]            */
  float *w1, *b1;         /* mlp.model.1.{weight,bias}           [M, E], [M]    */
  float *w2, *b2;         /* mlp.model.4.{weight,bias}           [E, M], [E]    */
  float *bw0, *bb0;       /* b-mlp.models.<k>.0.{weight,bias}    [E/2, bdim], [E/2] */
  float *bw3, *bb3;       /* b-mlp.models.<k>.3.{weight,bias}    [E, E/2], [E]  */
} v1t_block_ptrs;

typedef struct v1t_core_ptrs { /* used both for parameters (read) and for their gradients (written) */
  float *cls;  /* patch_embedding.cls_token      [1,1,E]      */
  float *pos;  /* patch_embedding.pos_embedding  [T,E]        */
  float *wpe;  /* patch_embedding.projection.2.weight [E, C*p*p] */
  float *bpe;  /* patch_embedding.projection.2.bias   [E]     */
  v1t_block_ptrs blk[V1T_MAX_BLOCKS];
} v1t_core_ptrs;

const char* v1t_last_error(void);
int v1t_version(void);

/* instrumentation for bench.py: kernels launched by this library since load, and optional per-phase device
 * timing with CUDA events recorded on the launching stream around each phase of the path */
#define V1T_PHASE_PATCH 0       /* patch embedding + CLS/pos (K1) */
#define V1T_PHASE_LN_QKV 1      /* behaviour add + LayerNorm + QKV GEMM (K2,K3) */
#define V1T_PHASE_ATTN_FWD 2    /* softmax(QK^T)V (K4) */
#define V1T_PHASE_PROJ 3        /* out projection + residual (K5) */
#define V1T_PHASE_MLP 4         /* LN + MLP (K6,K7) */
#define V1T_PHASE_ATTN_BWD 5    /* attention backward (K4b) */
#define V1T_PHASE_LINEAR_BWD 6  /* dgrad/wgrad GEMMs, LN/GELU backward, reductions (K10) */
#define V1T_PHASE_READOUT_FWD 7 /* readout + ELU1 + Poisson forward (K8) */
#define V1T_PHASE_READOUT_BWD 8 /* readout backward (K9) */
#define V1T_PHASE_ATTN_FWD_KERNEL 9  /* the fused attention forward launch alone (inside V1T_PHASE_ATTN_FWD) */
#define V1T_PHASE_ATTN_BWD_KERNEL 10 /* the fused attention backward launch alone (inside V1T_PHASE_ATTN_BWD) */
#define V1T_PHASE_ATTN_BWD_PAIR 11   /* the two-CTA-cluster dV + dK launch alone (inside V1T_PHASE_ATTN_BWD_KERNEL) */
#define V1T_PHASE_ATTN_BWD_DQ 12     /* the dQ launch alone (inside V1T_PHASE_ATTN_BWD_KERNEL) */
#define V1T_NUM_PHASES 13
uint64_t v1t_launch_count(void);
int v1t_prof_enable(int on);
int v1t_prof_reset(void);
int v1t_prof_read(int phase, float* total_ms, int* count);

int v1t_core_dims_of(const v1t_core_shape* shape, v1t_core_dims* out);

/* bytes of the activations kept between forward and backward, and of the transient scratch */
size_t v1t_core_saved_bytes(const v1t_core_shape* shape);
size_t v1t_core_scratch_bytes(const v1t_core_shape* shape);

/* ViTCore.forward.  images [B,C,H,W]; behaviors [B,bdim] (= cat(behaviors,pupil_centers) for mode 3/4) or NULL.
 * tokens_out [B, T, emb_ld]: final residual stream; the core output map is rows 1.. (CLS dropped), cols 0..E-1.
 * saved: v1t_core_saved_bytes (only written when keep_for_backward != 0); scratch: v1t_core_scratch_bytes. */
int v1t_core_forward(const v1t_core_shape* shape, const v1t_core_ptrs* params, const float* images,
                     const float* behaviors, float* tokens_out, void* saved, void* scratch,
                     int keep_for_backward, void* stream);

/* autograd of v1t_core_forward.  d_tokens [B,T,emb_ld] is dL/d(tokens_out) and is CLOBBERED.
 * grads: every non-NULL pointer is overwritten with that parameter's gradient.  d_images may be NULL. */
int v1t_core_backward(const v1t_core_shape* shape, const v1t_core_ptrs* params, const float* images,
                      const float* behaviors, float* d_tokens, const void* saved, void* scratch,
                      const v1t_core_ptrs* grads, float* d_images, void* stream);

/* softmax(QK^T * E^-0.5) of block `block` for the given input: probs [B,H,T,T] (the tensor the reference's
 * Attention.attend emits); uses the fp32 kernels.  Needs the `saved` buffer of a keep_for_backward forward. */
int v1t_attention_probs(const v1t_core_shape* shape, const void* saved, int block, float* probs, void* stream);

/* Fused attention on the tcgen05 tensor cores (K4 of SURVEY.md): qkv [B,T,3*H*E] fp32 packed as to_qkv emits it
 * (q | k | v, heads concatenated; vit.py:269-272) -> out [B,T,H*E] = softmax(q k^T E^-0.5) (dropout) v with heads
 * concatenated ('b h n d -> b n (h d)').  lse_out [B*H, roundup(T,128)] receives the base-2 log-sum-exp per row
 * (may be NULL).  impl = V1T_IMPL_BF16X3 | V1T_IMPL_BF16.  scratch: v1t_attn_scratch_bytes. */
size_t v1t_attn_scratch_bytes(int B, int H, int T, int E);
int v1t_attn_forward(const float* qkv, int B, int H, int T, int E, int impl, float p_drop, uint64_t seed,
                     uint32_t site, float* out, float* lse_out, void* scratch, void* stream);

/* autograd of v1t_attn_forward: out / lse are the forward's results, d_out = dL/d(out); writes d_qkv [B,T,3*H*E]
 * (dq | dk | dv).  Two atomic-free kernels (per key tile: dK, dV; per query tile: dQ), P recomputed from lse. */
int v1t_attn_backward(const float* qkv, const float* out, const float* d_out, const float* lse, int B, int H, int T,
                      int E, int impl, float p_drop, uint64_t seed, uint32_t site, float* d_qkv, void* scratch,
                      void* stream);

/* ---- Gaussian2d readout ------------------------------------------------------------------------------ */
typedef struct v1t_readout_shape {
  int32_t batch;        /* B */
  int32_t neurons;      /* N */
  int32_t channels;     /* C = core emb dim */
  int32_t gh, gw;       /* feature-map height, width (29, 57) */
  int64_t fs_b, fs_y, fs_x; /* fmap element strides for batch / row / col; channel stride must be 1 */
} v1t_readout_shape;

size_t v1t_readout_scratch_bytes(const v1t_readout_shape* s);

/* z[b,n] = sum_c bilinear(fmap[b,:,:,c]; grid[b,n]) * features[c,n] + bias[n]
 * grid = clamp(mu[n] + sigma[n] @ noise[b,n], -1, 1) + shifts[b]      (gaussian2d.py:219-235,267-270)
 * mu [N,2]; sigma [N,2,2]; noise [B,N,2] or NULL (eval); shifts [B,2] or NULL; features [C,N]; bias [N] or NULL.
 * y_true [B,N] or NULL.  Outputs: z [B,N] (pre-activation); if y_out != NULL: y = elu(z)+1;
 * if loss_out != NULL (needs y_true): loss = loss_scale * sum((y+eps) - (y_true+eps) log(y+eps)). */
int v1t_readout_forward(const v1t_readout_shape* s, const float* fmap, const float* mu, const float* sigma,
                        const float* noise, const float* shifts, const float* features, const float* bias,
                        const float* y_true, float loss_scale, float* z, float* y_out, float* loss_out,
                        void* scratch, void* stream);

/* backward.  dz [B,N] = dL/dz, or NULL with y_true given: then dz = dloss * loss_scale * dPoisson/dz (fused).
 * Outputs (any may be NULL): d_fmap (same strides as fmap, ACCUMULATED into: caller zero-fills; V1T_READOUT_DFMAP=sorted
 * selects the pixel-major atomic-free pass, bitwise reproducible),
 * d_mu [N,2], d_sigma [N,2,2], d_shifts [B,2], d_features [C,N], d_bias [N]. */
int v1t_readout_backward(const v1t_readout_shape* s, const float* fmap, const float* mu, const float* sigma,
                         const float* noise, const float* shifts, const float* features, const float* z,
                         const float* dz, const float* y_true, float loss_scale, float dloss, float* d_fmap,
                         float* d_mu, float* d_sigma, float* d_shifts, float* d_features, float* d_bias,
                         void* scratch, void* stream);

/* standalone ELU1 + Poisson (strict drop-in mode: Model owns ELU1, train_step owns the criterion) */
int v1t_elu1_forward(const float* z, float* y, int64_t n, void* stream);
int v1t_elu1_backward(const float* z, const float* dy, float* dz, int64_t n, void* stream);
size_t v1t_poisson_scratch_bytes(int64_t n);
int v1t_poisson_forward(const float* y_pred, const float* y_true, int64_t n, float eps, float loss_scale,
                        float* loss_out, void* scratch, void* stream);
int v1t_poisson_backward(const float* y_pred, const float* y_true, int64_t n, float eps, float loss_scale,
                         const float* dloss, float* dy, void* stream);

/* ---- callers either side of the path (SURVEY.md section 8f "next" rows) ------------------------------- */

/* n1: fused L1 regulariser + AdamW over every parameter tensor in one launch.  Replaces the regulariser graph
 * (vit.py:419-421, gaussian2d.py:83-100, core_shifter.py:35-36, model.py:141-149, train.py:71-73) and
 * torch.optim.AdamW(weight_decay=0).step() (train.py:217-223,77-80).  Per element:
 *   g = grad_scale * grad + l1 * sign(p);  p *= 1 - lr * weight_decay;  m += (g - m)(1 - beta1);
 *   v = beta2 v + (1 - beta2) g g;  p -= lr / bias_corr1 * m / (sqrt(v) / bias_corr2_sqrt + eps)
 * tensors_dev: DEVICE array of n_tensors records; chunk_prefix_dev: DEVICE int32 [n_tensors + 1], prefix sums of
 * ceil(numel / v1t_opt_chunk_elems()); n_chunks = chunk_prefix[n_tensors].  bias_corr1 = 1 - beta1^t,
 * bias_corr2_sqrt = sqrt(1 - beta2^t) for the step count t the caller keeps; the hyper-parameters are doubles so that
 * 1 - beta is formed in double precision as torch does (1 - float(0.9999) is off by 1.7e-4 relative).  zero_grad != 0 clears the gradients in
 * the same pass.  l1_sums_dev (optional, [n_groups]) receives sum |p| per group BEFORE the update (what the
 * reference logs as reg_loss / reg_scale); needs scratch of v1t_adamw_l1_scratch_bytes(n_chunks). */
typedef struct v1t_opt_tensor {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
  float lr;            /* learning rate of the tensor's param group */
  float l1;            /* reg_scale * (times the regulariser entered the loss this step); 0 = not regularised */
  float weight_decay;  /* decoupled decay (the reference passes 0) */
  int32_t group;       /* index into l1_sums_dev */
} v1t_opt_tensor;
int v1t_opt_chunk_elems(void);
size_t v1t_adamw_l1_scratch_bytes(int n_chunks);
int v1t_adamw_l1_step(const v1t_opt_tensor* tensors_dev, const int32_t* chunk_prefix_dev, int n_tensors, int n_chunks,
                      double beta1, double beta2, double eps, double bias_corr1, double bias_corr2_sqrt,
                      double grad_scale, int zero_grad, float* l1_sums_dev, int n_groups, void* scratch,
                      void* stream);

/* a11 / n3: the small MLPs either side of the readout as one kernel per direction: the readout's grid predictor
 * Linear-ELU-Linear-Tanh over the neurons (gaussian2d.py:102-136,188-193) and the core / image shifters'
 * Linear-Tanh stacks over the batch (core_shifter.py:24-40, image_cropper.py:27-48).  Weights in nn.Linear layout
 * ([out, in], bias [out] or NULL).  The backward recomputes the activations from x (nothing saved) and OVERWRITES
 * every non-NULL gradient pointer; no gradient with respect to x (both inputs are data). */
#define V1T_MLP_MAX_LAYERS 3
#define V1T_MLP_MAX_WIDTH 32
#define V1T_ACT_NONE 0
#define V1T_ACT_TANH 1
#define V1T_ACT_ELU 2
typedef struct v1t_mlp_spec {
  int32_t rows;                          /* neurons (grid predictor) or samples (shifters) */
  int32_t layers;                        /* 1..V1T_MLP_MAX_LAYERS Linear layers, layer l followed by act[l] */
  int32_t width[V1T_MLP_MAX_LAYERS + 1]; /* width[0] = input features, width[l+1] = outputs of layer l; <= 32 */
  int32_t act[V1T_MLP_MAX_LAYERS];       /* V1T_ACT_* */
  int64_t x_ld;                          /* row stride of x in floats (>= width[0]) */
} v1t_mlp_spec;
typedef struct v1t_mlp_ptrs {
  float* w[V1T_MLP_MAX_LAYERS];
  float* b[V1T_MLP_MAX_LAYERS];
} v1t_mlp_ptrs;
size_t v1t_small_mlp_scratch_bytes(const v1t_mlp_spec* s);
int v1t_small_mlp_forward(const v1t_mlp_spec* s, const v1t_mlp_ptrs* params, const float* x, float* y, void* stream);
int v1t_small_mlp_backward(const v1t_mlp_spec* s, const v1t_mlp_ptrs* params, const float* x, const float* dy,
                           const v1t_mlp_ptrs* grads, void* scratch, void* stream);

/* n3: ImageCropper.forward (image_cropper.py:120-140) as one gather: images [B,C,in_h,in_w]; grid [crop_h,crop_w,2]
 * is the module's (x,y) crop-grid buffer (image_cropper.py:104-112); shifts [B,2] or NULL (ImageShifter output);
 * nearest sampling with align_corners and zero padding, then bilinear resize (antialias off) to (out_h,out_w) when
 * it differs from the crop, then `behavior_planes` constant planes from behaviors [B,behavior_planes]
 * (behavior_mode 1).  out [B, C + behavior_planes, out_h, out_w].  No gradient: nearest sampling has none. */
typedef struct v1t_crop_shape {
  int32_t batch, channels, in_h, in_w;
  int32_t crop_h, crop_w;
  int32_t out_h, out_w;
  int32_t behavior_planes;
} v1t_crop_shape;
int v1t_crop_resize(const v1t_crop_shape* s, const float* images, const float* grid, const float* shifts,
                    const float* behaviors, float* out, void* stream);

/* n2: attention rollout of recorded attention maps (attention_rollout.py:92-133, attention_rollouts): attn
 * [B,L,H,T,T] (what Recorder stacks, attention_rollout.py:72-75) -> heatmaps [B,out_h,out_w]: head max, + identity,
 * row-normalise, product over the blocks, row 0 without the CLS column as a (gh,gw) map, min-max normalised and
 * bilinear-resized.  Only row 0 of the product is needed, so each block is one vector-matrix pass over its
 * attention matrices (each read once; the last block contributes only its first row). */
size_t v1t_rollout_scratch_bytes(int B, int T);
int v1t_attention_rollout(const float* attn, int B, int L, int H, int T, int gh, int gw, int out_h, int out_w,
                          float* heatmaps, void* scratch, void* stream);

/* n4: ensemble output module (ensemble.py:30-80 OutputModule, :131-151 EnsembleModel.forward): the K members'
 * pre-activation responses x_k [n = B*N each] -> y = elu(sum_k weight[k] x_k + bias) + 1, or with weight == NULL
 * the mean over members (ensemble_mode 0); the [B,N,K] stack the reference concatenates is never built.  Backward
 * (modes 1/2, the only trained parameters of fit_ensemble): d_weight [K], d_bias [1] from dy. */
#define V1T_ENSEMBLE_MAX 16
typedef struct v1t_ensemble_members {
  const float* x[V1T_ENSEMBLE_MAX];
  int32_t count;
} v1t_ensemble_members;
size_t v1t_ensemble_scratch_bytes(int64_t n, int count);
int v1t_ensemble_forward(const v1t_ensemble_members* members, const float* weight, const float* bias, int64_t n,
                         float* y, void* stream);
int v1t_ensemble_backward(const v1t_ensemble_members* members, const float* weight, const float* bias,
                          const float* dy, int64_t n, float* d_weight, float* d_bias, void* scratch, void* stream);

/* ---- building blocks, exported for unit tests -------------------------------------------------------- */
/* C[b][m,n] = alpha * sum_k A[b][m,k] * B[b][k,n] (+ bias[n]) (+ R[b][m,n]); arbitrary element strides */
typedef struct v1t_gemm_desc {
  int32_t m, n, k, batch1, batch2;
  int64_t a_m, a_k, a_b1, a_b2;
  int64_t b_k, b_n, b_b1, b_b2;
  int64_t c_m, c_b1, c_b2;      /* C column stride is 1 */
  int64_t r_m, r_b1, r_b2;      /* residual operand (optional) */
  float alpha;
  int32_t accumulate;           /* C += ... instead of C = ... */
} v1t_gemm_desc;
int v1t_gemm_fp32(const v1t_gemm_desc* d, const float* A, const float* B, float* C, const float* bias,
                  const float* R, void* stream);
/* same contract on the tcgen05 tensor cores; impl = V1T_IMPL_BF16X3 or V1T_IMPL_BF16 */
/* test knob: stage M/N-contiguous operands un-transposed and use MN-major UMMA descriptors (default on) */
int v1t_gemm_tc_set_mn_major(int on);
int v1t_gemm_tc(const v1t_gemm_desc* d, const float* A, const float* B, float* C, const float* bias,
                const float* R, int impl, void* stream);

/* GEMM operands as pre-swizzled bf16 planes.  A row-major matrix X[rows, cols] (leading dim ld) is converted once
 * into a hi plane (and a lo plane, x ~= hi + lo) laid out like the shared-memory tiles tcgen05.mma reads: 32-column
 * atoms, [ceil(cols/32)][round_up(rows,32)][64 B], 16-byte chunks XOR-swizzled with ((row >> 1) & 3), zero padded.
 * v1t_gemm_tc_planes is v1t_gemm_tc with either operand optionally replaced by the planes of the matrix it views
 * (a_hi/b_hi NULL = use the fp32 pointer): the planes are bulk-copied into the MMA stages without conversion work.
 * The operand must be the whole matrix the planes were made from (K-major when d->a_k / d->b_k == 1, else
 * M/N-major); unbatched problems only.  This is how the core feeds weights (nn.Linear, vit.py:146,221,230). */
size_t v1t_matrix_plane_bytes(int64_t rows, int64_t cols);
int v1t_matrix_planes(const float* X, int64_t ld, int64_t rows, int64_t cols, void* hi, void* lo, void* stream);
int v1t_gemm_tc_planes(const v1t_gemm_desc* d, const float* A, const float* B, float* C, const float* bias,
                       const float* R, int impl, const void* a_hi, const void* a_lo, int64_t a_rows, int64_t a_cols,
                       const void* b_hi, const void* b_lo, int64_t b_rows, int64_t b_cols, void* stream);

/* the inverted-dropout multipliers (0 or 1/(1-p)) the kernels apply at dropout site `site` =
 * block*8 + {0 tokens, 1 attention probs, 2 proj out, 3 MLP hidden, 4 MLP out}; element index = row-major
 * index in the logical tensor ([B,T,E], [B,T,M], [B,H,T,T]) with the LAST dimension's stride rounded up to 4
 * (so one Philox call covers 4 adjacent columns).  Lets tests replay the exact masks. */
int v1t_dropout_mask(float* out, int64_t n, uint64_t seed, uint32_t site, float p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* V1T_B200_H */
