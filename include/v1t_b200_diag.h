/* v1t_b200 diagnostics — NOT part of the product C-ABI (include/v1t_b200.h).
 *
 * Micro-benchmarks and a self-test of the sm_100a primitives the kernels are built on (tcgen05.mma issue cost in its
 * SS / TS forms, per-SM cp.async.bulk fill rate, tensor-memory A operand).  Built into a separate library,
 * v1t_b200/libv1t_b200_diag.so, from csrc/diag/; used by scripts/*_microbench.py and one GPU test only. */
#ifndef V1T_B200_DIAG_H
#define V1T_B200_DIAG_H

#ifdef __cplusplus
extern "C" {
#endif

const char* v1t_diag_last_error(void);

/* measurement helper: cycles for iters*8 tcgen05.mma (M=128, K=16, bf16) of width N on all SMs; ts=1: A operand
 * from tensor memory, mn_b=1: MN-major B.  out_dev: 148 int64 cycle counts (device memory). */
int v1t_mma_microbench(int N, int ts, int iters, int mn_b, long long* out_dev, void* stream);

/* measurement helper: every SM streams iters rounds of `copies` cp.async.bulk copies of `bytes` bytes from a global
 * buffer (span bytes, wrapped) into a ring of `slots` shared-memory slots; out_dev: 148 int64 cycle counts.  Measures
 * the per-SM global->shared fill ceiling that bounds the bf16x3 main loops. */
int v1t_bulk_microbench(const void* src, long long span, int bytes, int copies, int slots, int iters,
                        long long* out_dev, void* stream);

/* self-test of the tensor-memory A operand (tcgen05.st + TS-form tcgen05.mma): C[128,N] = bf16(A[128,K]) bf16(B[N,K])^T */
int v1t_ts_selftest(const float* A, const float* B, float* C, int N, int K, void* stream);

/* hook exported by the PRODUCT library (libv1t_b200.so), not by the diagnostics library: device buffer of
 * 2 items x 2 ranks x 32 tiles x 8 events (int64 cycles since the cluster's start barrier; tile row 31 = the hand-over
 * between items, see attn_bwd2.cu) that the first cluster of the
 * pair attention-backward kernel fill (events: MMA warp 0 = scores wait begins, 1 = scores issue, 2 = accumulate wait
 * begins, 3 = accumulate issue; softmax warp 0: 4 = scores arrived, 5 = exchange wait passed, 6 = operand slot free,
 * 7 = operand written); NULL switches the trace off.  scripts/pair_trace.py prints the timeline. */
int v1t_diag_attn_pair_trace(long long* buf);

#ifdef __cplusplus
}
#endif
#endif
