"""Run a few representative GEMM shapes through the tcgen05 GEMM (for ncu captures / quick event timing)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from v1t_b200 import _lib

lib = _lib.load()
DEV = "cuda:0"
impl = {"bf16x3": _lib.IMPL_BF16X3, "bf16": _lib.IMPL_BF16}[sys.argv[1] if len(sys.argv) > 1 else "bf16x3"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5


def run(name, m, n, k, b1=1, b2=1, lda=None, ldb=None, ldc=None, nt=True):
    lda, ldb, ldc = lda or k, ldb or (k if nt else n), ldc or n
    A = torch.randn(b1 * b2 * m * lda, device=DEV)
    B = torch.randn(b1 * b2 * (n if nt else k) * ldb, device=DEV)
    Cm = torch.empty(b1 * b2 * m * ldc, device=DEV)
    d = _lib.GemmDesc(m=m, n=n, k=k, batch1=b1, batch2=b2, alpha=1.0, accumulate=0)
    d.a_m, d.a_k, d.a_b1, d.a_b2 = lda, 1, b2 * m * lda, m * lda
    if nt:
        d.b_k, d.b_n, d.b_b1, d.b_b2 = 1, ldb, b2 * n * ldb, n * ldb
    else:
        d.b_k, d.b_n, d.b_b1, d.b_b2 = ldb, 1, b2 * k * ldb, k * ldb
    d.c_m, d.c_b1, d.c_b2 = ldc, b2 * m * ldc, m * ldc
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: lib.v1t_gemm_tc(C.byref(d), A.data_ptr(), B.data_ptr(), Cm.data_ptr(), None, None, impl, st)
    for _ in range(2):
        assert call() == 0, _lib.last_error()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * m * n * k * b1 * b2
    print(f"{name:28s} m={m} n={n} k={k} b={b1}x{b2}: {ms * 1e3:9.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)


R = 16 * 1654
run("qkv  [R,160]x[1860,160]^T", R, 1860, 155, lda=160, ldb=160, ldc=1860)
run("proj [R,620]x[155,620]^T", R, 155, 620, lda=620, ldb=620, ldc=160)
run("mlp1 [R,160]x[488,160]^T", R, 488, 155, lda=160, ldb=160, ldc=512)
run("scores QK^T per (b,h)", 1654, 1654, 155, 16, 4, lda=1860, ldb=1860, ldc=1656)
run("PV   [T,T]x[T,155]", 1654, 155, 1654, 16, 4, lda=1656, ldb=1860, ldc=620, nt=False)


def run_g(name, m, n, k, a_m, a_k, b_k, b_n, ldc, asz, bsz):
    """generic strides; single batch"""
    A = torch.randn(asz, device=DEV)
    B = torch.randn(bsz, device=DEV)
    Cm = torch.empty(m * ldc, device=DEV)
    d = _lib.GemmDesc(m=m, n=n, k=k, batch1=1, batch2=1, alpha=1.0, accumulate=0)
    d.a_m, d.a_k, d.b_k, d.b_n, d.c_m = a_m, a_k, b_k, b_n, ldc
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: lib.v1t_gemm_tc(C.byref(d), A.data_ptr(), B.data_ptr(), Cm.data_ptr(), None, None, impl, st)
    for _ in range(2):
        assert call() == 0, _lib.last_error()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} m={m} n={n} k={k}: {ms * 1e3:9.1f} us  {2.0 * m * n * k / ms / 1e9:8.1f} TFLOP/s", flush=True)


print("-- backward shapes (no split-K here: single-pass over K)")
# dgrad: dO[R,620] = da[R,160] @ Wproj[155,620]  (B is N-contiguous)
run_g("dgrad dO = da Wproj", R, 620, 155, 160, 1, 620, 1, 620, R * 160, 155 * 620)
# dgrad: dh1[R,155] = dqkv[R,1860] @ Wqkv[1860,155]
run_g("dgrad dh1 = dqkv Wqkv", R, 155, 1860, 1860, 1, 155, 1, 160, R * 1860, 1860 * 155)
# dgrad: dg[R,488] = dm[R,160] @ W2[155,488]
run_g("dgrad dg = dm W2", R, 488, 155, 160, 1, 488, 1, 512, R * 160, 155 * 488)
# dgrad: dh2[R,155] = du[R,512] @ W1[488,155]
run_g("dgrad dh2 = du W1", R, 155, 488, 512, 1, 155, 1, 160, R * 512, 488 * 155)
# wgrad: dWqkv[1860,155] = dqkv^T[1860,R] @ h1[R,160]   (A is M-contiguous, K = R)
run_g("wgrad dWqkv = dqkv^T h1 (1 pass)", 1860, 155, R, 1, 1860, 160, 1, 156, R * 1860, R * 160)
run_g("wgrad dW1 = du^T h2 (1 pass)", 488, 155, R, 1, 512, 160, 1, 156, R * 512, R * 160)
run_g("wgrad dW2 = dm^T g (1 pass)", 155, 488, R, 1, 160, 512, 1, 488, R * 160, R * 512)
run_g("wgrad dWproj = da^T o (1 pass)", 155, 620, R, 1, 160, 620, 1, 620, R * 160, R * 620)
