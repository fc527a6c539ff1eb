"""Time the model's GEMM shapes with fp32 (converted) operands vs pre-swizzled plane operands."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from v1t_b200 import _lib
lib = _lib.load(); DEV = "cuda:0"; impl = _lib.IMPL_BF16X3
st = torch.cuda.current_stream().cuda_stream


def planes(X):
    nb = lib.v1t_matrix_plane_bytes(X.shape[0], X.shape[1])
    hi = torch.empty(nb, dtype=torch.uint8, device=DEV); lo = torch.empty(nb, dtype=torch.uint8, device=DEV)
    assert lib.v1t_matrix_planes(X.data_ptr(), X.stride(0), X.shape[0], X.shape[1], hi.data_ptr(), lo.data_ptr(), st) == 0
    return hi, lo


def run(name, m, n, k, ta, tb, lda=None, ldb=None, ldc=None):
    A = torch.randn((k, lda or m) if ta else (m, lda or k), device=DEV)[:, :(m if ta else k)]
    B = torch.randn((n, ldb or k) if tb else (k, ldb or n), device=DEV)[:, :(k if tb else n)]
    ldc = ldc or n
    Cm = torch.empty(m, ldc, device=DEV)
    d = _lib.GemmDesc(m=m, n=n, k=k, batch1=1, batch2=1, alpha=1.0, accumulate=0)
    d.a_m, d.a_k = (1, A.stride(0)) if ta else (A.stride(0), 1)
    d.b_k, d.b_n = (1, B.stride(0)) if tb else (B.stride(0), 1)
    d.c_m = ldc
    pa, pb = planes(A), planes(B)
    out = []
    for which in ("", "b", "a", "ab"):
        a = pa if "a" in which else (None, None)
        b = pb if "b" in which else (None, None)
        p = lambda t: None if t is None else t.data_ptr()
        call = lambda: lib.v1t_gemm_tc_planes(C.byref(d), A.data_ptr(), B.data_ptr(), Cm.data_ptr(), None, None, impl,
                                              p(a[0]), p(a[1]), A.shape[0], A.shape[1], p(b[0]), p(b[1]), B.shape[0],
                                              B.shape[1], st)
        for _ in range(2):
            assert call() == 0, _lib.last_error()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            call()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 5 * 1e3)
    print(f"{name:30s} m={m:6d} n={n:5d} k={k:6d}  fp32 {out[0]:7.1f}  B-pl {out[1]:7.1f}  A-pl {out[2]:7.1f}  AB-pl {out[3]:7.1f} us", flush=True)


R = 16 * 1654
run("qkv   h Wqkv^T", R, 1860, 155, False, True, lda=160)
run("proj  o Wp^T", R, 155, 620, False, True, ldc=160)
run("mlp1  h W1^T", R, 488, 155, False, True, lda=160, ldc=512)
run("mlp2  g W2^T", R, 155, 488, False, True, lda=512, ldc=160)
run("dg    dm W2", R, 488, 155, False, False, lda=160, ldc=512)
run("dh2   du W1", R, 155, 488, False, False, lda=512, ldc=160)
run("dO    da Wp", R, 620, 155, False, False, lda=160)
run("dh1   dqkv Wqkv", R, 155, 1860, False, False, ldc=160)
print("-- weight gradients, single pass over K = R (the core runs them split-K)")
run("dW2   dm^T g", 155, 488, R, True, False, lda=160, ldb=512)
run("dW1   du^T h", 488, 155, R, True, False, lda=512, ldb=160)
run("dWp   da^T o", 155, 620, R, True, False, lda=160)
run("dWqkv dqkv^T h", 1860, 155, R, True, False, ldb=160)
