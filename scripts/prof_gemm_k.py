import ctypes as C, sys, torch
sys.path.insert(0, ".")
from v1t_b200 import _lib
lib = _lib.load(); DEV = "cuda:0"
impl = {"bf16x3": _lib.IMPL_BF16X3, "bf16": _lib.IMPL_BF16}[sys.argv[1] if len(sys.argv) > 1 else "bf16x3"]
R = 16 * 1654
def run(m, n, k, reps=5):
    A = torch.randn(m * k, device=DEV); B = torch.randn(n * k, device=DEV); Cm = torch.empty(m * n, device=DEV)
    d = _lib.GemmDesc(m=m, n=n, k=k, batch1=1, batch2=1, alpha=1.0, accumulate=0)
    d.a_m, d.a_k, d.b_k, d.b_n, d.c_m = k, 1, 1, k, n
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: lib.v1t_gemm_tc(C.byref(d), A.data_ptr(), B.data_ptr(), Cm.data_ptr(), None, None, impl, st)
    for _ in range(2): assert call() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for n in (1920, 160, 480):
    print(f"N={n} (M={R}):", "  ".join(f"K={k}: {run(R, n, k):7.1f} us" for k in (32, 160, 320, 640)))
