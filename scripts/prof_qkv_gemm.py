"""One QKV-shaped GEMM as the model issues it (A = LN output [R,160], B = Wqkv [155,1860] N-contiguous) for ncu."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from v1t_b200 import _lib
lib = _lib.load(); DEV = "cuda:0"; impl = _lib.IMPL_BF16X3
R = 16 * 1654
m, n, k = R, 1860, 155
A = torch.randn(R * 160, device=DEV); B = torch.randn(155 * 1860, device=DEV); Cm = torch.empty(m * 1860, device=DEV)
d = _lib.GemmDesc(m=m, n=n, k=k, batch1=1, batch2=1, alpha=1.0, accumulate=0)
d.a_m, d.a_k, d.b_k, d.b_n, d.c_m = 160, 1, 1860, 1, 1860
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    assert lib.v1t_gemm_tc(C.byref(d), A.data_ptr(), B.data_ptr(), Cm.data_ptr(), None, None, impl, st) == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    lib.v1t_gemm_tc(C.byref(d), A.data_ptr(), B.data_ptr(), Cm.data_ptr(), None, None, impl, st)
e1.record(); torch.cuda.synchronize()
print("qkv gemm us:", e0.elapsed_time(e1) / 5 * 1e3)
