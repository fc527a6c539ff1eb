import sys, numpy as np, torch, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from v1t_b200 import _lib
import test_gpu_parity as T
lib = _lib.load()
DEV="cuda:0"
def case(m,n,k,ta,tb,which):
    rng = torch.Generator(device=DEV).manual_seed(1)
    A = torch.randn((k, m) if ta else (m, k), device=DEV, generator=rng)
    Bm = torch.randn((n, k) if tb else (k, n), device=DEV, generator=rng)
    Cm = torch.full((m, n), float("nan"), device=DEV)
    d = _lib.GemmDesc(m=m, n=n, k=k, batch1=1, batch2=1, alpha=1.0, accumulate=0)
    d.a_m, d.a_k = (1, m) if ta else (k, 1)
    d.b_k, d.b_n = (1, k) if tb else (n, 1)
    d.c_m = n
    ah, al = T._planes_of(lib, A) if "a" in which else (None, None)
    bh, bl = T._planes_of(lib, Bm) if "b" in which else (None, None)
    ptr = lambda t: None if t is None else t.data_ptr()
    rc = lib.v1t_gemm_tc_planes(C.byref(d), A.data_ptr(), Bm.data_ptr(), Cm.data_ptr(), None, None, _lib.IMPL_BF16X3,
                                ptr(ah), ptr(al), A.shape[0], A.shape[1], ptr(bh), ptr(bl), Bm.shape[0], Bm.shape[1],
                                torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    ref = ((A.t() if ta else A).double() @ (Bm.t() if tb else Bm).double())
    err = (Cm.double()-ref).abs()
    rowerr = err.max(dim=1).values
    bad = (rowerr > 1e-3).nonzero().flatten()
    print(m,n,k,ta,tb,which, "max err", err.max().item(), "bad rows", bad[:5].tolist(), "...", bad[-5:].tolist(), len(bad))
for w in ["", "b", "a", "ab"]:
    case(128,160,160,False,True,w)
    case(300,155,155,False,True,w)
    case(300,155,155,True,True,w)
    case(1000,620,488,False,False,w)
