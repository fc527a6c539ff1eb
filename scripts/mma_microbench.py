import sys, torch
sys.path.insert(0, ".")
from v1t_b200 import _lib
lib = _lib.load_diag()
out = torch.zeros(148, dtype=torch.int64, device="cuda:0")
iters = 200
print("cycles per tcgen05.mma (M=128, K=16, bf16), all 148 SMs busy; compute floor = N/2")
for ts in (0, 1):
    for mn in (0, 1):
        row = []
        for N in (16, 32, 64, 128, 160, 256):
            rc = lib.v1t_mma_microbench(N, ts, iters, mn, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            assert rc == 0, lib.v1t_diag_last_error()
            torch.cuda.synchronize()
            row.append(f"N={N}: {out.max().item() / (iters * 8):6.1f}")
        print(("TS (A in TMEM)" if ts else "SS (A in smem)"), ("B MN-major" if mn else "B K-major "), " | ".join(row))
