"""Per-SM cp.async.bulk fill rate vs copy size and bytes in flight (all 148 SMs streaming concurrently)."""
import sys, torch
sys.path.insert(0, ".")
from v1t_b200 import _lib
lib = _lib.load_diag(); DEV = "cuda:0"
st = torch.cuda.current_stream().cuda_stream
out = torch.zeros(148, dtype=torch.int64, device=DEV)
clk = 1.965e9
for span_mb, label in ((64, "L2-resident 64 MB"), (2048, "HBM 2 GB")):
    src = torch.empty(span_mb << 20, dtype=torch.uint8, device=DEV).random_(0, 255)
    print(f"-- source: {label}")
    for nbytes, copies, slots in ((4096, 10, 2), (4096, 10, 4), (8192, 5, 2), (8192, 5, 4), (20480, 2, 2), (20480, 2, 4),
                                  (40960, 1, 2), (40960, 1, 4), (2048, 18, 3), (2048, 18, 5), (16384, 2, 6)):
        iters = 400
        for _ in range(2):
            assert lib.v1t_bulk_microbench(src.data_ptr(), src.numel(), nbytes, copies, slots, iters, out.data_ptr(), st) == 0, lib.v1t_diag_last_error()
        torch.cuda.synchronize()
        cyc = out.max().item()
        total = nbytes * copies * iters
        gbs = total / (cyc / clk) / 1e9
        print(f"copy {nbytes:6d} B x {copies:2d} per slot, {slots} slots ({nbytes*copies*slots/1024:5.0f} KB in flight): "
              f"{gbs:6.1f} GB/s per SM, {gbs*148/1e3:5.2f} TB/s aggregate", flush=True)
