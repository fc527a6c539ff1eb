"""Per-tile timeline of one cluster of the pair attention-backward kernel (GPU box):  python scripts/pair_trace.py"""
import os, sys
import torch
sys.path.insert(0, ".")
from v1t_b200 import _lib
lib = _lib.load(); diag = _lib.load_diag()
DEV = "cuda:0"
os.environ["V1T_ATTN_BWD"] = "pair"
B, H, T, E, p = 16, 4, 1654, 155, float(sys.argv[1]) if len(sys.argv) > 1 else 0.2544
impl = _lib.IMPL_BF16X3
g = torch.Generator(device=DEV).manual_seed(1)
qkv = torch.randn(B, T, 3 * H * E, device=DEV, generator=g)
d_out = torch.randn(B, T, H * E, device=DEV, generator=g)
out = torch.empty(B, T, H * E, device=DEV)
Tp = (T + 127) // 128 * 128
lse = torch.zeros(B * H, Tp, device=DEV)
d_qkv = torch.empty((B, T, 3 * H * E), device=DEV)
scratch = torch.empty(lib.v1t_attn_scratch_bytes(B, H, T, E), dtype=torch.uint8, device=DEV)
st = torch.cuda.current_stream().cuda_stream
assert lib.v1t_attn_forward(qkv.data_ptr(), B, H, T, E, impl, p, 4242, 3, out.data_ptr(), lse.data_ptr(), scratch.data_ptr(), st) == 0
trace = torch.zeros(2 * 2 * 32 * 8, dtype=torch.int64, device=DEV)
for it in range(3):
    if it == 2:
        assert diag.v1t_diag_attn_pair_trace(trace.data_ptr()) == 0
    assert lib.v1t_attn_backward(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), lse.data_ptr(), B, H, T, E, impl, p, 4242, 3,
                                 d_qkv.data_ptr(), scratch.data_ptr(), st) == 0
torch.cuda.synchronize()
diag.v1t_diag_attn_pair_trace(None)
t = trace.view(2, 2, 32, 8).cpu()  # [item][rank][tile][event]
names = ["mma:S wait", "mma:S issue", "mma:O wait", "mma:O issue", "sm:S arrived", "sm:xchg ok", "sm:slot free", "sm:A written"]
nt = (T + 63) // 64
for k in range(2):
    for r in range(2):
        print(f"== item {k} rank {r} (cycles since cluster start; dropout p={p})")
        print("tile " + " ".join(f"{n:>13s}" for n in names))
        for j in list(range(4)) + list(range(nt - 3, nt)):
            print(f"{j:4d} " + " ".join(f"{int(t[k, r, j, e]):13d}" for e in range(8)))
        d = (t[k, r, 20, 1] - t[k, r, 4, 1]).item() / 16
        x = t[k, r, 31]
        print(f"   steady-state cycles per tile (S issue, tiles 4..20): {d:.0f}")
        print(f"   hand-over: accumulator complete {int(x[0])}, epilogue done {int(x[1])}, resident stored {int(x[2])}, "
              f"resident visible to the MMA warp {int(x[3])}")
