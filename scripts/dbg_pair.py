"""Locate differences between the pair and the three-pass attention backward (GPU box)."""
import os, sys
import torch
sys.path.insert(0, ".")
from v1t_b200 import _lib
lib = _lib.load()
DEV = "cuda:0"

def run(B, H, T, E, p, variant, impl=_lib.IMPL_BF16X3):
    os.environ["V1T_ATTN_BWD"] = variant
    g = torch.Generator(device=DEV).manual_seed(B * 1000 + T + 1)
    qkv = torch.randn(B, T, 3 * H * E, device=DEV, generator=g)
    d_out = torch.randn(B, T, H * E, device=DEV, generator=g)
    out = torch.empty(B, T, H * E, device=DEV)
    Tp = (T + 127) // 128 * 128
    lse = torch.zeros(B * H, Tp, device=DEV)
    d_qkv = torch.full((B, T, 3 * H * E), float("nan"), device=DEV)
    scratch = torch.empty(lib.v1t_attn_scratch_bytes(B, H, T, E), dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.v1t_attn_forward(qkv.data_ptr(), B, H, T, E, impl, p, 4242, 3, out.data_ptr(), lse.data_ptr(), scratch.data_ptr(), st) == 0
    assert lib.v1t_attn_backward(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), lse.data_ptr(), B, H, T, E, impl, p, 4242, 3,
                                 d_qkv.data_ptr(), scratch.data_ptr(), st) == 0, _lib.last_error()
    torch.cuda.synchronize()
    return d_qkv

for (B, H, T, E, p) in [(1, 1, 64, 32, 0.0), (1, 1, 200, 155, 0.0), (1, 1, 333, 24, 0.0), (1, 1, 128, 155, 0.0), (1, 1, 333, 155, 0.0),
                        (1, 1, 333, 64, 0.0), (1, 1, 333, 96, 0.0), (1, 1, 333, 128, 0.0), (1, 2, 1654, 155, 0.0)]:
    a = run(B, H, T, E, p, "three")
    b = run(B, H, T, E, p, "pair")
    I = H * E
    print(f"== B{B} H{H} T{T} E{E} p{p}")
    for name, sl in (("dq", slice(0, I)), ("dk", slice(I, 2 * I)), ("dv", slice(2 * I, 3 * I))):
        x, y = a[..., sl], b[..., sl]
        bad = ~torch.isfinite(y)
        diff = (x - y).abs()
        diff[bad] = 0
        rows = bad.any(-1).nonzero()
        cols = bad.any(0).any(0).nonzero().flatten()
        big = (diff > 1e-3 * x.abs().max()).any(-1).nonzero()
        print(f"  {name}: nonfinite {int(bad.sum())} rows[{rows[:, 1].min().item() if len(rows) else '-'}..{rows[:, 1].max().item() if len(rows) else '-'}] "
              f"cols[{cols.min().item() if len(cols) else '-'}..{cols.max().item() if len(cols) else '-'}] maxdiff {diff.max().item():.3e} (ref max {x.abs().max().item():.3e}) "
              f"rows with big diff: {len(big)} [{big[:, 1].min().item() if len(big) else '-'}..{big[:, 1].max().item() if len(big) else '-'}]")

print("---- detail: dk per (row block of 64, column atom of 32) max |diff| / ref max")
for (B, H, T, E, p) in [(1, 1, 64, 64, 0.0), (1, 1, 128, 64, 0.0), (1, 1, 200, 64, 0.0)]:
    a = run(B, H, T, E, p, "three")
    b = run(B, H, T, E, p, "pair")
    I = H * E
    x, y = a[0, :, I:2 * I], b[0, :, I:2 * I]
    y = torch.nan_to_num(y, nan=1e30, posinf=1e30, neginf=-1e30)
    print(f"== T{T} E{E}")
    for r in range(0, T, 64):
        print("  rows", r, [f"{((x[r:r+64, c:c+32] - y[r:r+64, c:c+32]).abs().max() / x.abs().max()).item():.1e}" for c in range(0, E, 32)])
    for impl in (_lib.IMPL_BF16,):
        a = run(B, H, T, E, p, "three", impl)
        b = run(B, H, T, E, p, "pair", impl)
        x, y = a[0, :, I:2 * I], torch.nan_to_num(b[0, :, I:2 * I], nan=1e30, posinf=1e30, neginf=-1e30)
        for r in range(0, T, 64):
            print("  bf16 rows", r, [f"{((x[r:r+64, c:c+32] - y[r:r+64, c:c+32]).abs().max() / x.abs().max()).item():.1e}" for c in range(0, E, 32)])
