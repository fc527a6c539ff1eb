"""Time the fused attention forward / backward through the C-ABI (CUDA events), default V1T shape."""
import sys

import torch

sys.path.insert(0, ".")
from v1t_b200 import _lib

lib = _lib.load()
DEV = "cuda:0"
impl_name = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
B, H, T, E = 16, 4, 1654, 155
impl = _lib.IMPL_NAMES[impl_name]
qkv = torch.randn(B, T, 3 * H * E, device=DEV)
d_out = torch.randn(B, T, H * E, device=DEV)
out = torch.empty(B, T, H * E, device=DEV)
Tp = (T + 127) // 128 * 128
lse = torch.zeros(B * H, Tp, device=DEV)
d_qkv = torch.empty(B, T, 3 * H * E, device=DEV)
scratch = torch.empty(lib.v1t_attn_scratch_bytes(B, H, T, E), dtype=torch.uint8, device=DEV)
st = torch.cuda.current_stream().cuda_stream
fwd = lambda: lib.v1t_attn_forward(qkv.data_ptr(), B, H, T, E, impl, p, 1, 1, out.data_ptr(), lse.data_ptr(),
                                   scratch.data_ptr(), st)
bwd = lambda: lib.v1t_attn_backward(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), lse.data_ptr(), B, H, T, E, impl,
                                    p, 1, 1, d_qkv.data_ptr(), scratch.data_ptr(), st)
fl = 4.0 * B * H * T * T * E
for name, fn, mult in (("forward (incl. planes)", fwd, 1.0), ("backward (incl. planes, delta)", bwd, 2.0)):
    for _ in range(2):
        assert fn() == 0, _lib.last_error()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{impl_name} p={p} {name:32s}: {ms * 1e3:9.1f} us   {mult * fl / ms / 1e9:7.1f} algorithmic TFLOP/s", flush=True)
