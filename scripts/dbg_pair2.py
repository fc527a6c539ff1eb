import os, sys, math
import torch
sys.path.insert(0, ".")
from v1t_b200 import _lib
lib = _lib.load()
DEV = "cuda:0"
B, H, T, E, p = 1, 1, 64, 64, 0.0
impl = _lib.IMPL_BF16
os.environ["V1T_ATTN_BWD"] = "pair"
g = torch.Generator(device=DEV).manual_seed(B * 1000 + T + 1)
qkv = torch.randn(B, T, 3 * H * E, device=DEV, generator=g)
d_out = torch.randn(B, T, H * E, device=DEV, generator=g)
out = torch.empty(B, T, H * E, device=DEV)
Tp = 128
lse = torch.zeros(B * H, Tp, device=DEV)
d_qkv = torch.full((B, T, 3 * H * E), float("nan"), device=DEV)
scratch = torch.empty(lib.v1t_attn_scratch_bytes(B, H, T, E), dtype=torch.uint8, device=DEV)
st = torch.cuda.current_stream().cuda_stream
assert lib.v1t_attn_forward(qkv.data_ptr(), B, H, T, E, impl, p, 4242, 3, out.data_ptr(), lse.data_ptr(), scratch.data_ptr(), st) == 0
assert lib.v1t_attn_backward(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), lse.data_ptr(), B, H, T, E, impl, p, 4242, 3,
                             d_qkv.data_ptr(), scratch.data_ptr(), st) == 0
torch.cuda.synchronize()
q, k, v = (qkv[0, :, i * E:(i + 1) * E].double() for i in range(3))
do = d_out[0].double()
sc = E ** -0.5
P = torch.softmax(q @ k.T * sc, -1)
dP = do @ v.T
delta = (P * dP).sum(-1, keepdim=True)
dS = P * (dP - delta)
dk = d_qkv[0, :, E:2 * E].double()
cands = {"dS^T q": dS.T @ q * sc, "dS^T do": dS.T @ do * sc, "dS^T k": dS.T @ k * sc, "dS^T v": dS.T @ v * sc, "P^T do": P.T @ do * sc,
         "P^T q": P.T @ q * sc}
for name, c in cands.items():
    for a0 in (0, 32):
        for b0 in (0, 32):
            e = (dk[:, a0:a0 + 32] - c[:, b0:b0 + 32]).abs().max() / c.abs().max()
            print(f"dk[:, {a0}:{a0+32}] vs {name}[:, {b0}:{b0+32}]: {e.item():.2e}")
# rows permuted?  compare atom 1 with the expected atom 1 under row shifts
exp = cands["dS^T q"]
for sh in (0, 8, 16, 32):
    e = (dk[:, 32:64] - torch.roll(exp[:, 32:64], sh, 0)).abs().max() / exp.abs().max()
    print("row shift", sh, f"{e.item():.2e}")
# per-column error of atom 1
print("per-col err atom1:", [(f"{((dk[:, c] - exp[:, c]).abs().max() / exp.abs().max()).item():.1e}") for c in range(32, 64)])
# k-step structure: partial sums over query groups of 16
for grp in range(4):
    part = dS[grp * 16:(grp + 1) * 16].T @ q[grp * 16:(grp + 1) * 16] * sc
    print("queries", grp * 16, "partial atom1 corr:", f"{torch.corrcoef(torch.stack([part[:, 32:64].flatten(), dk[:, 32:64].flatten()]))[0, 1].item():.3f}")
