"""Timeline of the pair attention-backward kernel inside a real training step (operand-plane epilogue, the bench path):
python scripts/pair_trace_model.py  -- prints the first two items of cluster 0 of the step's last attention backward."""
import os, sys
import torch
sys.path.insert(0, ".")
import v1t_b200
from v1t_b200 import _lib
from bench import make_args, make_ds
lib = _lib.load(); diag = _lib.load_diag()
dev = torch.device("cuda", 0)
mice = {"A": 8000}
B, T = 16, 1654
args = make_args(mice, dev, impl="bf16x3")
torch.manual_seed(1)
model = v1t_b200.Model(args, ds=make_ds(mice)).to(dev)
crit = v1t_b200.get_criterion(args, ds=make_ds(mice))
model.train(True)
g = torch.Generator(device=dev).manual_seed(3)
batch = dict(image=torch.randn(B, 1, 36, 64, device=dev, generator=g), behavior=torch.rand(B, 3, device=dev, generator=g),
             pupil_center=torch.rand(B, 2, device=dev, generator=g), response=torch.rand(B, 8000, device=dev, generator=g) * 2)
trace = torch.zeros(2 * 2 * 32 * 8, dtype=torch.int64, device=dev)
for it in range(3):
    if it == 2:
        assert diag.v1t_diag_attn_pair_trace(trace.data_ptr()) == 0
    model.zero_grad(set_to_none=True)
    y, _, _ = model(batch["image"], mouse_id="A", behaviors=batch["behavior"], pupil_centers=batch["pupil_center"])
    crit(y_true=batch["response"], y_pred=y, mouse_id="A", batch_size=B).backward()
torch.cuda.synchronize()
diag.v1t_diag_attn_pair_trace(None)
t = trace.view(2, 2, 32, 8).cpu()
names = ["mma:S wait", "mma:S issue", "mma:O wait", "mma:O issue", "sm:S arrived", "sm:xchg ok", "sm:slot free", "sm:A written"]
nt = (T + 63) // 64
for k in range(2):
    for r in range(2):
        print(f"== item {k} rank {r}")
        print("tile " + " ".join(f"{n:>13s}" for n in names))
        for j in list(range(3)) + list(range(nt - 2, nt)):
            print(f"{j:4d} " + " ".join(f"{int(t[k, r, j, e]):13d}" for e in range(8)))
        x = t[k, r, 31]
        print(f"   steady-state cycles per tile: {(t[k, r, 20, 1] - t[k, r, 4, 1]).item() / 16:.0f};  hand-over: accumulator complete "
              f"{int(x[0])}, epilogue done {int(x[1])}, resident stored {int(x[2])}, visible to the MMA warp {int(x[3])}, first accumulator chunk read {int(x[4])}")
