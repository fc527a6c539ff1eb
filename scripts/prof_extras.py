"""Workload for the ncu capture of the HBM-bound callers (fused L1+AdamW, attention rollout, readout):
    ncu --set full --import-source on -k regex:'adamw_l1_kernel|rollout_step_kernel' -c 6 -o gpurun_out/extras \
        python scripts/prof_extras.py
Sizes = the bench workload: 11.4 M parameters (7 mice x 8000 neurons + core), rollout stack [4,4,4,1654,1654]."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import v1t_b200  # noqa: E402
from v1t_b200 import functional as VF  # noqa: E402
from v1t_b200.optim import FusedAdamWL1, l1_coefficients  # noqa: E402

dev = torch.device("cuda", 0)
neurons = bench.neuron_counts(7, 8000)
margs = bench.make_args(neurons, dev)
torch.manual_seed(0)
model = v1t_b200.Model(margs, ds=bench.make_ds(neurons)).to(dev)
for p in model.parameters():
    p.grad = torch.randn_like(p) * 1e-3
opt = FusedAdamWL1(model.get_parameters(core_lr=1e-3), lr=1e-3, betas=(0.9, 0.9999), eps=1e-8, weight_decay=0,
                   l1=l1_coefficients(model, list(neurons)))
for _ in range(3):
    opt.step()
attn = torch.softmax(torch.randn((4, 4, 4, 1654, 1654), device=dev) * 2.0, dim=-1)
for _ in range(2):
    heat = VF.attention_rollouts(attn, (36, 64), (29, 57))
torch.cuda.synchronize()
print("ok", float(heat.mean()))
