"""Readout backward alone at the bench shape (B = 16, N = 8000, 29 x 57 x 155), clustered or spread positions:
ncu --metrics gpu__time_duration.sum --csv python scripts/prof_readout_bwd.py [clustered|spread]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from v1t_b200 import functional as VF
dev = "cuda:0"
mode = sys.argv[1] if len(sys.argv) > 1 else "clustered"
rng = np.random.default_rng(0)
B, gh, gw, N, C = 16, 29, 57, 8000, 155
base = torch.randn(B, gh * gw + 1, 160, device=dev)
fmap = base[:, 1:, :C].unflatten(1, (gh, gw)).permute(0, 3, 1, 2).requires_grad_(True)
cu = lambda a: torch.from_numpy(a).to(dev)
mu = (rng.normal(0, 0.02, (N, 2)) if mode == "clustered" else rng.uniform(-1, 1, (N, 2))).astype(np.float32)
sigma = rng.uniform(-0.1, 0.1, (N, 2, 2)).astype(np.float32)
noise = rng.standard_normal((B, N, 2)).astype(np.float32)
feats = cu(rng.standard_normal((C, N)).astype(np.float32)).requires_grad_(True)
dz = cu(rng.standard_normal((B, N)).astype(np.float32))
for _ in range(3):
    fmap.grad = None
    z = VF.readout_forward(fmap, cu(mu), cu(sigma), cu(noise), None, feats, None)
    z.backward(dz)
torch.cuda.synchronize()
