"""Generate tests/golden/*.npz by running the LIVE reference in the build container.

    python scripts/make_golden.py            # needs /root/reference (read-only)

Each fixture holds: the reference Model.state_dict(), seeded synthetic inputs, the
injected readout noise (SURVEY.md Appendix C "train-mode parity recipe"), and the
reference's outputs — core feature map, pre-activation z, responses y, Poisson loss and
the gradient of the loss w.r.t. every parameter (fp32, as the reference computes them).
The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (neurons, in_shape, B, overrides, mode)
    "tiny_eval": ({"A": 37}, (1, 12, 16), 3, dict(emb_dim=24, num_heads=2, mlp_dim=40, num_blocks=2), "eval"),
    "tiny_train": ({"A": 37}, (1, 12, 16), 3, dict(emb_dim=24, num_heads=2, mlp_dim=40, num_blocks=2), "train"),
    "color_mode4": ({"K": 29, "L": 41}, (2, 11, 13), 4,
                    dict(emb_dim=20, num_heads=3, mlp_dim=36, num_blocks=2, behavior_mode=4, patch_size=4,
                         patch_stride=2, ds_name="franke2022"), "train"),
    "nobias_mu_param": ({"A": 33}, (1, 12, 16), 2,
                        dict(emb_dim=16, num_heads=2, mlp_dim=24, num_blocks=1, disable_bias=True,
                             disable_grid_predictor=True, behavior_mode=2, shift_mode=0), "train"),
    "nobehav": ({"A": 21}, (1, 10, 12), 2,
                dict(emb_dim=16, num_heads=1, mlp_dim=24, num_blocks=1, behavior_mode=0), "eval"),
    # default widths (E=155,H=4,M=488: exercises the 155->160 padding paths), short sequence
    "default_dims": ({"A": 200}, (1, 12, 20), 2, dict(num_blocks=1), "train"),
}


def run_case(name, neurons, in_shape, B, over, mode, seed=1234):
    over = dict(over)
    over.update(p_dropout=0.0, t_dropout=0.0)  # deterministic (Appendix C)
    args = rh.make_args(neurons, in_shape=in_shape, **over)
    ds = rh.make_fake_ds(neurons, ds_size=4500, seed=seed)
    model, crit = rh.build_reference_model(args, ds, seed=seed, trained_like=True)
    g = torch.Generator().manual_seed(seed + 7)
    out = {}
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()
    model.train(mode == "train")
    for mouse_id, n in neurons.items():
        images = torch.randn((B,) + tuple(in_shape), generator=g)
        behaviors = torch.rand((B, 3), generator=g)
        pupil = torch.rand((B, 2), generator=g)
        y_true = torch.rand((B, n), generator=g) * 2.0
        model.zero_grad(set_to_none=True)
        noise = None
        if mode == "train":
            torch.manual_seed(seed + 13)
            noise = torch.empty(B, n, 1, 2).normal_()
            torch.manual_seed(seed + 13)  # the reference draws the same tensor (gaussian2d.py:221)
        images.requires_grad_(True)
        # step-by-step Model.forward (model.py:151-177) so the core map can be captured
        imgs, _ = model.image_cropper(images, mouse_id=mouse_id, behaviors=behaviors, pupil_centers=pupil)
        fmap = model.core(imgs, mouse_id=mouse_id, behaviors=behaviors, pupil_centers=pupil)
        shifts = None
        if model.core_shifter is not None:
            shifts = model.core_shifter(pupil, mouse_id=mouse_id)
        z = model.readouts(fmap, mouse_id=mouse_id, shifts=shifts)
        y = model.elu1(z)
        loss = crit(y_true=y_true, y_pred=y, mouse_id=mouse_id, batch_size=B)
        loss.backward()
        pfx = f"{mouse_id}/"
        out[pfx + "images"] = images.detach().numpy()
        out[pfx + "behaviors"] = behaviors.numpy()
        out[pfx + "pupil_centers"] = pupil.numpy()
        out[pfx + "y_true"] = y_true.numpy()
        if noise is not None:
            out[pfx + "noise"] = noise.view(B, n, 2).numpy()
        out[pfx + "fmap"] = fmap.detach().permute(0, 2, 3, 1).contiguous().numpy()  # [B,gh,gw,E]
        out[pfx + "z"] = z.detach().numpy()
        out[pfx + "y"] = y.detach().numpy()
        out[pfx + "loss"] = loss.detach().numpy()
        out[pfx + "dimages"] = images.grad.numpy()
        for k, p in model.named_parameters():
            if p.grad is not None:
                out[pfx + "grad/" + k] = p.grad.detach().clone().numpy()
    meta = dict(neurons=neurons, in_shape=list(in_shape), B=B, mode=mode, ds_size=4500,
                args={k: v for k, v in vars(args).items() if isinstance(v, (int, float, str, bool))})
    out["meta"] = np.array(repr(meta))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for name, (neurons, in_shape, B, over, mode) in CASES.items():
        run_case(name, neurons, in_shape, B, over, mode)
