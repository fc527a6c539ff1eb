"""Generate tests/golden/extras.npz from the LIVE reference and from torch itself (build container only).

    python scripts/make_golden_extras.py          # needs /root/reference (read-only)

Holds, for the callers either side of the hot path:
  rollout/*   random softmax attention stacks and the reference's attention_rollouts() of them
              (src/v1t/utils/attention_rollout.py:124-133)
  crop<k>/*   the reference ImageCropper (src/v1t/models/image_cropper.py) on seeded inputs: state dict, inputs,
              output images and grids, for crop scales / shift modes / resize / behaviour planes
  opt/*       torch.optim.AdamW(weight_decay=0) driven the way the reference drives it (train.py:71-80,217-223):
              data gradients + autograd of reg_scale * |p|.sum(), three steps, parameters and moments after each
  ens<k>/*    the ensemble OutputModule's arithmetic (ensemble.py:68-80): mean / Linear over the member stack + ELU1
  mlp<k>/*    nn.Sequential Linear/ELU/Tanh stacks (gaussian2d.py:102-136, core_shifter.py:24-40): output and
              autograd weight gradients
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "extras.npz")


def rollout_cases(out):
    _, _, ar = rh.import_reference()
    g = torch.Generator().manual_seed(4321)
    # (B, L, H, gh, gw, image_shape): even and odd token counts, one block, default 4-block/4-head stack
    cases = {"a": (2, 3, 2, 5, 7, (12, 16)), "b": (3, 1, 3, 4, 3, (9, 11)), "c": (2, 4, 4, 6, 11, (36, 64)),
             "d": (1, 2, 1, 3, 3, (5, 5))}
    for k, (B, L, H, gh, gw, shape) in cases.items():
        T = gh * gw + 1
        logits = torch.randn((B, L, H, T, T), generator=g) * 2.0
        attn = torch.softmax(logits, dim=-1)
        heat = ar.attention_rollouts(attn, image_shape=list(shape))
        assert ar.find_shape(T - 1) == (gh, gw), (ar.find_shape(T - 1), gh, gw)
        out[f"rollout/{k}/attn"] = attn.numpy()
        out[f"rollout/{k}/heat"] = heat.numpy()
        out[f"rollout/{k}/grid"] = np.array([gh, gw])


def cropper_cases(out):
    rh.import_reference()
    from v1t.models.image_cropper import ImageCropper  # type: ignore

    g = torch.Generator().manual_seed(99)
    cases = {
        # name: (input_shape, B, overrides)
        "0": ((1, 144, 256), 2, dict(shift_mode=2, center_crop=1.0, resize_image=1, behavior_mode=3)),
        "1": ((1, 20, 30), 3, dict(shift_mode=3, center_crop=0.8, resize_image=0, behavior_mode=1,
                                    cropper_reg_scale=0.01)),
        "2": ((2, 40, 60), 2, dict(shift_mode=4, center_crop=0.7, resize_image=1, behavior_mode=4,
                                    cropper_reg_scale=0.0)),
        "3": ((1, 36, 64), 2, dict(shift_mode=1, center_crop=0.5, resize_image=0, behavior_mode=0)),
    }
    for k, (shape, B, over) in cases.items():
        args = rh.make_args({"A": 5}, in_shape=shape, **over)
        torch.manual_seed(7 + int(k))
        ds = rh.make_fake_ds({"A": 5})
        crop = ImageCropper(args, ds=ds)
        if crop.image_shifter is not None:
            with torch.no_grad():
                for p in crop.image_shifter.parameters():
                    p.mul_(4.0)
        images = torch.randn((B,) + shape, generator=g)
        behaviors = torch.rand((B, 3), generator=g)
        pupil = torch.rand((B, 2), generator=g) * 2 - 1
        with torch.no_grad():
            o, grid = crop(images, mouse_id="A", behaviors=behaviors, pupil_centers=pupil)
        for name, v in crop.state_dict().items():
            out[f"crop{k}/sd/{name}"] = v.numpy()
        out[f"crop{k}/images"] = images.numpy()
        out[f"crop{k}/behaviors"] = behaviors.numpy()
        out[f"crop{k}/pupil_centers"] = pupil.numpy()
        out[f"crop{k}/out"] = o.numpy()
        out[f"crop{k}/grid"] = grid.numpy()
        out[f"crop{k}/meta"] = np.array(repr(dict(in_shape=list(shape), B=B, over=over,
                                                  output_shape=list(crop.output_shape))))


def optimizer_case(out):
    g = torch.Generator().manual_seed(5)
    shapes = [(7,), (33, 5), (4, 1, 1030), (1, 4099)]  # unaligned tails, one tensor spanning two chunks
    l1 = [0.0, 0.5379 * 2, 0.0076, 0.3]
    lrs = [1e-3, 1e-3, 1.6e-3, 1.6e-3]
    params = [nn.Parameter(torch.randn(s, generator=g) * 0.5) for s in shapes]
    with torch.no_grad():
        params[1].view(-1)[::7] = 0.0  # sign(0) = 0
    opt = torch.optim.AdamW([{"params": params[:2], "lr": lrs[0]}, {"params": params[2:]}], lr=lrs[2],
                            betas=(0.9, 0.9999), eps=1e-8, weight_decay=0)
    for i, p in enumerate(params):
        out[f"opt/p0/{i}"] = p.detach().clone().numpy()
    out["opt/l1"] = np.array(l1)
    out["opt/lr"] = np.array(lrs)
    out["opt/hyper"] = np.array([0.9, 0.9999, 1e-8])
    for step in range(1, 4):
        opt.zero_grad()
        grads = [torch.randn(s, generator=g) * 0.1 for s in shapes]
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        reg = sum(c * p.abs().sum() for c, p in zip(l1, params))
        reg.backward()  # accumulates c * sign(p) like train.py:71-73
        opt.step()
        for i, p in enumerate(params):
            out[f"opt/g{step}/{i}"] = grads[i].numpy()
            out[f"opt/p{step}/{i}"] = p.detach().clone().numpy()
            out[f"opt/m{step}/{i}"] = opt.state[p]["exp_avg"].clone().numpy()
            out[f"opt/v{step}/{i}"] = opt.state[p]["exp_avg_sq"].clone().numpy()


def mlp_cases(out):
    g = torch.Generator().manual_seed(11)
    cases = {
        "0": (1000, 3, [(2, 30, "elu"), (30, 2, "tanh")], 2),           # grid predictor on source_grid[:, :2]
        "1": (16, 2, [(2, 5, "tanh"), (5, 5, "tanh"), (5, 2, "tanh")], 2),  # core shifter
        "2": (131, 5, [(5, 10, "tanh"), (10, 10, "tanh"), (10, 2, "tanh")], 5),  # image shifter, shift_mode 4
        "3": (1, 4, [(3, 32, "elu")], 3),                                # single row / single layer / max width
    }
    for k, (rows, x_cols, layers, used) in cases.items():
        mods = []
        for i, o, a in layers:
            mods.append(nn.Linear(i, o))
            mods.append(nn.ELU() if a == "elu" else nn.Tanh())
        net = nn.Sequential(*mods)
        with torch.no_grad():
            for p in net.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.8)
        x = torch.randn((rows, x_cols), generator=g)
        y = net(x[:, :used])
        dy = torch.randn(y.shape, generator=g)
        (y * dy).sum().backward()
        out[f"mlp{k}/x"] = x.numpy()
        out[f"mlp{k}/used"] = np.array(used)
        out[f"mlp{k}/acts"] = np.array([a for _, _, a in layers])
        out[f"mlp{k}/y"] = y.detach().numpy()
        out[f"mlp{k}/dy"] = dy.numpy()
        lin = [m for m in net if isinstance(m, nn.Linear)]
        for i, m in enumerate(lin):
            out[f"mlp{k}/w{i}"] = m.weight.detach().numpy()
            out[f"mlp{k}/b{i}"] = m.bias.detach().numpy()
            out[f"mlp{k}/gw{i}"] = m.weight.grad.numpy()
            out[f"mlp{k}/gb{i}"] = m.bias.grad.numpy()


def ensemble_cases(out):
    """OutputModule arithmetic (ensemble.py:68-80) restated with the same torch ops: mean(dim=-1) | nn.Linear(K, 1)
    on the [B,N,K] stack, then ELU + 1; gradients of the Linear by autograd."""
    g = torch.Generator().manual_seed(23)
    for k, (B, N, K, linear) in {"0": (3, 41, 5, False), "1": (2, 1000, 5, True), "2": (4, 17, 1, True)}.items():
        xs = [torch.randn((B, N), generator=g) * 1.5 for _ in range(K)]
        stack = torch.cat([x[..., None] for x in xs], dim=-1)
        if linear:
            lin = nn.Linear(K, 1)
            with torch.no_grad():
                lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * 0.5)
                lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.2)
            z = lin(stack)[..., 0]
        else:
            z = torch.mean(stack, dim=-1)
        y = torch.nn.functional.elu(z) + 1
        out[f"ens{k}/x"] = torch.stack(xs).numpy()
        out[f"ens{k}/y"] = y.detach().numpy()
        if linear:
            dy = torch.randn(y.shape, generator=g)
            (y * dy).sum().backward()
            out[f"ens{k}/dy"] = dy.numpy()
            out[f"ens{k}/w"] = lin.weight.detach().numpy()
            out[f"ens{k}/b"] = lin.bias.detach().numpy()
            out[f"ens{k}/gw"] = lin.weight.grad.numpy()
            out[f"ens{k}/gb"] = lin.bias.grad.numpy()


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    out = {}
    rollout_cases(out)
    cropper_cases(out)
    optimizer_case(out)
    mlp_cases(out)
    ensemble_cases(out)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({os.path.getsize(OUT) / 1e6:.2f} MB, {len(out)} arrays)")
