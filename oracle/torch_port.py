"""CPU port of the reference hot path with the SAME ATen ops the reference calls (TEST/BASELINE INFRASTRUCTURE).

The reference is Python and cannot travel to the GPU box, so bench.py's ``cpu_baseline`` / ``--impl reference``
legs time this port instead (``kind: "port"``): nn.functional.unfold / linear / layer_norm / softmax / gelu /
grid_sample + autograd, fp32, all host threads — op for op what vit.py:122-129,253-275,348-362,
gaussian2d.py:195-278, models/utils.py:117-118 and losses.py:153-166 execute on a CPU device.
tests/test_oracle.py pins it against the golden fixtures produced by the live reference.
Never imported by the product.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

EPS = torch.finfo(torch.float32).eps


def core_forward(sd, cfg, images, behaviors, pupil_centers, mouse_id="A", p_drop=0.0, t_drop=0.0, training=False):
    """sd: state dict with 'core.' keys (tensors).  Returns the (B,E,h,w) channel-last VIEW like the reference."""
    E, H = cfg.emb_dim, cfg.num_heads
    B = images.shape[0]
    g = lambda k: sd["core." + k]
    x = F.unfold(images, kernel_size=cfg.patch_size, stride=cfg.patch_stride).transpose(1, 2)
    x = F.linear(x, g("patch_embedding.projection.2.weight"), g("patch_embedding.projection.2.bias"))
    x = torch.cat((g("patch_embedding.cls_token").expand(B, -1, -1), x), dim=1)
    x = x + g("patch_embedding.pos_embedding")
    x = F.dropout(x, p_drop, training)
    if cfg.behavior_mode in (3, 4):
        beh = torch.cat((behaviors, pupil_centers), dim=-1)
    elif cfg.behavior_mode == 2:
        beh = behaviors
    else:
        beh = None
    key = mouse_id if cfg.behavior_mode == 4 else "share"
    scale = E ** -0.5
    for i in range(cfg.num_blocks):
        p = f"transformer.blocks.{i}."
        opt = lambda k: sd.get("core." + p + k)
        if beh is not None:
            h = torch.tanh(F.linear(beh, g(p + f"b-mlp.models.{key}.0.weight"), opt(f"b-mlp.models.{key}.0.bias")))
            h = torch.tanh(F.linear(h, g(p + f"b-mlp.models.{key}.3.weight"), opt(f"b-mlp.models.{key}.3.bias")))
            x = x + h[:, None, :]
        h = F.layer_norm(x, (E,), g(p + "mha.layer_norm.weight"), g(p + "mha.layer_norm.bias"))
        q, k, v = torch.chunk(F.linear(h, g(p + "mha.to_qkv.weight")), 3, dim=-1)
        heads = lambda a: a.reshape(B, -1, H, E).transpose(1, 2)
        q, k, v = heads(q), heads(k), heads(v)
        attn = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * scale, dim=-1)
        attn = F.dropout(attn, t_drop, training)
        o = torch.matmul(attn, v).transpose(1, 2).reshape(B, -1, H * E)
        o = F.dropout(F.linear(o, g(p + "mha.projection.0.weight"), opt("mha.projection.0.bias")), t_drop, training)
        x = o + x
        h = F.layer_norm(x, (E,), g(p + "mlp.model.0.weight"), g(p + "mlp.model.0.bias"))
        h = F.dropout(F.gelu(F.linear(h, g(p + "mlp.model.1.weight"), opt("mlp.model.1.bias"))), t_drop, training)
        h = F.dropout(F.linear(h, g(p + "mlp.model.4.weight"), opt("mlp.model.4.bias")), t_drop, training)
        x = h + x
    gh, gw = cfg.out_hw
    return x[:, 1:, :].reshape(B, gh, gw, E).permute(0, 3, 1, 2)


def readout_forward(sd, mouse_id, fmap, shifts=None, noise=None):
    r = lambda k: sd[f"readouts.{mouse_id}.{k}"]
    B, C = fmap.shape[:2]
    feats = r("features")
    N = feats.shape[-1]
    if f"readouts.{mouse_id}._mu" in sd:
        mu = r("_mu")
    else:
        h = F.elu(F.linear(r("source_grid"), r("mu_transform.0.weight"), r("mu_transform.0.bias")))
        mu = torch.tanh(F.linear(h, r("mu_transform.2.weight"), r("mu_transform.2.bias"))).view(1, N, 1, 2)
    norm = noise.view(B, N, 1, 2) if noise is not None else mu.new_zeros(B, N, 1, 2)
    grid = torch.clamp(torch.einsum("ancd,bnid->bnic", r("sigma"), norm) + mu, min=-1, max=1)
    if shifts is not None:
        grid = grid + shifts[:, None, None, :]
    out = F.grid_sample(fmap, grid=grid, align_corners=True).squeeze(-1)
    out = (out * feats.view(1, C, N)).sum(dim=1)
    bias = sd.get(f"readouts.{mouse_id}.bias")
    return out + bias if bias is not None else out


def shifter_forward(sd, mouse_id, pupil_centers):
    if f"core_shifter.{mouse_id}.mlp.0.weight" not in sd:
        return None
    x = pupil_centers
    for j in (0, 2, 4):
        x = torch.tanh(F.linear(x, sd[f"core_shifter.{mouse_id}.mlp.{j}.weight"], sd[f"core_shifter.{mouse_id}.mlp.{j}.bias"]))
    return x


def step(sd, cfg, mouse_id, images, behaviors, pupil_centers, y_true, ds_size, batch_size=None, noise=None,
         p_drop=0.0, t_drop=0.0, training=False, core_reg_scale=None):
    """One forward+backward of core + readout + ELU1 + Poisson (+ optional core L1) — train_step's body
    (train.py:56-73).  Returns (loss, y).  Gradients land in the .grad of the tensors in ``sd`` that require grad."""
    B = images.shape[0]
    batch_size = B if batch_size is None else batch_size
    if training and noise is None:
        noise = torch.empty(B, sd[f"readouts.{mouse_id}.features"].shape[-1], 1, 2).normal_()
    fmap = core_forward(sd, cfg, images, behaviors, pupil_centers, mouse_id, p_drop, t_drop, training)
    z = readout_forward(sd, mouse_id, fmap, shifter_forward(sd, mouse_id, pupil_centers), noise)
    y = F.elu(z) + 1.0
    yp, yt = y + EPS, y_true + EPS
    loss = torch.sum(yp - yt * torch.log(yp)) * math.sqrt(ds_size / batch_size)
    total = loss
    if core_reg_scale is not None:
        total = total + (B / batch_size) * core_reg_scale * sum(
            v.abs().sum() for k, v in sd.items() if k.startswith("core.") and v.requires_grad)
    total.backward()
    return loss.detach(), y.detach()
