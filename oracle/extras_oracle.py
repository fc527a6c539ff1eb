"""CPU restatement (numpy, fp64) of the callers either side of the hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(v1t_b200/) never does.  Each function cites the reference lines it follows (paths under /root/reference); where the
arithmetic lives in PyTorch (absent from /root/reference: the reference pins only "PyTorch 2.0", README.md:71-74;
this image has torch 2.11) the published ATen / torch.optim algorithm is restated and named.

Parity is PINNED: tests/golden/extras.npz holds outputs of the live reference (attention_rollouts, ImageCropper) and
of torch itself (torch.optim.AdamW + autograd of the L1 term, nn.Sequential MLPs), written by
scripts/make_golden_extras.py; tests/test_extras.py checks every function here against them.
"""
from __future__ import annotations

import math

import numpy as np


# ---------------------------------------------------------------------------------------------------
# n1: L1 regulariser + AdamW
# ---------------------------------------------------------------------------------------------------
def adamw_l1_step(p, g, m, v, step, lr, beta1, beta2, eps, l1=0.0, weight_decay=0.0, grad_scale=1.0):
    """One optimizer step on one tensor; returns (p, m, v) new.  ``step`` is the 1-based step count.

    Gradient: data gradient + l1 * sign(p) — autograd of ``reg_scale * p.abs().sum()`` (vit.py:419-421,
    gaussian2d.py:83-100, core_shifter.py:35-36; added to the loss at train.py:71-73).
    Update: torch/optim/adamw.py::_single_tensor_adamw (amsgrad=False, maximize=False), the optimizer the reference
    builds at train.py:217-223:
        p *= 1 - lr * wd;  m = lerp(m, g, 1 - b1);  v = b2 v + (1 - b2) g^2
        p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
    """
    p, g, m, v = (np.asarray(a, dtype=np.float64) for a in (p, g, m, v))
    g = grad_scale * g + l1 * np.sign(p)
    p = p * (1.0 - lr * weight_decay)
    m = m + (g - m) * (1.0 - beta1)
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2_sqrt = math.sqrt(1.0 - beta2 ** step)
    p = p - (lr / bc1) * m / (np.sqrt(v) / bc2_sqrt + eps)
    return p, m, v


# ---------------------------------------------------------------------------------------------------
# a11 / n3: small MLPs (grid predictor gaussian2d.py:102-136,188-193; shifters core_shifter.py:24-40,
# image_cropper.py:27-48)
# ---------------------------------------------------------------------------------------------------
def _act(kind, x):
    if kind == "tanh":
        return np.tanh(x)
    if kind == "elu":  # nn.ELU(alpha=1)
        return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    return x


def _act_grad(kind, pre, y):
    if kind == "tanh":
        return 1.0 - y * y
    if kind == "elu":
        return np.where(pre > 0, 1.0, y + 1.0)
    return np.ones_like(pre)


def small_mlp_forward(x, weights, biases, acts):
    """x [R,in]; weights[l] [out,in] (nn.Linear layout), biases[l] [out] or None; acts[l] in {"tanh","elu",None}.
    Returns (y, cache)."""
    a = np.asarray(x, dtype=np.float64)
    cache = [a]
    pres = []
    for w, b, k in zip(weights, biases, acts):
        pre = a @ np.asarray(w, dtype=np.float64).T
        if b is not None:
            pre = pre + np.asarray(b, dtype=np.float64)
        a = _act(k, pre)
        pres.append(pre)
        cache.append(a)
    return a, (cache, pres)


def small_mlp_backward(dy, weights, biases, acts, cache):
    """Returns ([dW_l], [db_l]) (db_l None where the layer has no bias)."""
    acts_out, pres = cache
    d = np.asarray(dy, dtype=np.float64)
    gw, gb = [None] * len(weights), [None] * len(weights)
    for l in range(len(weights) - 1, -1, -1):
        dpre = d * _act_grad(acts[l], pres[l], acts_out[l + 1])
        gw[l] = dpre.T @ acts_out[l]
        gb[l] = dpre.sum(axis=0) if biases[l] is not None else None
        d = dpre @ np.asarray(weights[l], dtype=np.float64)
    return gw, gb


# ---------------------------------------------------------------------------------------------------
# n3: image cropper (image_cropper.py:104-112 build_grid, :120-140 forward)
# ---------------------------------------------------------------------------------------------------
def _nearest_sample(img, gx, gy):
    """F.grid_sample(mode="nearest", padding_mode="zeros", align_corners=True) for one image [C,H,W] on grid points
    (gx, gy) [h,w] (ATen GridSampler.h: unnormalise ((g+1)/2)(size-1), nearbyint = round-half-even, zeros outside).
    Coordinates are evaluated in fp32 like ATen so that ties resolve identically."""
    c, h, w = img.shape
    ix = ((gx.astype(np.float32) + np.float32(1)) / np.float32(2)) * np.float32(w - 1)
    iy = ((gy.astype(np.float32) + np.float32(1)) / np.float32(2)) * np.float32(h - 1)
    nx, ny = np.rint(ix).astype(np.int64), np.rint(iy).astype(np.int64)
    ok = (nx >= 0) & (nx <= w - 1) & (ny >= 0) & (ny <= h - 1)
    out = img[:, np.clip(ny, 0, h - 1), np.clip(nx, 0, w - 1)]
    return np.where(ok[None], out, 0.0)


def _bilinear_resize(x, oh, ow):
    """torchvision Resize(antialias=False) == F.interpolate(mode="bilinear", align_corners=False) on [..., H, W]
    (ATen UpSample.h: src = max(scale (dst + 0.5) - 0.5, 0), scale = in / out; taps i0, min(i0 + 1, in - 1))."""
    h, w = x.shape[-2:]

    def taps(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        src = np.maximum(scale * (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) - np.float32(0.5), 0)
        i0 = np.minimum(src.astype(np.int64), n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        lam = (src - i0.astype(np.float32)).astype(np.float64)
        return i0, i1, lam

    y0, y1, ly = taps(h, oh)
    x0, x1, lx = taps(w, ow)
    top = x[..., y0, :][..., :, x0] * (1 - lx) + x[..., y0, :][..., :, x1] * lx
    bot = x[..., y1, :][..., :, x0] * (1 - lx) + x[..., y1, :][..., :, x1] * lx
    return top * (1 - ly)[:, None] + bot * ly[:, None]


def crop_resize(images, grid, shifts=None, out_hw=None, behaviors=None):
    """images [B,C,H,W]; grid [crop_h,crop_w,2] (x,y); shifts [B,2] or None; out_hw None = no resize;
    behaviors [B,K] appended as constant planes (behavior_mode 1)."""
    images = np.asarray(images, dtype=np.float64)
    grid = np.asarray(grid, dtype=np.float32).reshape(grid.shape[-3], grid.shape[-2], 2)
    outs = []
    for b in range(images.shape[0]):
        gx, gy = grid[..., 0], grid[..., 1]
        if shifts is not None:
            gx = gx + np.float32(shifts[b][0])
            gy = gy + np.float32(shifts[b][1])
        o = _nearest_sample(images[b], gx, gy)
        if out_hw is not None and tuple(out_hw) != o.shape[-2:]:
            o = _bilinear_resize(o, out_hw[0], out_hw[1])
        if behaviors is not None:
            planes = np.broadcast_to(np.asarray(behaviors[b], dtype=np.float64)[:, None, None],
                                     (len(behaviors[b]),) + o.shape[-2:])
            o = np.concatenate([o, planes], axis=0)
        outs.append(o)
    return np.stack(outs)


# ---------------------------------------------------------------------------------------------------
# n2: attention rollout (attention_rollout.py:78-133)
# ---------------------------------------------------------------------------------------------------
def find_shape(num_patches: int):
    """attention_rollout.py:78-83."""
    dim1 = math.ceil(math.sqrt(num_patches))
    while num_patches % dim1 != 0 and dim1 > 0:
        dim1 -= 1
    return dim1, num_patches // dim1


def attention_rollout(attention, image_shape):
    """One sample, attention [L,H,T,T] -> heat map [*image_shape]; the matrix-chain form the reference uses
    (attention_rollout.py:92-121), deliberately NOT the row-vector form of the CUDA kernel."""
    a = np.asarray(attention, dtype=np.float64).max(axis=1)
    a = a + np.eye(a.shape[-1])
    a = a / a.sum(axis=-1, keepdims=True)
    joint = a[0]
    for n in range(1, a.shape[0]):
        joint = a[n] @ joint
    heat = joint[0, 1:]
    heat = heat.reshape(find_shape(len(heat)))
    heat = (heat - heat.min()) / (heat.max() - heat.min())
    return _bilinear_resize(heat, image_shape[0], image_shape[1])


def attention_rollouts(attentions, image_shape):
    """attention_rollout.py:124-133."""
    return np.stack([attention_rollout(a, image_shape) for a in attentions])


# ---------------------------------------------------------------------------------------------------
# n4: ensemble output module (ensemble.py:30-80 OutputModule, :131-151 EnsembleModel.forward)
# ---------------------------------------------------------------------------------------------------
def ensemble_combine(members, weight=None, bias=None):
    """members [K,B,N] pre-activation responses; weight [1,K] / bias [1] of nn.Linear(K,1), or None = mean
    (ensemble_mode 0).  Returns (y = elu(z) + 1, z)."""
    x = np.asarray(members, dtype=np.float64)
    if weight is None:
        z = x.mean(axis=0)
    else:
        z = np.tensordot(np.asarray(weight, dtype=np.float64).reshape(-1), x, axes=(0, 0))
        if bias is not None:
            z = z + float(np.asarray(bias).reshape(-1)[0])
    return np.where(z > 0, z, np.expm1(np.minimum(z, 0))) + 1.0, z


def ensemble_backward(members, z, dy):
    """Gradients of the Linear output module: (d_weight [1,K], d_bias [1])."""
    x = np.asarray(members, dtype=np.float64)
    dz = np.asarray(dy, dtype=np.float64) * np.where(z > 0, 1.0, np.exp(np.minimum(z, 0)))
    return np.tensordot(x, dz, axes=([1, 2], [0, 1]))[None, :], np.array([dz.sum()])
