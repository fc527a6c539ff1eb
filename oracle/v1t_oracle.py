"""CPU oracle for the V1T hot path (TEST INFRASTRUCTURE — never imported by the product).

A float64 numpy restatement, forward AND hand-derived backward, of the reference
path that v1t_b200 re-implements in CUDA:

  ViT core      /root/reference/src/v1t/models/core/vit.py:41-129 (Image2Patches, mode 0),
                :157-202 (BehaviorMLP), :205-284 (Attention), :132-154 (MLP),
                :348-362 (Transformer.forward), :423-436 (ViTCore.forward)
  readout       /root/reference/src/v1t/models/readout/gaussian2d.py:188-278
  activation    /root/reference/src/v1t/models/utils.py:109-118 (ELU1)
  loss          /root/reference/src/v1t/losses.py:114-119,141-166 (PoissonLoss + scale_ds)
  shifter       /root/reference/src/v1t/models/core_shifter.py:24-40

All arithmetic of the path lives in PyTorch ATen (third-party; reference pins
"PyTorch 2.0", README.md:71-74; this container has torch 2.11).  The published
semantics restated here: nn.Unfold (c,kh,kw ordering), nn.LayerNorm (biased var,
eps 1e-5), exact-erf GELU, softmax over the last dim, F.grid_sample(bilinear,
zeros padding, align_corners=True) as in ATen/native/GridSampler.h:25-60.

PINNING: the reference ships no tests / golden vectors, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build
container: tests/golden/*.npz are produced by scripts/make_golden.py importing
/root/reference/src, and tests/test_oracle.py checks this file against them
(forward, loss and every gradient).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.  Parameters are addressed by the reference's state-dict keys
 so a reference ``Model.state_dict()`` plugs in directly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

try:  # scipy is in the image; fall back to a vectorised math.erf otherwise
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

F64 = np.float64
EPS_F32 = float(np.finfo(np.float32).eps)  # losses.py:22


@dataclass
class CoreConfig:
    """Shape-defining args of ViTCore (vit.py:374-405)."""

    in_ch: int = 1
    in_h: int = 36
    in_w: int = 64
    patch_size: int = 8
    patch_stride: int = 1
    emb_dim: int = 155
    num_heads: int = 4
    mlp_dim: int = 488
    num_blocks: int = 4
    behavior_mode: int = 3
    use_bias: bool = True

    @property
    def grid_hw(self):
        # Image2Patches.unfold_dim (vit.py:112-115)
        gh = (self.in_h - self.patch_size) // self.patch_stride + 1
        gw = (self.in_w - self.patch_size) // self.patch_stride + 1
        return gh, gw

    @property
    def num_tokens(self):
        gh, gw = self.grid_hw
        return gh * gw + 1

    @property
    def out_hw(self):
        # ViTCore.find_shape (vit.py:411-417)
        n = self.num_tokens - 1
        d1 = math.ceil(math.sqrt(n))
        while n % d1 != 0 and d1 > 0:
            d1 -= 1
        return d1, n // d1


# --------------------------------------------------------------------------- #
# small helpers
# --------------------------------------------------------------------------- #
def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=F64)


def params_to_f64(sd, prefix=""):
    """state-dict (torch or numpy values) -> {key: float64 ndarray}, optional prefix strip."""
    out = {}
    for k, v in sd.items():
        if prefix and not k.startswith(prefix):
            continue
        out[k[len(prefix):]] = _np(v)
    return out


def layer_norm_fwd(x, g, b, eps=1e-5):
    mean = x.mean(-1, keepdims=True)
    var = ((x - mean) ** 2).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * rstd
    return xhat * g + b, (xhat, rstd)


def layer_norm_bwd(dy, g, cache):
    xhat, rstd = cache
    red = tuple(range(dy.ndim - 1))
    dg = (dy * xhat).sum(red)
    db = dy.sum(red)
    dxh = dy * g
    dx = rstd * (dxh - dxh.mean(-1, keepdims=True) - xhat * (dxh * xhat).mean(-1, keepdims=True))
    return dx, dg, db


def gelu_fwd(h):
    return 0.5 * h * (1.0 + _erf(h / math.sqrt(2.0)))


def gelu_grad(h):
    return 0.5 * (1.0 + _erf(h / math.sqrt(2.0))) + h * np.exp(-0.5 * h * h) / math.sqrt(2.0 * math.pi)


def unfold_patches(img, p, s):
    """nn.Unfold + 'b c l -> b l c' (vit.py:68-71): patch vector ordered (c, kh, kw)."""
    B, C, Hh, Ww = img.shape
    gh = (Hh - p) // s + 1
    gw = (Ww - p) // s + 1
    out = np.empty((B, gh * gw, C * p * p), dtype=F64)
    for r in range(gh):
        for c in range(gw):
            out[:, r * gw + c, :] = img[:, :, r * s:r * s + p, c * s:c * s + p].reshape(B, -1)
    return out


# --------------------------------------------------------------------------- #
# ViT core forward / backward
# --------------------------------------------------------------------------- #
def _blk(i, name):
    return f"transformer.blocks.{i}.{name}"


def behavior_latent(P, cfg, i, b_in, mouse_id="share"):
    """BehaviorMLP.forward (vit.py:200-202): tanh(L2(tanh(L1(b))))."""
    key = mouse_id if cfg.behavior_mode == 4 else "share"
    W0 = P[_blk(i, f"b-mlp.models.{key}.0.weight")]
    W3 = P[_blk(i, f"b-mlp.models.{key}.3.weight")]
    b0 = P.get(_blk(i, f"b-mlp.models.{key}.0.bias"), 0.0)
    b3 = P.get(_blk(i, f"b-mlp.models.{key}.3.bias"), 0.0)
    h = np.tanh(b_in @ W0.T + b0)
    o = np.tanh(h @ W3.T + b3)
    return o, (h, o, key)


def core_forward(P, cfg: CoreConfig, images, behaviors, pupil_centers, mouse_id="A", masks=None):
    """ViTCore.forward (vit.py:423-436), eval mode / dropout 0 unless `masks` given.

    P: core params keyed like the reference state-dict WITHOUT the ``core.`` prefix.
    masks (optional): dict of explicit inverted-dropout multipliers
        {"tokens": [B,T,E], (i,"attn"): [B,H,T,T], (i,"proj"): [B,T,E],
         (i,"mlp1"): [B,T,M], (i,"mlp2"): [B,T,E]}  (values 0 or 1/(1-p)).
    Returns (fmap [B, gh, gw, E] channel-last, cache).
    """
    masks = masks or {}
    images = _np(images)
    B = images.shape[0]
    E, H = cfg.emb_dim, cfg.num_heads
    T = cfg.num_tokens
    cache = {"cfg": cfg, "B": B, "blocks": [], "mouse_id": mouse_id}

    patches = unfold_patches(images, cfg.patch_size, cfg.patch_stride)
    Wpe = P["patch_embedding.projection.2.weight"]
    bpe = P["patch_embedding.projection.2.bias"]
    tok = patches @ Wpe.T + bpe
    cls = np.broadcast_to(P["patch_embedding.cls_token"].reshape(1, 1, E), (B, 1, E))
    x = np.concatenate([cls, tok], axis=1) + P["patch_embedding.pos_embedding"][None]
    if "tokens" in masks:
        x = x * masks["tokens"]
    cache["patches"] = patches

    if cfg.behavior_mode in (3, 4):
        b_in = np.concatenate([_np(behaviors), _np(pupil_centers)], axis=-1)  # vit.py:431-432
    elif cfg.behavior_mode == 2:
        b_in = _np(behaviors)
    else:
        b_in = None
    cache["b_in"] = b_in

    scale = float(E) ** -0.5  # vit.py:234
    for i in range(cfg.num_blocks):
        c = {}
        if b_in is not None:
            bl, c["bmlp"] = behavior_latent(P, cfg, i, b_in, mouse_id)
            x = x + bl[:, None, :]  # vit.py:356-359 (persists in the residual stream)
        # --- Attention.mha (vit.py:267-275)
        h, c["ln1"] = layer_norm_fwd(x, P[_blk(i, "mha.layer_norm.weight")], P[_blk(i, "mha.layer_norm.bias")])
        c["h_ln1"] = h
        qkv = h @ P[_blk(i, "mha.to_qkv.weight")].T
        q, k, v = np.split(qkv, 3, axis=-1)
        to_heads = lambda a: a.reshape(B, T, H, E).transpose(0, 2, 1, 3)  # b n (h d) -> b h n d
        q, k, v = to_heads(q), to_heads(k), to_heads(v)
        s = np.einsum("bhid,bhjd->bhij", q, k) * scale
        s = s - s.max(-1, keepdims=True)
        p = np.exp(s)
        p = p / p.sum(-1, keepdims=True)
        pd = p * masks[(i, "attn")] if (i, "attn") in masks else p
        o = np.einsum("bhij,bhjd->bhid", pd, v)
        o2 = o.transpose(0, 2, 1, 3).reshape(B, T, H * E)
        a = o2 @ P[_blk(i, "mha.projection.0.weight")].T
        if _blk(i, "mha.projection.0.bias") in P:
            a = a + P[_blk(i, "mha.projection.0.bias")]
        if (i, "proj") in masks:
            a = a * masks[(i, "proj")]
        c.update(q=q, k=k, v=v, p=p, pd=pd, o2=o2)
        x = a + x
        # --- MLP (vit.py:143-154)
        h2, c["ln2"] = layer_norm_fwd(x, P[_blk(i, "mlp.model.0.weight")], P[_blk(i, "mlp.model.0.bias")])
        c["h_ln2"] = h2
        u = h2 @ P[_blk(i, "mlp.model.1.weight")].T
        if _blk(i, "mlp.model.1.bias") in P:
            u = u + P[_blk(i, "mlp.model.1.bias")]
        g = gelu_fwd(u)
        gd = g * masks[(i, "mlp1")] if (i, "mlp1") in masks else g
        m = gd @ P[_blk(i, "mlp.model.4.weight")].T
        if _blk(i, "mlp.model.4.bias") in P:
            m = m + P[_blk(i, "mlp.model.4.bias")]
        if (i, "mlp2") in masks:
            m = m * masks[(i, "mlp2")]
        c.update(u=u, gd=gd)
        x = m + x
        cache["blocks"].append(c)

    gh, gw = cfg.out_hw
    fmap = x[:, 1:, :].reshape(B, gh, gw, E)  # drop CLS; 'b (h w) c' (vit.py:434-435); no final LN
    cache["masks"] = masks
    return fmap, cache


def core_backward(P, cache, dfmap):
    """Gradient of core_forward w.r.t. every core parameter (and images).

    dfmap: [B, gh, gw, E].  Returns (grads dict keyed like P, dimages).
    """
    cfg: CoreConfig = cache["cfg"]
    masks = cache["masks"]
    B, E, H, T = cache["B"], cfg.emb_dim, cfg.num_heads, cfg.num_tokens
    scale = float(E) ** -0.5
    G = {}
    dx = np.zeros((B, T, E), dtype=F64)
    dx[:, 1:, :] = _np(dfmap).reshape(B, T - 1, E)
    b_in = cache["b_in"]

    def acc(key, val):
        G[key] = G.get(key, 0.0) + val

    for i in reversed(range(cfg.num_blocks)):
        c = cache["blocks"][i]
        # MLP branch: x_out = m + x
        dm = dx * masks[(i, "mlp2")] if (i, "mlp2") in masks else dx
        W2 = P[_blk(i, "mlp.model.4.weight")]
        acc(_blk(i, "mlp.model.4.weight"), np.einsum("bte,btm->em", dm, c["gd"]))
        if _blk(i, "mlp.model.4.bias") in P:
            acc(_blk(i, "mlp.model.4.bias"), dm.sum((0, 1)))
        dgd = dm @ W2
        dg = dgd * masks[(i, "mlp1")] if (i, "mlp1") in masks else dgd
        du = dg * gelu_grad(c["u"])
        W1 = P[_blk(i, "mlp.model.1.weight")]
        acc(_blk(i, "mlp.model.1.weight"), np.einsum("btm,bte->me", du, c["h_ln2"]))
        if _blk(i, "mlp.model.1.bias") in P:
            acc(_blk(i, "mlp.model.1.bias"), du.sum((0, 1)))
        dh2 = du @ W1
        dxl, dgm, dbt = layer_norm_bwd(dh2, P[_blk(i, "mlp.model.0.weight")], c["ln2"])
        acc(_blk(i, "mlp.model.0.weight"), dgm)
        acc(_blk(i, "mlp.model.0.bias"), dbt)
        dx = dx + dxl
        # attention branch: x_mid = a + x
        da = dx * masks[(i, "proj")] if (i, "proj") in masks else dx
        Wp = P[_blk(i, "mha.projection.0.weight")]
        acc(_blk(i, "mha.projection.0.weight"), np.einsum("bte,bti->ei", da, c["o2"]))
        if _blk(i, "mha.projection.0.bias") in P:
            acc(_blk(i, "mha.projection.0.bias"), da.sum((0, 1)))
        do2 = da @ Wp
        do = do2.reshape(B, T, H, E).transpose(0, 2, 1, 3)
        dv = np.einsum("bhij,bhid->bhjd", c["pd"], do)
        dpd = np.einsum("bhid,bhjd->bhij", do, c["v"])
        dp = dpd * masks[(i, "attn")] if (i, "attn") in masks else dpd
        ds = c["p"] * (dp - (dp * c["p"]).sum(-1, keepdims=True))
        dq = np.einsum("bhij,bhjd->bhid", ds, c["k"]) * scale
        dk = np.einsum("bhij,bhid->bhjd", ds, c["q"]) * scale
        from_heads = lambda a: a.transpose(0, 2, 1, 3).reshape(B, T, H * E)
        dqkv = np.concatenate([from_heads(dq), from_heads(dk), from_heads(dv)], axis=-1)
        Wqkv = P[_blk(i, "mha.to_qkv.weight")]
        acc(_blk(i, "mha.to_qkv.weight"), np.einsum("btn,bte->ne", dqkv, c["h_ln1"]))
        dh = dqkv @ Wqkv
        dxl, dgm, dbt = layer_norm_bwd(dh, P[_blk(i, "mha.layer_norm.weight")], c["ln1"])
        acc(_blk(i, "mha.layer_norm.weight"), dgm)
        acc(_blk(i, "mha.layer_norm.bias"), dbt)
        dx = dx + dxl
        # behaviour add
        if b_in is not None:
            hb, ob, key = c["bmlp"]
            dbl = dx.sum(1)  # [B,E]
            dz3 = dbl * (1.0 - ob * ob)
            W3 = P[_blk(i, f"b-mlp.models.{key}.3.weight")]
            acc(_blk(i, f"b-mlp.models.{key}.3.weight"), dz3.T @ hb)
            if _blk(i, f"b-mlp.models.{key}.3.bias") in P:
                acc(_blk(i, f"b-mlp.models.{key}.3.bias"), dz3.sum(0))
            dz0 = (dz3 @ W3) * (1.0 - hb * hb)
            acc(_blk(i, f"b-mlp.models.{key}.0.weight"), dz0.T @ b_in)
            if _blk(i, f"b-mlp.models.{key}.0.bias") in P:
                acc(_blk(i, f"b-mlp.models.{key}.0.bias"), dz0.sum(0))

    if "tokens" in masks:
        dx = dx * masks["tokens"]
    G["patch_embedding.pos_embedding"] = dx.sum(0)
    G["patch_embedding.cls_token"] = dx[:, 0, :].sum(0).reshape(1, 1, E)
    dtok = dx[:, 1:, :]
    G["patch_embedding.projection.2.weight"] = np.einsum("ble,blk->ek", dtok, cache["patches"])
    G["patch_embedding.projection.2.bias"] = dtok.sum((0, 1))
    # d images (fold of dpatches) — used by the reference's aRF scripts (misc/estimate_aRFs.py)
    dpatch = dtok @ P["patch_embedding.projection.2.weight"]
    p, s = cfg.patch_size, cfg.patch_stride
    gh, gw = cfg.grid_hw
    dimg = np.zeros((B, cfg.in_ch, cfg.in_h, cfg.in_w), dtype=F64)
    for r in range(gh):
        for cc in range(gw):
            dimg[:, :, r * s:r * s + p, cc * s:cc * s + p] += dpatch[:, r * gw + cc, :].reshape(B, cfg.in_ch, p, p)
    return G, dimg


# --------------------------------------------------------------------------- #
# Gaussian2d readout + ELU1 + Poisson
# --------------------------------------------------------------------------- #
def mu_transform_fwd(R, source_grid):
    """Gaussian2DReadout.mu with the grid predictor (gaussian2d.py:102-136,188-193)."""
    W0, b0 = R["mu_transform.0.weight"], R["mu_transform.0.bias"]
    W2, b2 = R["mu_transform.2.weight"], R["mu_transform.2.bias"]
    a = source_grid @ W0.T + b0
    h = np.where(a > 0, a, np.expm1(a))  # nn.ELU
    o = np.tanh(h @ W2.T + b2)
    return o, (source_grid, a, h, o)


def mu_transform_bwd(R, cache, dmu):
    sg, a, h, o = cache
    dz = dmu * (1 - o * o)
    G = {"mu_transform.2.weight": dz.T @ h, "mu_transform.2.bias": dz.sum(0)}
    dh = dz @ R["mu_transform.2.weight"]
    da = dh * np.where(a > 0, 1.0, np.exp(a))
    G["mu_transform.0.weight"] = da.T @ sg
    G["mu_transform.0.bias"] = da.sum(0)
    return G


def shifter_fwd(S, pupil_centers):
    """CoreShifter.forward (core_shifter.py:24-40): Linear/Tanh x3."""
    x = _np(pupil_centers)
    acts = [x]
    for j in (0, 2, 4):
        x = np.tanh(x @ S[f"mlp.{j}.weight"].T + S[f"mlp.{j}.bias"])
        acts.append(x)
    return x, acts


def shifter_bwd(S, acts, dshift):
    G = {}
    d = dshift
    for idx, j in reversed(list(enumerate((0, 2, 4)))):
        y = acts[idx + 1]
        dz = d * (1 - y * y)
        G[f"mlp.{j}.weight"] = dz.T @ acts[idx]
        G[f"mlp.{j}.bias"] = dz.sum(0)
        d = dz @ S[f"mlp.{j}.weight"]
    return G


def readout_forward(fmap, mu, sigma, features, bias, noise=None, shifts=None):
    """Gaussian2DReadout.forward (gaussian2d.py:237-278), pre-activation z [B,N].

    fmap     [B, gh, gw, C] channel-last feature map
    mu       [N,2] (x,y);  sigma [N,2,2];  features [C,N];  bias [N] or None
    noise    [B,N,2] standard normal (train mode, gaussian2d.py:219-235) or None (eval)
    shifts   [B,2] or None (added AFTER the clamp, gaussian2d.py:267-268)
    """
    fmap = _np(fmap)
    B, gh, gw, C = fmap.shape
    N = mu.shape[0]
    pre = np.broadcast_to(mu[None], (B, N, 2)).copy()
    if noise is not None:
        pre = pre + np.einsum("ncd,bnd->bnc", sigma, _np(noise))
    grid = np.clip(pre, -1.0, 1.0)
    if shifts is not None:
        grid = grid + _np(shifts)[:, None, :]
    ix = (grid[..., 0] + 1) * 0.5 * (gw - 1)
    iy = (grid[..., 1] + 1) * 0.5 * (gh - 1)
    x0 = np.floor(ix).astype(np.int64)
    y0 = np.floor(iy).astype(np.int64)
    samp = np.zeros((B, N, C), dtype=F64)
    corners = []
    bidx = np.arange(B)[:, None]
    for dy in (0, 1):
        for dx in (0, 1):
            xc, yc = x0 + dx, y0 + dy
            wx = 1.0 - np.abs(ix - xc)
            wy = 1.0 - np.abs(iy - yc)
            valid = ((xc >= 0) & (xc <= gw - 1) & (yc >= 0) & (yc <= gh - 1)).astype(F64)
            xcc, ycc = np.clip(xc, 0, gw - 1), np.clip(yc, 0, gh - 1)
            val = fmap[bidx, ycc, xcc, :]  # [B,N,C]
            samp += (valid * wx * wy)[..., None] * val
            corners.append((dx, dy, xcc, ycc, wx, wy, valid, val))
    z = np.einsum("bnc,cn->bn", samp, features)
    if bias is not None:
        z = z + bias
    cache = dict(fmap_shape=fmap.shape, pre=pre, corners=corners, samp=samp, noise=None if noise is None else _np(noise))
    return z, cache


def readout_backward(cache, sigma, features, dz):
    """Backward of readout_forward (verified against autograd)."""
    B, gh, gw, C = cache["fmap_shape"]
    N = features.shape[1]
    samp = cache["samp"]
    G = {"bias": dz.sum(0), "features": np.einsum("bn,bnc->cn", dz, samp)}
    dfmap = np.zeros((B, gh, gw, C), dtype=F64)
    gx = np.zeros((B, N), dtype=F64)
    gy = np.zeros((B, N), dtype=F64)
    featT = features.T  # [N,C]
    bidx = np.broadcast_to(np.arange(B)[:, None], (B, N))
    for (dx, dy, xcc, ycc, wx, wy, valid, val) in cache["corners"]:
        contrib = (dz * valid * wx * wy)[..., None] * featT[None]  # [B,N,C]
        np.add.at(dfmap, (bidx, ycc, xcc), contrib)
        dot = np.einsum("bnc,nc->bn", val, featT)
        sx = 1.0 if dx == 1 else -1.0
        sy = 1.0 if dy == 1 else -1.0
        gx += valid * sx * wy * dot
        gy += valid * sy * wx * dot
    dgrid = np.stack([gx * dz * (gw - 1) * 0.5, gy * dz * (gh - 1) * 0.5], axis=-1)  # [B,N,2]
    G["shifts"] = dgrid.sum(1)
    pre = cache["pre"]
    dpre = dgrid * ((pre >= -1.0) & (pre <= 1.0))
    G["mu"] = dpre.sum(0)
    if cache["noise"] is not None:
        G["sigma"] = np.einsum("bnc,bnd->ncd", dpre, cache["noise"])
    else:
        G["sigma"] = np.zeros_like(sigma)
    return G, dfmap


def elu1_fwd(z):
    return np.where(z > 0, z, np.expm1(np.minimum(z, 0.0))) + 1.0


def poisson_loss_fwd(y_pred, y_true, ds_size, batch_size, eps=EPS_F32, ds_scale=True):
    """PoissonLoss.forward (losses.py:153-166) incl. Loss.scale_ds (:114-119)."""
    yp, yt = y_pred + eps, _np(y_true) + eps
    loss = (yp - yt * np.log(yp)).sum()
    s = math.sqrt(ds_size / batch_size) if ds_scale else 1.0
    return s * loss, s


def elu1_poisson_bwd(z, y_true, s, eps=EPS_F32, dloss=1.0):
    """dL/dz for loss = s * sum((y+eps) - (yt+eps) log(y+eps)),  y = elu(z)+1."""
    y = elu1_fwd(z)
    dy = dloss * s * (1.0 - (_np(y_true) + eps) / (y + eps))
    return dy * np.where(z > 0, 1.0, np.exp(np.minimum(z, 0.0)))


# --------------------------------------------------------------------------- #
# whole path: Model.forward (model.py:151-177) minus the cropper, + loss, + grads
# --------------------------------------------------------------------------- #
def path_forward_backward(sd, cfg: CoreConfig, mouse_id, images, behaviors, pupil_centers, y_true,
                          ds_size, batch_size=None, noise=None, masks=None, want_grads=True):
    """Full hot path on a reference ``Model.state_dict()`` (keys core.*, readouts.<id>.*, core_shifter.<id>.*).

    Returns dict(fmap, z, y, loss, grads{full state-dict key: ndarray}, dimages).
    """
    P = params_to_f64(sd, "core.")
    R = params_to_f64(sd, f"readouts.{mouse_id}.")
    S = params_to_f64(sd, f"core_shifter.{mouse_id}.")
    B = _np(images).shape[0]
    batch_size = B if batch_size is None else batch_size
    fmap, ccache = core_forward(P, cfg, images, behaviors, pupil_centers, mouse_id, masks)
    N = R["bias"].shape[0] if "bias" in R else R["features"].shape[-1]
    if "_mu" in R:
        mu, mcache = R["_mu"].reshape(N, 2), None
    else:
        mu, mcache = mu_transform_fwd(R, R["source_grid"])
    sigma = R["sigma"].reshape(N, 2, 2)
    C = cfg.emb_dim
    features = R["features"].reshape(C, N)
    shifts, sacts = (None, None)
    if S:
        shifts, sacts = shifter_fwd(S, pupil_centers)
    z, rcache = readout_forward(fmap, mu, sigma, features, R.get("bias"), noise=noise, shifts=shifts)
    y = elu1_fwd(z)
    loss, s = poisson_loss_fwd(y, y_true, ds_size, batch_size)
    out = dict(fmap=fmap, z=z, y=y, loss=loss, mu=mu, shifts=shifts)
    if not want_grads:
        return out
    dz = elu1_poisson_bwd(z, y_true, s)
    RG, dfmap = readout_backward(rcache, sigma, features, dz)
    grads = {}
    grads[f"readouts.{mouse_id}.features"] = RG["features"].reshape(R["features"].shape)
    if "bias" in R:
        grads[f"readouts.{mouse_id}.bias"] = RG["bias"]
    grads[f"readouts.{mouse_id}.sigma"] = RG["sigma"].reshape(R["sigma"].shape)
    if mcache is not None:
        for k, v in mu_transform_bwd(R, mcache, RG["mu"]).items():
            grads[f"readouts.{mouse_id}.{k}"] = v
    else:
        grads[f"readouts.{mouse_id}._mu"] = RG["mu"].reshape(R["_mu"].shape)
    if S:
        for k, v in shifter_bwd(S, sacts, RG["shifts"]).items():
            grads[f"core_shifter.{mouse_id}.{k}"] = v
    CG, dimg = core_backward(P, ccache, dfmap)
    for k, v in CG.items():
        grads["core." + k] = v
    out.update(grads=grads, dimages=dimg, dz=dz, dfmap=dfmap, dmu=RG["mu"], dshifts=RG["shifts"])
    return out
