"""Callers either side of the hot path: fused L1 + AdamW, small MLPs, image cropper, attention
rollout.  CPU part: the numpy oracle against tests/golden/extras.npz (outputs of the live reference / of torch,
scripts/make_golden_extras.py).  GPU part: the CUDA kernels through the C-ABI against the same fixtures and oracle."""
import ast
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR, rel_err
from oracle import extras_oracle as XO

Z = np.load(os.path.join(GOLDEN_DIR, "extras.npz"))
ROLLOUT_CASES = sorted({k.split("/")[1] for k in Z.files if k.startswith("rollout/")})
CROP_CASES = sorted({k.split("/")[0][4:] for k in Z.files if k.startswith("crop")})
MLP_CASES = sorted({k.split("/")[0][3:] for k in Z.files if k.startswith("mlp")})
ENS_CASES = sorted({k.split("/")[0][3:] for k in Z.files if k.startswith("ens")})
N_OPT = len([k for k in Z.files if k.startswith("opt/p0/")])

TOL_F32 = 2e-6   # fp32 kernels against fp64 oracle / fp32 torch, max|d| / max|ref|
TOL_ROLL = 2e-5  # rollout: T-long fp32 dot products chained over the blocks, then min-max normalised
TOL_ROLL_FULL = 1e-4  # same at T ~ 1654-2014 (absolute error on a [0,1]-normalised heat map)


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def _mlp(k):
    n = len([f for f in Z.files if f.startswith(f"mlp{k}/w")])
    w = [Z[f"mlp{k}/w{i}"] for i in range(n)]
    b = [Z[f"mlp{k}/b{i}"] for i in range(n)]
    acts = [str(a) for a in Z[f"mlp{k}/acts"]]
    return w, b, acts, int(Z[f"mlp{k}/used"])


def _crop_meta(k):
    return ast.literal_eval(str(Z[f"crop{k}/meta"]))


def _crop_shifts(k):
    """ImageShifter output recomputed with the oracle MLP from the fixture's state dict (image_cropper.py:40-48)."""
    sd = {f[len(f"crop{k}/sd/"):]: Z[f] for f in Z.files if f.startswith(f"crop{k}/sd/")}
    if not any(n.startswith("image_shifter") for n in sd):
        return None, sd
    over = _crop_meta(k)["over"]
    x = Z[f"crop{k}/pupil_centers"]
    if over["shift_mode"] == 4:
        x = np.concatenate([Z[f"crop{k}/behaviors"], x], axis=1)
    w = [sd[f"image_shifter.A.mlp.{i}.weight"] for i in (0, 2, 4)]
    b = [sd[f"image_shifter.A.mlp.{i}.bias"] for i in (0, 2, 4)]
    y, _ = XO.small_mlp_forward(x, w, b, ["tanh"] * 3)
    return y * float(sd["image_shifter.A.max_shift"]), sd


# ---------------------------------------------------------------------------------------------------
# CPU: oracle pinned to the reference / torch fixtures
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", ROLLOUT_CASES)
def test_oracle_rollout_matches_reference(k):
    attn, heat = Z[f"rollout/{k}/attn"], Z[f"rollout/{k}/heat"]
    assert XO.find_shape(attn.shape[-1] - 1) == tuple(Z[f"rollout/{k}/grid"])
    got = XO.attention_rollouts(attn, heat.shape[1:])
    assert rel_err(got, heat) < 1e-5


@pytest.mark.parametrize("k", CROP_CASES)
def test_oracle_cropper_matches_reference(k):
    meta = _crop_meta(k)
    shifts, sd = _crop_shifts(k)
    grid0 = sd["grid"][0]
    beh = Z[f"crop{k}/behaviors"] if meta["over"]["behavior_mode"] == 1 else None
    got = XO.crop_resize(Z[f"crop{k}/images"], grid0, shifts, meta["output_shape"][1:], beh)
    ref = Z[f"crop{k}/out"]
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-6
    if shifts is not None:
        assert rel_err(grid0[None] + shifts[:, None, None, :], Z[f"crop{k}/grid"]) < 1e-6


def test_oracle_adamw_l1_matches_torch():
    b1, b2, eps = Z["opt/hyper"]
    for i in range(N_OPT):
        p = Z[f"opt/p0/{i}"].astype(np.float64)
        m, v = np.zeros_like(p), np.zeros_like(p)
        for step in (1, 2, 3):
            p, m, v = XO.adamw_l1_step(p, Z[f"opt/g{step}/{i}"], m, v, step, Z["opt/lr"][i], b1, b2, eps,
                                       l1=Z["opt/l1"][i])
            assert rel_err(p, Z[f"opt/p{step}/{i}"]) < 1e-6
            assert rel_err(m, Z[f"opt/m{step}/{i}"]) < 1e-6
            assert rel_err(v, Z[f"opt/v{step}/{i}"]) < 1e-6


@pytest.mark.parametrize("k", MLP_CASES)
def test_oracle_small_mlp_matches_torch(k):
    w, b, acts, used = _mlp(k)
    x = Z[f"mlp{k}/x"][:, :used]
    y, cache = XO.small_mlp_forward(x, w, b, acts)
    assert rel_err(y, Z[f"mlp{k}/y"]) < 1e-6
    gw, gb = XO.small_mlp_backward(Z[f"mlp{k}/dy"], w, b, acts, cache)
    for i in range(len(w)):
        assert rel_err(gw[i], Z[f"mlp{k}/gw{i}"]) < 1e-5
        assert rel_err(gb[i], Z[f"mlp{k}/gb{i}"]) < 1e-5


@pytest.mark.parametrize("k", ENS_CASES)
def test_oracle_ensemble_matches_torch(k):
    x = Z[f"ens{k}/x"]
    linear = f"ens{k}/w" in Z.files
    y, z = XO.ensemble_combine(x, Z[f"ens{k}/w"] if linear else None, Z[f"ens{k}/b"] if linear else None)
    assert rel_err(y, Z[f"ens{k}/y"]) < 1e-6
    if linear:
        gw, gb = XO.ensemble_backward(x, z, Z[f"ens{k}/dy"])
        assert rel_err(gw, Z[f"ens{k}/gw"]) < 1e-5 and rel_err(gb, Z[f"ens{k}/gb"]) < 1e-5


def test_optimizer_refuses_cpu_parameters():
    from v1t_b200.optim import FusedAdamWL1

    p = torch.nn.Parameter(torch.ones(4))
    p.grad = torch.ones(4)
    opt = FusedAdamWL1([p], lr=1e-3)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        opt.step()


def test_small_mlp_and_rollout_refuse_cpu_tensors():
    from v1t_b200 import functional as VF

    lin = torch.nn.Linear(2, 3)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        VF.small_mlp(torch.zeros(4, 2), [lin], ["tanh"])
    with pytest.raises(RuntimeError, match="CUDA-only"):
        VF.attention_rollouts(torch.zeros(1, 1, 1, 5, 5), (4, 4), (2, 2))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        VF.crop_resize(torch.zeros(1, 1, 4, 4), torch.zeros(1, 4, 4, 2), None, (4, 4))


# ---------------------------------------------------------------------------------------------------
# GPU: the CUDA kernels through the C-ABI
# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("k", ROLLOUT_CASES)
def test_rollout_kernel_matches_reference(k):
    from v1t_b200 import functional as VF

    attn, heat = Z[f"rollout/{k}/attn"], Z[f"rollout/{k}/heat"]
    got = VF.attention_rollouts(cu(attn), heat.shape[1:], tuple(Z[f"rollout/{k}/grid"]))
    assert rel_err(got.cpu().numpy(), heat) < TOL_ROLL
    assert rel_err(got.cpu().numpy(), XO.attention_rollouts(attn, heat.shape[1:])) < TOL_ROLL


@pytest.mark.gpu
@pytest.mark.parametrize("T,B", [(1654, 2), (2014, 1), (515, 3)])
def test_rollout_kernel_full_size_against_oracle(T, B):
    """Default sequence length (and the other register-tile instantiations) against the matrix-chain oracle."""
    from v1t_b200 import functional as VF

    g = torch.Generator(device="cuda").manual_seed(T)
    L, H = 3, 2
    attn = torch.softmax(torch.randn((B, L, H, T, T), generator=g, device="cuda") * 3.0, dim=-1)
    gh, gw = XO.find_shape(T - 1)
    got = VF.attention_rollouts(attn, (36, 64), (gh, gw)).cpu().numpy()
    ref = XO.attention_rollouts(attn.cpu().numpy(), (36, 64))
    # min-max normalisation divides by the map's range, which for random attention is a fraction of its mean, so the
    # fp32 rounding of the T-term sums is amplified accordingly (the reference's fp32 matmuls have the same property)
    assert rel_err(got, ref) < TOL_ROLL_FULL
    assert np.isfinite(got).all() and got.min() >= 0.0 and got.max() <= 1.0 + 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("k", CROP_CASES)
def test_cropper_module_matches_reference(k):
    from v1t_b200.modules import ImageCropper

    meta = _crop_meta(k)
    a = dict(input_shape=tuple(meta["in_shape"]), ds_name="sensorium", cropper_reg_scale=0.0)
    a.update(meta["over"])
    crop = ImageCropper(SimpleNamespace(**a), ds={"A": None})
    sd = {f[len(f"crop{k}/sd/"):]: torch.from_numpy(Z[f]) for f in Z.files if f.startswith(f"crop{k}/sd/")}
    crop.load_state_dict(sd, strict=True)
    crop.cuda()
    assert tuple(crop.output_shape) == tuple(meta["output_shape"])
    out, grid = crop(cu(Z[f"crop{k}/images"]), mouse_id="A", behaviors=cu(Z[f"crop{k}/behaviors"]),
                     pupil_centers=cu(Z[f"crop{k}/pupil_centers"]))
    assert rel_err(grid.detach().cpu().numpy(), Z[f"crop{k}/grid"]) < TOL_F32
    ref = Z[f"crop{k}/out"]
    assert out.shape == ref.shape
    got = out.cpu().numpy()
    # nearest sampling is discontinuous: a shift computed in fp32 on the GPU may round a tie differently from the
    # reference's CPU fp32.  Allow a handful of pixels to differ, the rest must agree to fp32 accuracy.
    bad = np.abs(got - ref) > 1e-5 * max(1.0, np.abs(ref).max())
    assert bad.mean() < 2e-3, f"{bad.sum()} of {bad.size} pixels differ"


@pytest.mark.gpu
@pytest.mark.parametrize("k", MLP_CASES)
def test_small_mlp_kernels_match_torch(k):
    from v1t_b200 import functional as VF

    w, b, acts, used = _mlp(k)
    x = cu(Z[f"mlp{k}/x"])
    wt = [cu(a).requires_grad_(True) for a in w]
    bt = [cu(a).requires_grad_(True) for a in b]
    xin = x[:, :used]  # a strided view when the predictor reads only the first columns (gaussian2d.py:111)
    y = VF.small_mlp(xin, list(zip(wt, bt)), acts)
    assert rel_err(y.detach().cpu().numpy(), Z[f"mlp{k}/y"]) < TOL_F32
    (y * cu(Z[f"mlp{k}/dy"])).sum().backward()
    for i in range(len(w)):
        assert rel_err(wt[i].grad.cpu().numpy(), Z[f"mlp{k}/gw{i}"]) < 1e-5
        assert rel_err(bt[i].grad.cpu().numpy(), Z[f"mlp{k}/gb{i}"]) < 1e-5


@pytest.mark.gpu
def test_small_mlp_backward_is_deterministic_and_handles_missing_bias():
    from v1t_b200 import functional as VF

    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((8003, 2), generator=g, device="cuda")
    w0 = torch.randn((30, 2), generator=g, device="cuda").requires_grad_(True)
    w1 = torch.randn((2, 30), generator=g, device="cuda").requires_grad_(True)
    dy = torch.randn((8003, 2), generator=g, device="cuda")
    grads = []
    for _ in range(2):
        w0.grad = w1.grad = None
        y = VF.small_mlp(x, [(w0, None), (w1, None)], ["elu", "tanh"])
        (y * dy).sum().backward()
        grads.append((w0.grad.clone(), w1.grad.clone()))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
    yo, cache = XO.small_mlp_forward(x.cpu().numpy(), [w0.detach().cpu().numpy(), w1.detach().cpu().numpy()],
                                     [None, None], ["elu", "tanh"])
    gw, _ = XO.small_mlp_backward(dy.cpu().numpy(), [w0.detach().cpu().numpy(), w1.detach().cpu().numpy()],
                                  [None, None], ["elu", "tanh"], cache)
    assert rel_err(y.detach().cpu().numpy(), yo) < TOL_F32
    assert rel_err(grads[0][0].cpu().numpy(), gw[0]) < 1e-5 and rel_err(grads[0][1].cpu().numpy(), gw[1]) < 1e-5


def _fused_optimizer():
    from v1t_b200.optim import FusedAdamWL1

    b1, b2, eps = (float(v) for v in Z["opt/hyper"])
    params = [torch.nn.Parameter(cu(Z[f"opt/p0/{i}"])) for i in range(N_OPT)]
    lrs = Z["opt/lr"]
    opt = FusedAdamWL1([{"params": params[:2], "lr": float(lrs[0])}, {"params": params[2:]}], lr=float(lrs[2]),
                       betas=(b1, b2), eps=eps, weight_decay=0,
                       l1={p: (float(c), i % 3) for i, (p, c) in enumerate(zip(params, Z["opt/l1"]))})
    return opt, params


@pytest.mark.gpu
def test_fused_adamw_l1_matches_torch_adamw():
    opt, params = _fused_optimizer()
    for step in (1, 2, 3):
        before = [p.detach().abs().sum().item() for p in params]
        for i, p in enumerate(params):
            p.grad = cu(Z[f"opt/g{step}/{i}"])
        opt.step(zero_grad=(step == 2))
        for i, p in enumerate(params):
            assert rel_err(p.detach().cpu().numpy(), Z[f"opt/p{step}/{i}"]) < TOL_F32, (step, i)
            assert rel_err(opt.state[p]["exp_avg"].cpu().numpy(), Z[f"opt/m{step}/{i}"]) < TOL_F32
            assert rel_err(opt.state[p]["exp_avg_sq"].cpu().numpy(), Z[f"opt/v{step}/{i}"]) < TOL_F32
            assert float(opt.state[p]["step"]) == step
            if step == 2:
                assert float(p.grad.abs().max()) == 0.0
        sums = opt.last_l1_sums.cpu().numpy()
        want = np.zeros(opt.n_l1_groups)
        for i, s in enumerate(before):
            want[i % 3] += s
        assert rel_err(sums, want) < 1e-5


@pytest.mark.gpu
def test_fused_adamw_state_dict_interchanges_with_torch_adamw():
    opt, params = _fused_optimizer()
    for i, p in enumerate(params):
        p.grad = cu(Z[f"opt/g1/{i}"])
    opt.step()
    # reference-side optimizer (train.py:217-223) picks the state up and continues identically to torch-only
    tparams = [torch.nn.Parameter(p.detach().clone()) for p in params]
    topt = torch.optim.AdamW([{"params": tparams[:2], "lr": float(Z["opt/lr"][0])}, {"params": tparams[2:]}],
                             lr=float(Z["opt/lr"][2]), betas=tuple(float(v) for v in Z["opt/hyper"][:2]),
                             eps=float(Z["opt/hyper"][2]), weight_decay=0)
    topt.load_state_dict(opt.state_dict())
    for i, p in enumerate(tparams):
        p.grad = cu(Z[f"opt/g2/{i}"]) + float(Z["opt/l1"][i]) * torch.sign(p.detach())
    topt.step()
    for i, p in enumerate(tparams):
        assert rel_err(p.detach().cpu().numpy(), Z[f"opt/p2/{i}"]) < TOL_F32
    # and back: a torch.optim.AdamW state dict loads into the fused optimizer
    opt2, params2 = _fused_optimizer()
    with torch.no_grad():
        for p2, p in zip(params2, tparams):
            p2.copy_(p)
    opt2.load_state_dict(topt.state_dict())
    for i, p in enumerate(params2):
        p.grad = cu(Z[f"opt/g3/{i}"])
    opt2.step()
    for i, p in enumerate(params2):
        assert rel_err(p.detach().cpu().numpy(), Z[f"opt/p3/{i}"]) < TOL_F32


@pytest.mark.gpu
def test_fused_adamw_steps_regularised_parameter_without_gradient():
    from v1t_b200.optim import FusedAdamWL1

    p = torch.nn.Parameter(torch.tensor([0.5, -0.25, 0.0, 2.0], device="cuda"))
    q = torch.nn.Parameter(torch.ones(3, device="cuda"))  # no gradient, not regularised: untouched
    opt = FusedAdamWL1([p, q], lr=0.1, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, l1={p: 0.3})
    opt.step()
    ref, _, _ = XO.adamw_l1_step(np.array([0.5, -0.25, 0.0, 2.0]), np.zeros(4), np.zeros(4), np.zeros(4), 1, 0.1, 0.9,
                                 0.999, 1e-8, l1=0.3)
    assert rel_err(p.detach().cpu().numpy(), ref) < TOL_F32
    assert torch.equal(q.detach(), torch.ones(3, device="cuda")) and len(opt.state[q]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("k", ENS_CASES)
def test_ensemble_output_module_matches_torch(k):
    from v1t_b200 import functional as VF

    x = [cu(a) for a in Z[f"ens{k}/x"]]
    if f"ens{k}/w" not in Z.files:
        y = VF.ensemble_combine(x)
        assert rel_err(y.cpu().numpy(), Z[f"ens{k}/y"]) < TOL_F32
        return
    w, b = cu(Z[f"ens{k}/w"]).requires_grad_(True), cu(Z[f"ens{k}/b"]).requires_grad_(True)
    y = VF.ensemble_combine(x, w, b)
    assert rel_err(y.detach().cpu().numpy(), Z[f"ens{k}/y"]) < TOL_F32
    (y * cu(Z[f"ens{k}/dy"])).sum().backward()
    assert rel_err(w.grad.cpu().numpy(), Z[f"ens{k}/gw"]) < 1e-5
    assert rel_err(b.grad.cpu().numpy(), Z[f"ens{k}/gb"]) < 1e-5


@pytest.mark.gpu
def test_fused_adamw_unaligned_parameter_takes_the_scalar_path():
    """A parameter that is a 4-byte-aligned view of a flat buffer (not 16-byte aligned) must give the same update."""
    from v1t_b200.optim import FusedAdamWL1

    rng = np.random.default_rng(0)
    n = 4099 + 37
    p0, g0 = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    flat, gflat = torch.zeros(n + 1, device="cuda"), torch.zeros(n + 3, device="cuda")
    p = torch.nn.Parameter(flat[1:])
    with torch.no_grad():
        p.copy_(cu(p0))
    p.grad = gflat[3:]
    p.grad.copy_(cu(g0))
    assert p.data_ptr() % 16 != 0
    opt = FusedAdamWL1([p], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1, l1={p: 0.05})
    opt.step(grad_scale=0.5)
    want, m, v = XO.adamw_l1_step(p0, g0, np.zeros(n), np.zeros(n), 1, 1e-2, 0.9, 0.999, 1e-8, l1=0.05,
                                  weight_decay=0.1, grad_scale=0.5)
    assert rel_err(p.detach().cpu().numpy(), want) < TOL_F32
    assert rel_err(opt.state[p]["exp_avg"].cpu().numpy(), m) < TOL_F32


# ---------------------------------------------------------------------------------------------------
# CPU: size-independent properties of the oracle restatements (beyond the pinned fixtures)
# ---------------------------------------------------------------------------------------------------
def test_oracle_small_mlp_gradients_agree_with_finite_differences():
    rng = np.random.default_rng(4)
    w = [rng.standard_normal((6, 3)), rng.standard_normal((4, 6)), rng.standard_normal((2, 4))]
    b = [rng.standard_normal(6), None, rng.standard_normal(2)]
    acts = ["elu", "tanh", None]
    x, dy = rng.standard_normal((9, 3)), rng.standard_normal((9, 2))
    y, cache = XO.small_mlp_forward(x, w, b, acts)
    gw, gb = XO.small_mlp_backward(dy, w, b, acts, cache)
    assert gb[1] is None
    eps = 1e-6
    for l, idx in ((0, (2, 1)), (1, (3, 5)), (2, (1, 0))):
        wp = [a.copy() for a in w]
        wp[l][idx] += eps
        wm = [a.copy() for a in w]
        wm[l][idx] -= eps
        num = ((XO.small_mlp_forward(x, wp, b, acts)[0] - XO.small_mlp_forward(x, wm, b, acts)[0]) * dy).sum() / (2 * eps)
        assert abs(num - gw[l][idx]) < 1e-6 * max(1.0, abs(num))
    bp, bm = [None if a is None else a.copy() for a in b], [None if a is None else a.copy() for a in b]
    bp[0][3] += eps
    bm[0][3] -= eps
    num = ((XO.small_mlp_forward(x, w, bp, acts)[0] - XO.small_mlp_forward(x, w, bm, acts)[0]) * dy).sum() / (2 * eps)
    assert abs(num - gb[0][3]) < 1e-6 * max(1.0, abs(num))


def test_oracle_rollout_invariants():
    """Head order does not matter (max over heads); a stack whose blocks all attend uniformly gives a map that is
    constant before normalisation (0/0 -> NaN in the reference too), so use a one-hot perturbation instead: the
    perturbed patch is the arg-max of the map."""
    rng = np.random.default_rng(8)
    L, H, T = 3, 4, 21
    gh, gw = XO.find_shape(T - 1)  # (5, 4): the reference's own factorisation of the 20 patches
    a = rng.random((L, H, T, T)) + 0.1
    a /= a.sum(-1, keepdims=True)
    base = XO.attention_rollout(a, (gh, gw))
    assert np.allclose(XO.attention_rollout(a[:, ::-1], (gh, gw)), base)
    assert base.min() == 0.0 and base.max() == 1.0
    hot = np.full((L, H, T, T), 1.0 / T)
    hot[:, :, :, 7] += 0.5  # every token of every block looks harder at token 7 = patch 6
    hot /= hot.sum(-1, keepdims=True)
    heat = XO.attention_rollout(hot, (gh, gw))  # same size as the patch grid: the resize is the identity
    assert np.unravel_index(np.argmax(heat), heat.shape) == divmod(6, gw)
    # row-vector formulation used by the CUDA kernel == matrix chain of the reference
    m = a.max(1) + np.eye(T)
    m /= m.sum(-1, keepdims=True)
    r = m[-1][0]
    for n in range(L - 2, -1, -1):
        r = r @ m[n]
    chain = m[0]
    for n in range(1, L):
        chain = m[n] @ chain
    assert np.allclose(r, chain[0])


def test_oracle_cropper_identity_and_adamw_closed_form():
    rng = np.random.default_rng(2)
    img = rng.standard_normal((2, 3, 6, 9))
    ys, xs = np.linspace(-1, 1, 6, dtype=np.float32), np.linspace(-1, 1, 9, dtype=np.float32)
    grid = np.stack(np.meshgrid(xs, ys), axis=-1)  # (x, y) at [row, col]
    assert np.array_equal(XO.crop_resize(img, grid), img)
    far = XO.crop_resize(img, grid, shifts=np.array([[3.0, 0.0], [0.0, -3.0]]))
    assert np.all(far == 0.0)  # shifted fully outside: zero padding
    # first AdamW step without regulariser or decay moves every coordinate by lr * sign(g) (up to eps)
    p, g = rng.standard_normal(50), rng.standard_normal(50)
    p1, m1, v1 = XO.adamw_l1_step(p, g, np.zeros(50), np.zeros(50), 1, 0.01, 0.9, 0.999, 1e-12)
    assert np.allclose(p1, p - 0.01 * np.sign(g), atol=1e-9)
    assert np.allclose(m1, 0.1 * g) and np.allclose(v1, 0.001 * g * g)
    # the L1 term alone (zero data gradient) pulls towards zero and leaves exact zeros in place
    q = np.array([0.5, -0.25, 0.0])
    q1, _, _ = XO.adamw_l1_step(q, np.zeros(3), np.zeros(3), np.zeros(3), 1, 0.01, 0.9, 0.999, 1e-12, l1=0.3)
    assert np.allclose(q1, [0.49, -0.24, 0.0])
