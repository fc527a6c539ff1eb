"""GPU tests against the LIVE reference (oracle/_ref: the unmodified bryanlimy/V1T staged by oracle/make_ref.py).

What the golden fixtures cannot pin: (1) the real shape — default core (T = 1654 tokens, 4 blocks), batch 16,
8000 neurons — against the reference's own fp32 forward/backward run on the same box; (2) the drop-in seam: the
REFERENCE's ``Model.forward`` (model.py:151-177), ``train.train_step`` (train.py:42-81) and
``attention_rollout.Recorder`` (attention_rollout.py:15-75) executing with the B200 modules that
``dropin.install()`` put into its registries.  Skipped when oracle/_ref is absent.
"""
import numpy as np
import pytest
import torch

import v1t_b200
from v1t_b200 import dropin
from oracle import ref_harness as rh
from golden_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_FWD, TOL_GRAD = 1e-3, 1e-3  # north-star tolerance; the reference side runs fp32 (TF32 off) on the same GPU

needs_ref = pytest.mark.skipif(not rh.reference_available(), reason="oracle/_ref absent (run oracle/make_ref.py)")


def _pair(neurons, seed=1234, **over):
    """(reference Model, its criterion, the reference's Model CLASS rebuilt with the B200 modules installed, criterion)
    with identical weights, both on the GPU."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    rh.import_reference()
    import v1t.models.model as ref_model_mod
    import v1t.losses as ref_losses

    over.setdefault("p_dropout", 0.0)
    over.setdefault("t_dropout", 0.0)
    args = rh.make_args(neurons, device=torch.device(DEV), **over)
    ds = rh.make_fake_ds(neurons, ds_size=4500, seed=seed)
    ref, ref_crit = rh.build_reference_model(args, ds, seed=seed, trained_like=True)
    ref = ref.to(DEV)
    ref_crit = ref_crit.to(DEV)
    dropin.install()
    try:
        args2 = rh.make_args(neurons, device=torch.device(DEV), **over)
        new = ref_model_mod.Model(args2, ds=ds)  # the reference's own class, B200 modules inside
        assert isinstance(new.core, v1t_b200.ViTCore) and isinstance(new.readouts[next(iter(neurons))],
                                                                   v1t_b200.Gaussian2DReadout)
        new.load_state_dict(ref.state_dict(), strict=True)
        new = new.to(DEV)
        new_crit = ref_losses.get_criterion(args2, ds=ds)
        assert isinstance(new_crit, v1t_b200.PoissonLoss)
    finally:
        dropin.uninstall()
    return ref, ref_crit, new, new_crit


def _batch(n, B, seed=7, in_shape=(1, 36, 64)):
    g = torch.Generator().manual_seed(seed)
    return {"image": torch.randn((B,) + in_shape, generator=g).to(DEV), "behavior": torch.rand((B, 3), generator=g).to(DEV),
            "pupil_center": torch.rand((B, 2), generator=g).to(DEV), "response": (torch.rand((B, n), generator=g) * 2).to(DEV)}


def _fwd_bwd(model, crit, mouse, b, seed):
    model.zero_grad(set_to_none=True)
    torch.manual_seed(seed)  # train mode: both sides draw the readout position noise from the device generator
    y, _, _ = model(inputs=b["image"], mouse_id=mouse, behaviors=b["behavior"], pupil_centers=b["pupil_center"])
    loss = crit(y_true=b["response"], y_pred=y, mouse_id=mouse, batch_size=b["image"].shape[0])
    loss.backward()
    return y.detach(), float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


@needs_ref
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_full_shape_forward_backward_matches_live_reference(mode):
    """T = 1654, 4 blocks, B = 16, N = 8000 — the bench shape — values, not just invariants."""
    n = 8000
    ref, ref_crit, new, new_crit = _pair({"A": n})
    ref.train(mode == "train")
    new.train(mode == "train")
    b = _batch(n, 16)
    y0, l0, g0 = _fwd_bwd(ref, ref_crit, "A", b, 11)
    y1, l1, g1 = _fwd_bwd(new, new_crit, "A", b, 11)
    assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) < TOL_FWD
    assert abs(l1 - l0) / abs(l0) < TOL_FWD
    worst = ("", 0.0)
    for k, ref_g in g0.items():
        if float(ref_g.abs().max()) == 0.0:
            continue
        assert k in g1, k
        e = rel_err(g1[k].cpu().numpy(), ref_g.cpu().numpy())
        if e > worst[1]:
            worst = (k, e)
    print(f"[live-ref {mode}] responses {rel_err(y1.cpu().numpy(), y0.cpu().numpy()):.2e} loss {abs(l1 - l0) / abs(l0):.2e} "
          f"worst grad {worst[1]:.2e} ({worst[0]})")
    assert worst[1] < TOL_GRAD, worst


@needs_ref
def test_reference_train_step_runs_on_the_installed_modules():
    """The reference's own train_step (autocast off, GradScaler disabled, regulariser in the graph, AdamW) on the
    drop-in model: same loss / reg_loss / total_loss and the same accumulated gradients as on the reference model."""
    train = rh.import_train()
    n = 512
    over = dict(patch_stride=4)  # T = 121: the call sequence is what is under test here
    ref, ref_crit, new, new_crit = _pair({"A": n, "B": n}, **over)
    res, grads = [], []
    for model, crit in ((ref, ref_crit), (new, new_crit)):
        model.train(False)  # deterministic (no position noise); train_step itself does not switch modes
        opt = torch.optim.AdamW(model.get_parameters(core_lr=1e-3), lr=1e-3)
        scaler = torch.amp.GradScaler("cuda", enabled=False)
        out = {}
        for mouse in ("A", "B"):
            b = _batch(n, 8, seed=3 if mouse == "A" else 4)
            out[mouse] = train.train_step(mouse_id=mouse, batch=b, model=model, optimizer=opt, criterion=crit,
                                          scaler=scaler, update=False, micro_batch_size=4, device=torch.device(DEV))
        res.append(out)
        grads.append({k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None})
    for mouse in ("A", "B"):
        for key in ("loss/loss", "loss/reg_loss", "loss/total_loss"):
            a, c = float(res[0][mouse][key]), float(res[1][mouse][key])
            assert abs(a - c) / abs(a) < TOL_FWD, (mouse, key, a, c)
    for k, g in grads[0].items():
        assert rel_err(grads[1][k].cpu().numpy(), g.cpu().numpy()) < TOL_GRAD, k


@needs_ref
def test_reference_recorder_and_rollout_on_the_installed_modules():
    """attention_rollout.Recorder hooks ``mha.attend`` of the drop-in core and receives softmax(QK^T/sqrt(E)); the
    rebound attention_rollouts gives the reference's heat maps."""
    _, _, ar = rh.import_reference()
    n = 64
    ref, _, new, _ = _pair({"A": n})
    ref.train(False)
    new.train(False)
    b = _batch(n, 3)
    with torch.no_grad():
        rec0 = ar.Recorder(ref.core)
        _, attn0 = rec0(images=b["image"], behaviors=b["behavior"], pupil_centers=b["pupil_center"], mouse_id="A")
        rec0.eject()
        heat0 = ar.attention_rollouts(attn0, image_shape=(36, 64))  # the reference's per-sample matmul chain
    dropin.install()
    try:
        with torch.no_grad():
            rec1 = ar.Recorder(new.core)  # the reference's Recorder class, looking for the rebound Attention
            _, attn1 = rec1(images=b["image"], behaviors=b["behavior"], pupil_centers=b["pupil_center"], mouse_id="A")
            rec1.eject()
            heat1 = ar.attention_rollouts(attn1, image_shape=(36, 64))
    finally:
        dropin.uninstall()
    assert attn1 is not None and tuple(attn1.shape) == tuple(attn0.shape) == (3, 4, 4, 1654, 1654)
    assert rel_err(attn1.cpu().numpy(), attn0.cpu().numpy()) < TOL_FWD
    assert rel_err(np.asarray(heat1.cpu()), np.asarray(heat0.cpu())) < TOL_FWD


@needs_ref
def test_recorder_on_a_frozen_core_without_grad():
    """hooks on a frozen core (no input requires grad) must still get the probabilities."""
    from v1t_b200.rollout import Recorder

    ref, _, new, _ = _pair({"A": 32}, patch_stride=4)
    new.core.freeze()
    new.train(False)
    b = _batch(32, 2)
    rec = Recorder(new.core)
    for ctx in (torch.no_grad(), torch.enable_grad()):
        with ctx:
            _, attn = rec(images=b["image"], behaviors=b["behavior"], pupil_centers=b["pupil_center"], mouse_id="A")
        assert attn is not None and tuple(attn.shape) == (2, 4, 4, 121, 121)
        assert abs(float(attn.sum(-1).mean()) - 1.0) < 1e-4
    rec.eject()
