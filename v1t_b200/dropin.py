"""Plug the B200 modules into an UNMODIFIED reference checkout.

``train.py`` hard-codes the accepted ``--core`` / ``--readout`` names (train.py:524-650), so a new registry
name cannot be selected from its CLI; instead the existing keys are overwritten with plain dict writes:

    import v1t_b200.dropin as dropin
    dropin.install()            # after `import v1t` is possible (reference's src/ on sys.path)
    import train; train.main(args)

``train.py`` builds ``torch.optim.AdamW`` and adds ``model.regularizer`` to every loss itself (train.py:71,217-223);
the fused optimizer is opt-in: ``v1t_b200.optim.build_optimizer(model, args, mouse_ids)`` + ``optim.train_step``
(INTEGRATION.md shows the two-line patch).

After ``install()`` the reference's ``Model(args, ds)``, ``losses.get_criterion`` (train.py:215-224),
``ensemble.py`` and ``demo.ipynb`` construct B200 modules; checkpoints stay interchangeable because the
state-dict layout is identical.
"""
from __future__ import annotations

import sys

_SAVED = []  # (object, attribute-or-key, previous value) of everything install() rebinds


def _rebind(obj, name, value, is_key=False):
    if is_key:
        _SAVED.append((obj, name, obj.get(name), True))
        obj[name] = value
    else:
        _SAVED.append((obj, name, getattr(obj, name, None), False))
        setattr(obj, name, value)


def uninstall():
    """Undo install(): the reference's own classes are back in its registries (tests compare the two side by side)."""
    while _SAVED:
        obj, name, old, is_key = _SAVED.pop()
        if is_key:
            obj[name] = old
        else:
            setattr(obj, name, old)


def install(stub_torchinfo: bool = True):
    from . import modules as M

    if _SAVED:
        return M

    import v1t.models.core.core as ref_core  # type: ignore
    import v1t.models.readout.readout as ref_readout  # type: ignore
    import v1t.losses as ref_losses  # type: ignore
    import v1t.models.core.vit as ref_vit  # type: ignore
    import v1t.models.model as ref_model  # type: ignore

    _rebind(ref_core._CORES, "vit", M.ViTCore, is_key=True)  # core/core.py:13
    _rebind(ref_readout._READOUTS, "gaussian2d", M.Gaussian2DReadout, is_key=True)  # readout/readout.py:15
    _rebind(ref_losses._CRITERION, "poisson", M.PoissonLoss, is_key=True)  # losses.py:15
    _rebind(ref_model, "ELU1", M.ELU1)  # Model.__init__ instantiates ELU1() (model.py:105)
    # the callers either side of the path are constructed by name from model.py's own imports (model.py:14-15,66,85)
    _rebind(ref_model, "ImageCropper", M.ImageCropper)
    _rebind(ref_model, "CoreShifters", M.CoreShifters)
    # attention_rollout.Recorder finds blocks with isinstance(m, Attention) (attention_rollout.py:26-33)
    _rebind(ref_vit, "Attention", M.Attention)
    _rebind(ref_vit, "ViTCore", M.ViTCore)
    ar = sys.modules.get("v1t.utils.attention_rollout")
    if ar is not None:
        from . import rollout as R

        _rebind(ar, "Attention", M.Attention)
        _rebind(ar, "ViTCore", M.ViTCore)
        # batched GPU rollout behind the reference's function names (attention_rollout.py:92-133)
        _rebind(ar, "attention_rollout", R.attention_rollout)
        _rebind(ar, "attention_rollouts", R.attention_rollouts)
    if stub_torchinfo:
        # get_model() runs torchinfo forward passes on CPU tensors before model.to(device)
        # (model.py:187-226); the B200 modules are CUDA-only, so summaries are skipped.
        _rebind(ref_model, "get_model_info", lambda *a, **k: None)
    return M
