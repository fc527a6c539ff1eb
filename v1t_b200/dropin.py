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


def install(stub_torchinfo: bool = True):
    from . import modules as M

    import v1t.models.core.core as ref_core  # type: ignore
    import v1t.models.readout.readout as ref_readout  # type: ignore
    import v1t.losses as ref_losses  # type: ignore
    import v1t.models.core.vit as ref_vit  # type: ignore
    import v1t.models.model as ref_model  # type: ignore

    ref_core._CORES["vit"] = M.ViTCore  # core/core.py:13
    ref_readout._READOUTS["gaussian2d"] = M.Gaussian2DReadout  # readout/readout.py:15
    ref_losses._CRITERION["poisson"] = M.PoissonLoss  # losses.py:15
    ref_model.ELU1 = M.ELU1  # Model.__init__ instantiates ELU1() (model.py:105)
    # the callers either side of the path are constructed by name from model.py's own imports (model.py:14-15,66,85)
    ref_model.ImageCropper = M.ImageCropper
    ref_model.CoreShifters = M.CoreShifters
    # attention_rollout.Recorder finds blocks with isinstance(m, Attention) (attention_rollout.py:26-33)
    ref_vit.Attention = M.Attention
    ref_vit.ViTCore = M.ViTCore
    ar = sys.modules.get("v1t.utils.attention_rollout")
    if ar is not None:
        from . import rollout as R

        ar.Attention, ar.ViTCore = M.Attention, M.ViTCore
        # batched GPU rollout behind the reference's function names (attention_rollout.py:92-133)
        ar.attention_rollout, ar.attention_rollouts = R.attention_rollout, R.attention_rollouts
    if stub_torchinfo:
        # get_model() runs torchinfo forward passes on CPU tensors before model.to(device)
        # (model.py:187-226); the B200 modules are CUDA-only, so summaries are skipped.
        ref_model.get_model_info = lambda *a, **k: None
    return M
