"""Import the LIVE reference (bryanlimy/V1T) — TEST / BASELINE INFRASTRUCTURE.

Looks for the reference at $V1T_REFERENCE_SRC, then /root/reference/src (build container), then oracle/_ref/src
(the byte-identical copy staged by oracle/make_ref.py, which is what exists on the GPU box).  Used by
scripts/make_golden*.py, tests/, __graft_entry__.smoke() and bench.py's reference / eager-baseline legs only.
Recipe from SURVEY.md Appendix C: stub ``torchinfo`` (model.py:4) and ``v1t.utils.tensorboard`` (model.py:13, needs
matplotlib/seaborn) so ``from v1t.models import Model`` works without touching the reference tree; ``import_train``
additionally stubs ruamel.yaml / h5py (utils/yaml.py:7, data.py) so the reference's own train.py imports.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_ref_src() -> str:
    cands = [os.environ.get("V1T_REFERENCE_SRC"), "/root/reference/src", os.path.join(_HERE, "_ref", "src")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "v1t")):
            return c
    return cands[1]


REF_SRC = _find_ref_src()
REF_ROOT = os.path.dirname(REF_SRC)  # train.py / ensemble.py live next to src/


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "v1t"))


def import_reference():
    """Returns (Model, losses, attention_rollout module)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_SRC}")
    if "torchinfo" not in sys.modules:
        try:
            import torchinfo  # noqa: F401
        except Exception:
            sys.modules["torchinfo"] = types.ModuleType("torchinfo")
    if "v1t.utils.tensorboard" not in sys.modules:
        tb = types.ModuleType("v1t.utils.tensorboard")
        tb.Summary = object
        sys.modules["v1t.utils.tensorboard"] = tb
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from v1t.models import Model  # type: ignore
    from v1t import losses  # type: ignore
    from v1t.utils import attention_rollout  # type: ignore

    return Model, losses, attention_rollout


class _Stub:
    """Attribute/call sink for optional third-party modules the hot path never touches."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Stub()

    def __call__(self, *a, **k):
        return _Stub()


def import_train():
    """The reference's own top-level ``train`` module (train_step: train.py:42-81) with plotting / yaml / hdf5
    dependencies stubbed.  Returns the module."""
    import_reference()
    for name in ("ruamel", "ruamel.yaml", "matplotlib", "seaborn", "h5py"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["ruamel.yaml"], "YAML"):
        sys.modules["ruamel.yaml"].YAML = _Stub
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import train  # type: ignore

    return train


class FakeDataset:
    """Only what Model / PoissonLoss touch: readout.py:36, gaussian2d.py:186, losses.py:107-112."""

    def __init__(self, num_neurons: int, ds_size: int, rng: np.random.Generator):
        self.coordinates = rng.standard_normal((num_neurons, 3)).astype(np.float32)
        self.response_stats = {
            "mean": np.ones((num_neurons,), dtype=np.float32),
            "std": np.ones((num_neurons,), dtype=np.float32),
        }
        self._n = ds_size

    def __len__(self):
        return self._n


class FakeLoader:
    def __init__(self, dataset):
        self.dataset = dataset


def make_args(neurons: dict, in_shape=(1, 36, 64), **over):
    """Default train.py args of the hot path (train.py:543-590,636-657), overridable."""
    a = dict(
        input_shape=tuple(in_shape), output_shapes={k: (n,) for k, n in neurons.items()},
        ds_name="sensorium", device=torch.device("cpu"), core="vit", readout="gaussian2d",
        behavior_mode=3, shift_mode=2, center_crop=1.0, resize_image=0, verbose=0,
        patch_mode=0, patch_size=8, patch_stride=1, emb_dim=155, num_blocks=4, num_heads=4,
        mlp_dim=488, p_dropout=0.0229, t_dropout=0.2544, drop_path=0.0, use_lsa=False,
        disable_bias=False, grad_checkpointing=0, core_reg_scale=0.5379,
        readout_reg_scale=0.0076, disable_grid_predictor=False, grid_predictor_dim=2,
        bias_mode=0, shifter_reg_scale=0.0, cropper_reg_scale=0.0, criterion="poisson",
        ds_scale=1, gray_scale=False,
    )
    a.update(over)
    return SimpleNamespace(**a)


def make_fake_ds(neurons: dict, ds_size=4500, seed=1234):
    rng = np.random.default_rng(seed)
    return {k: FakeLoader(FakeDataset(n, ds_size, rng)) for k, n in neurons.items()}


def build_reference_model(args, ds, seed=1234, trained_like=True):
    """Reference Model + PoissonLoss under a seed; optionally perturb readout so outputs vary."""
    Model, losses, _ = import_reference()
    torch.manual_seed(seed)
    model = Model(args, ds=ds)
    crit = losses.get_criterion(args, ds=ds)
    if trained_like:  # SURVEY.md §8d "trained-like" weights
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for r in model.readouts.values():
                r.features.copy_(torch.randn(r.features.shape, generator=g) * 0.05 + 1.0 / r.features.shape[1])
                r.bias.copy_(torch.randn(r.bias.shape, generator=g) * 0.3)
                r.sigma.copy_((torch.rand(r.sigma.shape, generator=g) - 0.5) * 0.6)
                for m in r.mu_transform if hasattr(r, "mu_transform") else []:
                    if hasattr(m, "weight"):
                        m.weight.mul_(3.0)
            if getattr(model, "core_shifter", None) is not None:
                for s in model.core_shifter.values():
                    for p in s.parameters():
                        p.add_(torch.randn(p.shape, generator=g) * 0.3)
            for blk in model.core.transformer.blocks:
                for lin in (blk["mha"].projection[0], blk["mlp"].model[1], blk["mlp"].model[4]):
                    if lin.bias is not None:
                        lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.05)
                for ln in (blk["mha"].layer_norm, blk["mlp"].model[0]):
                    ln.weight.add_(torch.randn(ln.weight.shape, generator=g) * 0.1)
                    ln.bias.add_(torch.randn(ln.bias.shape, generator=g) * 0.1)
                if "b-mlp" in blk:
                    for mlp in blk["b-mlp"].models.values():
                        for m in mlp:
                            if hasattr(m, "weight"):
                                m.weight.mul_(10.0)
                                if m.bias is not None:
                                    m.bias.add_(torch.randn(m.bias.shape, generator=g) * 0.1)
    return model, crit
