"""v1t_b200 — B200-native (sm_100a) implementation of the V1T hot path: ViT core + Gaussian2d readout +
Poisson loss, forward and backward, behind the reference's nn.Module / registry API."""
from . import _lib  # noqa: F401
from .modules import (  # noqa: F401
    Attention, BehaviorMLP, Core, CoreShifter, CoreShifters, ELU1, Gaussian2DReadout, Image2Patches, ImageCropper,
    Loss, MLP, Model, PoissonLoss, Readout, Readouts, Transformer, ViTCore, get_core, get_criterion,
)
from . import functional  # noqa: F401
from . import ensemble, optim, rollout  # noqa: F401
from .ensemble import EnsembleModel, OutputModule  # noqa: F401
from .optim import FusedAdamWL1, build_optimizer  # noqa: F401
from .rollout import Recorder, attention_rollouts  # noqa: F401

__all__ = ["ViTCore", "Gaussian2DReadout", "PoissonLoss", "ELU1", "Model", "Readouts", "get_core", "get_criterion",
           "functional", "FusedAdamWL1", "build_optimizer", "Recorder", "attention_rollouts", "EnsembleModel",
           "OutputModule", "ImageCropper", "CoreShifters"]
