"""Attention-rollout extraction on the GPU — mirror of src/v1t/utils/attention_rollout.py.

``Recorder`` keeps the reference's hook contract (forward hooks on every block's ``mha.attend`` receive the softmax
probabilities, attention_rollout.py:31-36; the core materialises them on demand through ``v1t_attention_probs``);
``attention_rollouts`` replaces the per-sample Python loop of chained T x T matmuls (attention_rollout.py:92-133) by
one fused vector-matrix pass per block (csrc/rollout.cu)."""
from __future__ import annotations

import math
import typing as t

import torch
from torch import nn

from . import functional as VF
from .modules import Attention, ViTCore


def find_shape(num_patches: int):
    """attention_rollout.py:78-83."""
    dim1 = math.ceil(math.sqrt(num_patches))
    while num_patches % dim1 != 0 and dim1 > 0:
        dim1 -= 1
    return dim1, num_patches // dim1


class Recorder(nn.Module):
    """attention_rollout.py:15-75."""

    def __init__(self, core: ViTCore):
        super().__init__()
        self.core = core
        self.cache: t.List[torch.Tensor] = []
        self.hooks = []
        self.hook_registered = False
        self.ejected = False

    def _hook(self, _, inputs, outputs):
        self.cache.append(outputs.detach())  # the probabilities are a fresh tensor per call: no clone needed

    def _register_hook(self):
        for module in self.core.transformer.modules():
            if isinstance(module, Attention):
                self.hooks.append(module.attend.register_forward_hook(self._hook))
        self.hook_registered = True

    def eject(self):
        self.ejected = True
        for hook in self.hooks:
            hook.remove()
        self.hooks.clear()
        return self.core

    def clear(self):
        self.cache.clear()

    def forward(self, images, behaviors, pupil_centers, mouse_id: str):
        """Returns (core output, attentions [B, blocks, heads, T, T])."""
        assert not self.ejected, "recorder has been ejected, cannot be used anymore"
        self.clear()
        if not self.hook_registered:
            self._register_hook()
        outputs = self.core(inputs=images, behaviors=behaviors, pupil_centers=pupil_centers, mouse_id=mouse_id)
        attentions = torch.stack(self.cache, dim=1) if self.cache else None
        return outputs, attentions


def attention_rollouts(attentions: torch.Tensor, image_shape: t.Sequence[int]) -> torch.Tensor:
    """attentions [B,L,H,T,T] -> heat maps [B,*image_shape] (attention_rollout.py:124-133)."""
    assert attentions.dim() == 5
    return VF.attention_rollouts(attentions, image_shape, find_shape(attentions.shape[-1] - 1))


def attention_rollout(attention: torch.Tensor, image_shape: t.Sequence[int]) -> torch.Tensor:
    """One sample [L,H,T,T] (attention_rollout.py:92-121)."""
    assert attention.dim() == 4
    return attention_rollouts(attention[None], image_shape)[0]


@torch.no_grad()
def extract_attention_maps(ds, model, num_samples: int = None, device="cuda") -> t.Dict[str, "torch.Tensor"]:
    """attention_rollout.py:136-203 without the dataset's inverse transforms being optional: they are applied when
    the dataset provides them."""
    model.to(device)
    model.train(False)
    dataset = ds.dataset
    mouse_id = dataset.mouse_id
    ident = lambda x: x  # noqa: E731
    inv_image = getattr(dataset, "i_transform_image", ident)
    inv_behavior = getattr(dataset, "i_transform_behavior", ident)
    inv_pupil = getattr(dataset, "i_transform_pupil_center", ident)
    recorder = Recorder(model.core)
    results = {"images": [], "heatmaps": [], "pupil_centers": [], "behaviors": []}
    count = num_samples
    for batch in ds:
        images, behaviors = batch["image"].to(device), batch["behavior"].to(device)
        pupil_centers = batch["pupil_center"].to(device)
        images, _ = model.image_cropper(inputs=images, mouse_id=mouse_id, behaviors=behaviors,
                                        pupil_centers=pupil_centers)
        _, attentions = recorder(images=images, behaviors=behaviors, pupil_centers=pupil_centers, mouse_id=mouse_id)
        recorder.clear()
        heatmaps = attention_rollouts(attentions, image_shape=images.shape[2:])
        results["images"].append(inv_image(images.cpu()))
        results["heatmaps"].append(heatmaps.cpu())
        results["behaviors"].append(inv_behavior(behaviors.cpu()))
        results["pupil_centers"].append(inv_pupil(pupil_centers.cpu()))
        if num_samples is not None and (count := count - len(images)) <= 0:
            break
    recorder.eject()
    results = {k: torch.vstack(v).numpy() for k, v in results.items()}
    if num_samples is not None:
        results = {k: v[:num_samples] for k, v in results.items()}
    return results
