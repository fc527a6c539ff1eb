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
    """Collects the attention probabilities of every block during a core forward (the contract of
    attention_rollout.py:15-75: hooks on ``mha.attend``; ``forward`` returns (core output, [B, blocks, heads, T, T]);
    ``eject`` gives the core back without hooks)."""

    def __init__(self, core: ViTCore):
        super().__init__()
        self.core = core
        self.cache: t.List[torch.Tensor] = []
        self._handles: t.List = []
        self.ejected = False

    @property
    def hook_registered(self) -> bool:
        return bool(self._handles)

    def _attach(self):
        attends = [m.attend for m in self.core.transformer.modules() if isinstance(m, Attention)]
        # the probabilities handed to the hook are a fresh tensor per call: keep them without cloning
        self._handles = [a.register_forward_hook(lambda _m, _i, probs: self.cache.append(probs.detach()))
                         for a in attends]

    def clear(self):
        self.cache = []

    def eject(self) -> ViTCore:
        while self._handles:
            self._handles.pop().remove()
        self.ejected = True
        return self.core

    def forward(self, images, behaviors, pupil_centers, mouse_id: str):
        if self.ejected:
            raise AssertionError("recorder has been ejected, cannot be used anymore")
        self.clear()
        if not self._handles:
            self._attach()
        outputs = self.core(inputs=images, behaviors=behaviors, pupil_centers=pupil_centers, mouse_id=mouse_id)
        return outputs, (torch.stack(self.cache, dim=1) if self.cache else None)


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
    """Rollout heat maps for the samples of one mouse's DataLoader (attention_rollout.py:136-203): returns numpy arrays
    ``images``, ``heatmaps``, ``behaviors``, ``pupil_centers`` (inverse-transformed when the dataset provides the
    ``i_transform_*`` callables), truncated to ``num_samples``."""
    dataset = ds.dataset
    mouse_id = dataset.mouse_id
    undo = {name: getattr(dataset, attr, None) or (lambda x: x)
            for name, attr in (("images", "i_transform_image"), ("behaviors", "i_transform_behavior"),
                               ("pupil_centers", "i_transform_pupil_center"))}
    model.to(device).train(False)
    recorder = Recorder(model.core)
    chunks: t.Dict[str, list] = {"images": [], "heatmaps": [], "behaviors": [], "pupil_centers": []}
    seen = 0
    try:
        for batch in ds:
            if num_samples is not None and seen >= num_samples:
                break
            x = {k: batch[k].to(device) for k in ("image", "behavior", "pupil_center")}
            images, _ = model.image_cropper(inputs=x["image"], mouse_id=mouse_id, behaviors=x["behavior"],
                                            pupil_centers=x["pupil_center"])
            _, attentions = recorder(images=images, behaviors=x["behavior"], pupil_centers=x["pupil_center"],
                                     mouse_id=mouse_id)
            recorder.clear()  # one batch of [B, L, H, T, T] at a time
            chunks["heatmaps"].append(attention_rollouts(attentions, image_shape=images.shape[2:]).cpu())
            chunks["images"].append(undo["images"](images.cpu()))
            chunks["behaviors"].append(undo["behaviors"](x["behavior"].cpu()))
            chunks["pupil_centers"].append(undo["pupil_centers"](x["pupil_center"].cpu()))
            seen += len(images)
    finally:
        recorder.eject()
    return {k: torch.vstack(v).numpy()[:num_samples] for k, v in chunks.items()}
