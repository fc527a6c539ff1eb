"""Ensemble inference — mirror of ensemble.py:30-151 (OutputModule, EnsembleModel).

The members are ordinary ``v1t_b200.Model`` instances (frozen, ensemble.py:106) run back to back through the native
path on the same inputs with ``activate=False``; the output module (mean / shared Linear / per-mouse Linear, then
ELU+1) is one fused kernel over the K response tensors (csrc/ensemble.cu) instead of rearrange + cat + mean/Linear +
ELU1 over a materialised [B,N,K] stack.  State-dict keys match the reference (``ensemble.<name>.…``,
``output_module.linear.…``)."""
from __future__ import annotations

import typing as t

import torch
from torch import nn

from . import functional as VF
from .modules import ELU1, Model


class OutputModule(nn.Module):
    """ensemble.py:30-80.  ensemble_mode 0: mean; 1: one Linear(K,1); 2: one Linear(K,1) per mouse."""

    def __init__(self, args, in_features: int):
        super().__init__()
        self.in_features = in_features
        self.output_shapes = args.output_shapes
        self.ensemble_mode = args.ensemble_mode
        assert self.ensemble_mode in (0, 1, 2)
        if self.ensemble_mode == 1:
            self.linear = nn.Linear(in_features, 1)
        elif self.ensemble_mode == 2:
            self.linear = nn.ModuleDict({m: nn.Linear(in_features, 1) for m in self.output_shapes.keys()})
        self.activation = ELU1()
        for m in self.modules():
            if isinstance(m, nn.Linear):  # ensemble.py:57-66
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.constant_(m.bias, 0)

    def forward(self, members: t.Sequence[torch.Tensor], mouse_id: str):
        """members: the K pre-activation responses [B,N] (the reference passes them concatenated as [B,N,K])."""
        if self.ensemble_mode == 0:
            return VF.ensemble_combine(members)
        lin = self.linear if self.ensemble_mode == 1 else self.linear[mouse_id]
        return VF.ensemble_combine(members, lin.weight, lin.bias)


class EnsembleModel(nn.Module):
    """ensemble.py:83-151 with the members passed in (checkpoint discovery / args.yaml parsing stay with the
    caller: ``members`` maps the reference's model names to constructed-and-loaded Models)."""

    def __init__(self, args, members: t.Dict[str, Model], concurrent: bool = True):
        super().__init__()
        self.input_shape = args.input_shape
        self.output_shapes = args.output_shapes
        self.ensemble = nn.ModuleDict(members)
        self.ensemble.requires_grad_(False)
        self.output_module = OutputModule(args, in_features=len(members))
        # The K members share inputs but not weights: each runs its native launch sequence on its OWN CUDA stream (own
        # scratch arena), so their kernels interleave on the device instead of running back to back; the output module
        # joins the streams.  (One grouped launch sequence with a model index per row block is the design that would
        # remove the remaining per-member launches; see DESIGN.md.)
        self.concurrent = concurrent
        self._streams: t.List[torch.cuda.Stream] = []

    def regularizer(self, mouse_id: str):
        return torch.tensor(0.0)

    def forward(self, inputs, mouse_id: str, behaviors, pupil_centers):
        names = list(self.ensemble.keys())
        outs = []
        side = self.concurrent and inputs.is_cuda and len(names) > 1 and not torch.is_grad_enabled()
        if side:
            cur = torch.cuda.current_stream(inputs.device)
            while len(self._streams) < len(names):
                self._streams.append(torch.cuda.Stream(device=inputs.device))
            for name, st in zip(names, self._streams):
                st.wait_stream(cur)  # inputs (and the members' weights) are ready on the caller's stream
                with torch.cuda.stream(st):
                    y, _, _ = self.ensemble[name](inputs, mouse_id=mouse_id, behaviors=behaviors,
                                                  pupil_centers=pupil_centers, activate=False)
                    y.record_stream(cur)
                outs.append(y)
            for st in self._streams[: len(names)]:
                cur.wait_stream(st)
            for x in (inputs, behaviors, pupil_centers):
                for st in self._streams[: len(names)]:
                    x.record_stream(st)
        else:
            for name in names:
                y, _, _ = self.ensemble[name](inputs, mouse_id=mouse_id, behaviors=behaviors,
                                              pupil_centers=pupil_centers, activate=False)
                outs.append(y)
        return self.output_module(outs, mouse_id=mouse_id), None, None
