"""Fused L1-regulariser + AdamW: the optimizer half of the reference's train step.

Reference behaviour replaced (file:line under /root/reference):
  * ``reg_loss = (micro / batch) * model.regularizer(mouse_id)`` added to every micro-batch loss (train.py:71-73):
    ``reg_scale * sum|p|`` over the core parameters (vit.py:419-421), the mouse's readout features
    (gaussian2d.py:83-100) and the mouse's shifter (core_shifter.py:35-36), summed in model.py:141-149.  Its autograd
    contribution to ``p.grad`` is ``reg_scale * sign(p)`` per application; the fractions of one batch sum to 1.
  * ``torch.optim.AdamW(model.get_parameters(core_lr), lr, betas, eps, weight_decay=0)`` (train.py:217-223) stepped
    once per sweep over the mice (train.py:77-80, 97-111).

Here both are one kernel launch over all parameter tensors (``v1t_adamw_l1_step``): the regulariser never enters the
autograd graph, its value (``sum|p|`` per group) is a by-product of the same pass.  ``FusedAdamWL1`` keeps
``torch.optim.AdamW``'s param-group keys and per-parameter state (``step``, ``exp_avg``, ``exp_avg_sq``), so the
``optimizer`` entry of a reference checkpoint (scheduler.py:88-100) loads into it and vice versa.
"""
from __future__ import annotations

import ctypes as C
import math
import typing as t

import numpy as np
import torch

from . import _lib
from .functional import _stream_ptr


def l1_coefficients(model, mouse_ids: t.Sequence[str], per_mouse_groups: bool = False
                    ) -> t.Dict[torch.nn.Parameter, t.Tuple[float, int]]:
    """(coefficient, group) of the L1 term each parameter receives over ONE optimizer step that visits every mouse
    in ``mouse_ids`` once (train.py:97-111).  Groups: 0 = core (regularised once per mouse step, model.py:143-144),
    1 = readout features, 2 = shifters (parameters without a regulariser land in the optimizer's last group).
    ``per_mouse_groups``: mouse i gets its own groups 1+3i (readout features), 2+3i (core shifter), 3+3i (image
    shifter), so that the per-mouse ``loss/reg_loss`` of train.py:71 can be read back (see ``l1_group_count``)."""
    coef: t.Dict[torch.nn.Parameter, t.Tuple[float, int]] = {}
    if not model.core.frozen:
        scale = float(model.core.reg_scale) * len(mouse_ids)
        for p in model.core.parameters():
            coef[p] = (scale, 0)
    for i, m in enumerate(mouse_ids):
        g_feat, g_cs, g_is = (1 + 3 * i, 2 + 3 * i, 3 + 3 * i) if per_mouse_groups else (1, 2, 2)
        readout = model.readouts[m]
        coef[readout.features] = (float(readout.reg_scale), g_feat)
        if getattr(model, "core_shifter", None) is not None:
            shifter = model.core_shifter[m]
            for p in shifter.parameters():
                coef[p] = (float(shifter.reg_scale), g_cs)
        image_shifter = getattr(model.image_cropper, "image_shifter", None)
        if image_shifter is not None:
            for p in image_shifter[m].parameters():
                coef[p] = (float(image_shifter[m].reg_scale), g_is)
    return coef


def l1_group_count(n_mice: int) -> int:
    """Number of L1 groups of the per-mouse layout: core + 3 per mouse + the trailing unregularised group."""
    return 1 + 3 * n_mice + 1


L1_GROUP_NAMES = ("core", "readout_features", "shifters", "unregularised")  # the last one: every other parameter


class FusedAdamWL1(torch.optim.Optimizer):
    """AdamW with the L1 regulariser's gradient folded in; CUDA-only (no CPU path).

    ``l1``: {parameter: coefficient} or {parameter: (coefficient, group)} (see ``l1_coefficients``).  A regularised
    parameter without a gradient is stepped with a zero data gradient — in the reference such a parameter always has
    one, from the regulariser."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 l1: t.Optional[dict] = None, n_l1_groups: int = len(L1_GROUP_NAMES)):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("FusedAdamWL1: invalid hyper-parameter")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None,
                        decoupled_weight_decay=True)
        super().__init__(params, defaults)
        self.n_l1_groups = n_l1_groups
        self.set_l1(l1 or {})
        self._table_key = None
        self._table = None
        self.last_l1_sums: t.Optional[torch.Tensor] = None

    def set_l1(self, l1: dict):
        self._l1 = {}
        for p, v in l1.items():
            c, g = (v if isinstance(v, tuple) else (v, 0))  # a bare coefficient goes to group 0
            if not 0 <= g < self.n_l1_groups:
                raise ValueError(f"FusedAdamWL1: l1 group {g} outside 0..{self.n_l1_groups - 1}")
            self._l1[p] = (float(c), int(g))
        self._table_key = None

    # ---- device table -------------------------------------------------------------------------------
    def _entries(self):
        out = []
        for group in self.param_groups:
            if group["amsgrad"] or group["maximize"]:
                raise NotImplementedError("FusedAdamWL1: amsgrad / maximize are not implemented")
            for p in group["params"]:
                c, g = self._l1.get(p, (0.0, self.n_l1_groups - 1))
                if p.grad is None:
                    if c == 0.0:
                        continue
                    p.grad = torch.zeros_like(p)
                if not p.is_cuda:
                    raise RuntimeError("FusedAdamWL1 is CUDA-only (no CPU path): parameter on " + str(p.device))
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdamWL1: parameters and gradients must be contiguous fp32")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                out.append((p, group, st, c, g))
        return out

    def _build_table(self, entries, device):
        lib = _lib.load()
        chunk = lib.v1t_opt_chunk_elems()
        key = tuple((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                     p.numel(), float(group["lr"]), c, float(group["weight_decay"]), g)
                    for p, group, st, c, g in entries)
        if key == self._table_key:
            return self._table
        rec = (_lib.OptTensor * len(entries))()
        prefix = np.zeros(len(entries) + 1, dtype=np.int32)
        for i, (pp, gp, mp, vp, n, lr, c, wd, g) in enumerate(key):
            rec[i] = _lib.OptTensor(param=pp, grad=gp, exp_avg=mp, exp_avg_sq=vp, numel=n, lr=lr, l1=c,
                                    weight_decay=wd, group=g)
            prefix[i + 1] = prefix[i] + (n + chunk - 1) // chunk
        raw = np.frombuffer(rec, dtype=np.uint8).copy()
        table = torch.from_numpy(raw).to(device)
        prefix_dev = torch.from_numpy(prefix).to(device)
        n_chunks = int(prefix[-1])
        scratch = torch.empty(max(1, lib.v1t_adamw_l1_scratch_bytes(n_chunks)), dtype=torch.uint8, device=device)
        sums = torch.zeros(self.n_l1_groups, dtype=torch.float32, device=device)
        self._table_key = key
        self._table = (table, prefix_dev, len(entries), n_chunks, scratch, sums)
        return self._table

    # ---- step ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0, zero_grad: bool = False):
        """One AdamW step.  ``grad_scale`` multiplies the data gradients (e.g. 1/loss_scale); ``zero_grad`` clears
        them in the same pass (the reference calls optimizer.zero_grad() right after, train.py:80).  After the call
        ``last_l1_sums`` holds sum|p| per L1 group as it was BEFORE the update."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries = self._entries()
        if not entries:
            return loss
        by_hyper: t.Dict[tuple, list] = {}
        for e in entries:
            group, st = e[1], e[2]
            hk = (float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), float(st["step"]),
                  e[0].device)
            by_hyper.setdefault(hk, []).append(e)
        if len(by_hyper) != 1:
            raise NotImplementedError("FusedAdamWL1: all parameters must share betas, eps, step count and device")
        (beta1, beta2, eps, step, device), _ = next(iter(by_hyper.items()))
        step += 1.0
        lib = _lib.load()
        table, prefix, n_tensors, n_chunks, scratch, sums = self._build_table(entries, device)
        bc1 = 1.0 - beta1 ** step
        bc2_sqrt = math.sqrt(1.0 - beta2 ** step)
        with torch.cuda.device(device):
            _lib.check(lib.v1t_adamw_l1_step(table.data_ptr(), prefix.data_ptr(), n_tensors, n_chunks, beta1, beta2,
                                             eps, bc1, bc2_sqrt, float(grad_scale), int(zero_grad), sums.data_ptr(),
                                             self.n_l1_groups, scratch.data_ptr(), _stream_ptr(device)),
                       "adamw_l1_step")
        for e in entries:
            e[2]["step"] += 1.0
        self.last_l1_sums = sums
        return loss

    def reg_loss(self, reg_scales: t.Sequence[float]) -> torch.Tensor:
        """What the reference reports as loss/reg_loss for the parameters as they were before the last step: the sum
        over groups of reg_scale[group] * sum|p|."""
        if self.last_l1_sums is None:
            raise RuntimeError("FusedAdamWL1.reg_loss: no step taken yet")
        scales = list(reg_scales)[: self.n_l1_groups]
        w = torch.tensor(scales + [0.0] * (self.n_l1_groups - len(scales)), dtype=torch.float32,
                         device=self.last_l1_sums.device)
        return (self.last_l1_sums * w).sum()


def build_optimizer(model, args, mouse_ids: t.Sequence[str]) -> FusedAdamWL1:
    """The reference's optimizer construction (train.py:216-223) with the regulariser folded in; per-mouse L1 groups
    so that ``train_step`` can report the reference's per-mouse ``loss/reg_loss``."""
    core_lr = args.lr if getattr(args, "core_lr", None) is None else args.core_lr
    mouse_ids = list(mouse_ids)
    opt = FusedAdamWL1(model.get_parameters(core_lr=core_lr), lr=args.lr,
                       betas=(args.adam_beta1, args.adam_beta2), eps=args.adam_eps, weight_decay=0,
                       l1=l1_coefficients(model, mouse_ids, per_mouse_groups=True),
                       n_l1_groups=l1_group_count(len(mouse_ids)))
    image_shifter = getattr(model.image_cropper, "image_shifter", None)
    opt.l1_layout = {
        "core": 0.0 if model.core.frozen else float(model.core.reg_scale),
        "mice": {m: (i, float(model.readouts[m].reg_scale),
                     float(model.core_shifter[m].reg_scale) if getattr(model, "core_shifter", None) is not None else 0.0,
                     float(image_shifter[m].reg_scale) if image_shifter is not None else 0.0)
                 for i, m in enumerate(mouse_ids)}}
    return opt


def reg_loss_of(optimizer: FusedAdamWL1, model, mouse_id: str) -> torch.Tensor:
    """model.regularizer(mouse_id) (model.py:141-149) from the |p| sums of the optimizer's last pass (the parameters
    as they were before that update); computed directly while no step has been taken yet."""
    layout = getattr(optimizer, "l1_layout", None)
    if layout is None or optimizer.last_l1_sums is None or mouse_id not in layout["mice"]:
        with torch.no_grad():
            return torch.as_tensor(model.regularizer(mouse_id), dtype=torch.float32)
    i, r_feat, r_cs, r_is = layout["mice"][mouse_id]
    s = optimizer.last_l1_sums
    return layout["core"] * s[0] + r_feat * s[1 + 3 * i] + r_cs * s[2 + 3 * i] + r_is * s[3 + 3 * i]


def train_step(mouse_id: str, batch: t.Dict[str, torch.Tensor], model, optimizer: FusedAdamWL1, criterion,
               scaler=None, update: bool = True, micro_batch_size: int = 0, device: torch.device = "cuda"
               ) -> t.Dict[str, torch.Tensor]:
    """Same signature and result keys as the reference's train_step (train.py:42-81), so the call at train.py:97-108
    works unchanged.  ``scaler`` is accepted and ignored (fp16 AMP is replaced by the declared bf16x3 / bf16 modes,
    SURVEY F7).  The regulariser is not part of the graph: its gradient is applied by ``optimizer.step``; its value
    (``loss/reg_loss``, per mouse) comes from the optimizer's by-product, i.e. one update behind the parameters."""
    model.to(device)
    micro_batch_size = micro_batch_size or batch["image"].size(0)
    batch_size = batch["image"].size(0)
    losses = []
    for lo in range(0, batch_size, micro_batch_size):
        sl = slice(lo, lo + micro_batch_size)
        y_true = batch["response"][sl].to(device)
        y_pred, _, _ = model(inputs=batch["image"][sl].to(device), mouse_id=mouse_id,
                             behaviors=batch["behavior"][sl].to(device),
                             pupil_centers=batch["pupil_center"][sl].to(device))
        loss = criterion(y_true=y_true, y_pred=y_pred, mouse_id=mouse_id, batch_size=batch_size)
        loss.backward()
        losses.append(loss.detach())
    loss = torch.stack(losses).sum()
    reg = reg_loss_of(optimizer, model, mouse_id).to(loss.device)  # the micro-batch fractions of train.py:71 sum to 1
    result = {"loss/loss": loss.cpu(), "loss/reg_loss": reg.cpu(), "loss/total_loss": (loss + reg).cpu()}
    if update:
        optimizer.step(zero_grad=True)
    return result
