"""nn.Module boundary of the B200 hot path: same constructor arguments, attribute paths, registry mechanism
and state-dict layout as the reference, with forward/backward running in libv1t_b200 (sm_100a CUDA).

Mirrors (file:line under /root/reference/src/v1t/):
  Core, register, get_core          models/core/core.py:1-65
  Image2Patches / MLP / BehaviorMLP / Attention / Transformer / ViTCore
                                    models/core/vit.py:41-129,132-154,157-202,205-284,287-362,365-436
  Readout, Readouts, register       models/readout/readout.py:1-85
  Gaussian2DReadout                 models/readout/gaussian2d.py:13-278
  ELU1                              models/utils.py:109-118
  Loss / PoissonLoss / get_criterion   losses.py:96-119,141-166,193-197
  CoreShifter(s)                    models/core_shifter.py:7-69     (tiny tanh MLP: stays in torch)

The sub-modules (Attention, MLP, ...) are parameter containers with the reference's names so that
``state_dict()`` keys match Appendix B of SURVEY.md exactly; the arithmetic happens once, fused, in
``ViTCore.forward``.  Flags outside the hot-path scope raise NotImplementedError (no fallback path).
"""
from __future__ import annotations

import math
import os
import typing as t

import numpy as np
import torch
from torch import nn

from . import _lib
from . import functional as VF

_CORES: t.Dict[str, t.Any] = {}
_READOUTS: t.Dict[str, t.Any] = {}
_CRITERION: t.Dict[str, t.Any] = {}


def _registrar(table):
    def register(name):
        def add(cls):
            table[name] = cls
            return cls

        return add

    return register


register_core = _registrar(_CORES)
register_readout = _registrar(_READOUTS)
register_criterion = _registrar(_CRITERION)


def get_core(args):
    if args.core not in _CORES:
        raise NotImplementedError(f"Core {args.core} has not been implemented.")
    return _CORES[args.core]


# ------------------------------------------------------------------------------------------------------
# core
# ------------------------------------------------------------------------------------------------------
class Core(nn.Module):
    def __init__(self, args, input_shape, name: str = "Core"):
        super().__init__()
        self.input_shape = input_shape
        self.name = name
        self.behavior_mode = args.behavior_mode
        self.frozen = False
        self.verbose = getattr(args, "verbose", 0)

    def freeze(self):
        for p in self.parameters():
            p.requires_grad_(False)
        self.frozen = True

    def unfreeze(self):
        for p in self.parameters():
            p.requires_grad_(True)
        self.frozen = False

    def regularizer(self):
        raise NotImplementedError


class DropPath(nn.Module):
    """Holds the ``keep_prop`` buffer of the reference (models/utils.py:121-141); only drop_path = 0 is in scope."""

    def __init__(self, dropout: float = 0.0):
        super().__init__()
        if dropout != 0:
            raise NotImplementedError("v1t_b200: --drop_path > 0 is outside the hot-path scope (default 0)")
        self.register_buffer("keep_prop", torch.tensor(1.0 - dropout))

    def forward(self, inputs):
        return inputs


class Image2Patches(nn.Module):
    def __init__(self, image_shape, patch_mode: int, patch_size: int, stride: int, emb_dim: int, dropout: float = 0.0):
        super().__init__()
        if patch_mode != 0:
            raise NotImplementedError(f"v1t_b200: --patch_mode {patch_mode} is out of scope (only 0: nn.Unfold + Linear)")
        if not 1 <= stride <= patch_size:
            raise AssertionError("need 1 <= stride <= patch_size")
        c, h, w = image_shape
        self.input_shape = image_shape
        n = (math.floor((h - patch_size) / stride) + 1) * (math.floor((w - patch_size) / stride) + 1)
        # indices 0/1 are the parameter-free unfold / rearrange steps; index 2 owns the weights (key "projection.2.*")
        self.projection = nn.Sequential(nn.Unfold(kernel_size=patch_size, stride=stride), nn.Identity(),
                                        nn.Linear(patch_size * patch_size * c, emb_dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, emb_dim))
        self.num_patches = n + 1
        self.pos_embedding = nn.Parameter(torch.randn(self.num_patches, emb_dim))
        self.dropout = nn.Dropout(p=dropout)
        self.output_shape = (self.num_patches, emb_dim)


class MLP(nn.Module):
    def __init__(self, in_dim: int, hidden_dim: int, out_dim: int = None, dropout: float = 0.0, use_bias: bool = True):
        super().__init__()
        out_dim = in_dim if out_dim is None else out_dim
        self.model = nn.Sequential(nn.LayerNorm(in_dim), nn.Linear(in_dim, hidden_dim, bias=use_bias), nn.GELU(),
                                   nn.Dropout(p=dropout), nn.Linear(hidden_dim, out_dim, bias=use_bias),
                                   nn.Dropout(p=dropout))


class BehaviorMLP(nn.Module):
    def __init__(self, behavior_mode: int, out_dim: int, dropout: float = 0.0, mouse_ids=None, use_bias: bool = True):
        super().__init__()
        assert behavior_mode in (2, 3, 4)
        self.behavior_mode = behavior_mode
        in_dim = 3 if behavior_mode == 2 else 5
        keys = list(mouse_ids) if behavior_mode == 4 else ["share"]
        self.models = nn.ModuleDict({
            k: nn.Sequential(nn.Linear(in_dim, out_dim // 2, bias=use_bias), nn.Tanh(), nn.Dropout(p=dropout),
                             nn.Linear(out_dim // 2, out_dim, bias=use_bias), nn.Tanh())
            for k in keys})

    def key(self, mouse_id: str) -> str:
        return mouse_id if self.behavior_mode == 4 else "share"


class _Attend(nn.Module):
    """Stands in for the reference's ``Attention.attend`` (nn.Softmax): forward hooks registered on it (as
    attention_rollout.Recorder does, attention_rollout.py:31-36) receive the softmax probabilities, which
    ViTCore.forward materialises on demand through v1t_attention_probs."""

    def forward(self, probs):
        return probs


class Attention(nn.Module):
    def __init__(self, num_patches: int, emb_dim: int, num_heads: int = 8, dropout: float = 0.0,
                 use_lsa: bool = False, use_bias: bool = True, grad_checkpointing: bool = False):
        super().__init__()
        if use_lsa:
            raise NotImplementedError("v1t_b200: --use_lsa is out of scope (default off)")
        self.grad_checkpointing = grad_checkpointing  # accepted, ignored: the fused kernels recompute by design
        inner = emb_dim * num_heads  # head_dim == emb_dim (vit.py:218)
        self.layer_norm = nn.LayerNorm(emb_dim)
        self.to_qkv = nn.Linear(emb_dim, inner * 3, bias=False)
        self.attend = _Attend()
        self.dropout = nn.Dropout(p=dropout)
        self.projection = nn.Sequential(nn.Linear(inner, emb_dim, bias=use_bias), nn.Dropout(p=dropout))
        self.mask = None
        self.register_buffer("scale", torch.tensor(emb_dim ** -0.5))


class Transformer(nn.Module):
    def __init__(self, input_shape, emb_dim, num_blocks, num_heads, mlp_dim, dropout, behavior_mode, mouse_ids,
                 use_lsa=False, drop_path=0.0, use_bias=True, grad_checkpointing=False):
        super().__init__()
        self.blocks = nn.ModuleList()
        for _ in range(num_blocks):
            block = nn.ModuleDict({
                "mha": Attention(input_shape[0], emb_dim, num_heads, dropout, use_lsa, use_bias, grad_checkpointing),
                "mlp": MLP(emb_dim, mlp_dim, dropout=dropout, use_bias=use_bias),
            })
            if behavior_mode in (2, 3, 4):
                block["b-mlp"] = BehaviorMLP(behavior_mode, emb_dim, mouse_ids=mouse_ids, use_bias=use_bias)
            self.blocks.append(block)
        self.drop_path = DropPath(dropout=drop_path)
        self.output_shape = (input_shape[0], emb_dim)
        self.apply(self._init)

    @staticmethod
    def _init(m):  # vit.py:338-346
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)


@register_core("vit")
class ViTCore(Core):
    """Drop-in for the reference ViTCore (vit.py:365-436); forward/backward in hand-written sm_100a CUDA."""

    def __init__(self, args, input_shape, name: str = "ViTCore"):
        super().__init__(args, input_shape=input_shape, name=name)
        self.register_buffer("reg_scale", torch.tensor(args.core_reg_scale))
        self.behavior_mode = args.behavior_mode
        if not hasattr(args, "grad_checkpointing"):
            args.grad_checkpointing = False
        elif args.grad_checkpointing is None:
            args.grad_checkpointing = "cuda" in torch.device(args.device).type
        self.patch_embedding = Image2Patches(input_shape, args.patch_mode, args.patch_size, args.patch_stride,
                                             args.emb_dim, args.p_dropout)
        self.transformer = Transformer(
            input_shape=self.patch_embedding.output_shape, emb_dim=args.emb_dim, num_blocks=args.num_blocks,
            num_heads=args.num_heads, mlp_dim=args.mlp_dim, dropout=args.t_dropout, behavior_mode=self.behavior_mode,
            mouse_ids=list(args.output_shapes.keys()), use_lsa=args.use_lsa, drop_path=args.drop_path,
            use_bias=not args.disable_bias, grad_checkpointing=bool(args.grad_checkpointing))
        h, w = self.find_shape(self.patch_embedding.num_patches - 1)
        self.output_shape = (args.emb_dim, h, w)
        self.p_dropout, self.t_dropout = float(args.p_dropout), float(args.t_dropout)
        if args.num_blocks > _lib.V1T_MAX_BLOCKS:
            raise NotImplementedError(f"v1t_b200: at most {_lib.V1T_MAX_BLOCKS} blocks")
        impl = getattr(args, "b200_impl", None) or os.environ.get("V1T_IMPL", "bf16x3")
        c, ih, iw = input_shape
        bdim = {0: 0, 1: 0, 2: 3, 3: 5, 4: 5}[self.behavior_mode]
        self.spec = VF.CoreSpec(in_ch=c, in_h=ih, in_w=iw, patch=args.patch_size, stride=args.patch_stride,
                                emb=args.emb_dim, heads=args.num_heads, mlp=args.mlp_dim, blocks=args.num_blocks,
                                bdim=bdim, impl=_lib.IMPL_NAMES[impl] if isinstance(impl, str) else int(impl))
        self.dropout_seed: t.Optional[int] = None  # set to pin the dropout masks (tests)
        self.grad_sink: t.Optional[VF.GradSink] = None  # see fused_grad_accumulation()
        self.last_dropout_seed = 0

    @staticmethod
    def find_shape(num_patches: int):
        d1 = math.ceil(math.sqrt(num_patches))
        while num_patches % d1 != 0 and d1 > 0:
            d1 -= 1
        return d1, num_patches // d1

    @property
    def attention_path(self) -> str:
        """"fused" (tcgen05 flash-style kernels, head dim <= 160) or "materialised" (batched GEMMs around a [chunk,H,T,T]
        probability buffer: the fp32 impl, and head dims the 512 TMEM columns cannot hold) — decided by the library
        from shape and impl, reported so that nothing is switched silently."""
        return "fused" if self.spec.dims().attn_path == _lib.ATTN_FUSED else "materialised"

    def regularizer(self):
        """L1 over all core parameters (vit.py:419-421)."""
        return self.reg_scale * sum(p.abs().sum() for p in self.parameters())

    def flat_params(self, mouse_id: str) -> t.List[t.Optional[torch.Tensor]]:
        """Parameters in the order of the C-ABI structs (v1t_core_ptrs / v1t_block_ptrs)."""
        pe = self.patch_embedding
        out = [pe.cls_token, pe.pos_embedding, pe.projection[2].weight, pe.projection[2].bias]
        for blk in self.transformer.blocks:
            mha, mlp = blk["mha"], blk["mlp"].model
            out += [mha.layer_norm.weight, mha.layer_norm.bias, mha.to_qkv.weight, mha.projection[0].weight,
                    mha.projection[0].bias, mlp[0].weight, mlp[0].bias, mlp[1].weight, mlp[1].bias, mlp[4].weight,
                    mlp[4].bias]
            if "b-mlp" in blk:
                seq = blk["b-mlp"].models[blk["b-mlp"].key(mouse_id)]
                out += [seq[0].weight, seq[0].bias, seq[3].weight, seq[3].bias]
            else:
                out += [None, None, None, None]
        return out

    def fused_grad_accumulation(self, on: bool = True):
        """Arm (or disarm) the flat gradient sink of the core parameters: while armed, every backward through the
        core adds all its parameter gradients into ``p.grad`` with one kernel instead of one per tensor."""
        if not on or self.frozen:
            if self.grad_sink is not None:
                self.grad_sink.disarm()
            return None
        if self.grad_sink is None or [id(p) for p in self.grad_sink.params] != [id(p) for p in self.parameters()
                                                                                   if p.requires_grad]:
            self.grad_sink = VF.GradSink(list(self.parameters()))
        self.grad_sink.arm()
        return self.grad_sink

    def _hooked_attends(self):
        return [(i, blk["mha"].attend) for i, blk in enumerate(self.transformer.blocks)
                if len(blk["mha"].attend._forward_hooks) > 0]

    def forward(self, inputs: torch.Tensor, mouse_id: str, behaviors: torch.Tensor, pupil_centers: torch.Tensor):
        if self.behavior_mode in (3, 4):
            beh = torch.cat((behaviors, pupil_centers), dim=-1)  # vit.py:431-432
        elif self.behavior_mode == 2:
            beh = behaviors
        else:
            beh = None
        training = self.training
        p_tok = self.p_dropout if training else 0.0
        p_blk = self.t_dropout if training else 0.0
        seed = 0
        if p_tok > 0 or p_blk > 0:
            seed = self.dropout_seed if self.dropout_seed is not None else int(
                torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        self.last_dropout_seed = seed
        hooked = self._hooked_attends()
        keep = {} if hooked else None
        tokens = VF.core_forward(self.spec, inputs, beh, self.flat_params(mouse_id), p_tok, p_blk, seed, keep,
                                 sink=self.grad_sink)
        for i, attend in hooked:  # emit softmax probabilities for hooks (attention rollout)
            attend(VF.attention_probs(self.spec, keep, i))
        e, h, w = self.output_shape
        # drop CLS, 'b (h w) c -> b c h w' as a VIEW of the padded token buffer (vit.py:434-435)
        return tokens[:, 1:, :e].unflatten(1, (h, w)).permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------------
# readout
# ------------------------------------------------------------------------------------------------------
class Readout(nn.Module):
    def __init__(self, args, input_shape, output_shape, ds, name: str = None):
        super().__init__()
        self.name = "Readout" if name is None else name
        self.input_shape = input_shape
        self.output_shape = output_shape
        self.neuron_coordinates = ds.dataset.coordinates
        self.register_buffer("reg_scale", torch.tensor(args.readout_reg_scale))

    @property
    def num_neurons(self):
        return self.output_shape[-1]

    def regularizer(self, reduction: str):
        return self.reg_scale * sum(p.abs().sum() for p in self.parameters())


@register_readout("gaussian2d")
class Gaussian2DReadout(Readout):
    """Drop-in for gaussian2d.py:13-278 ("full" Gaussian only — the only type any caller constructs)."""

    def __init__(self, args, input_shape, output_shape, ds, use_bias: bool = True, init_mu_range: float = 0.3,
                 init_sigma: float = 0.1, gaussian_type: str = "full", name: str = "Gaussian2DReadout"):
        super().__init__(args, input_shape=input_shape, output_shape=output_shape, ds=ds, name=name)
        if init_mu_range > 1.0 or init_mu_range <= 0.0 or init_sigma <= 0.0:
            raise ValueError("either init_mu_range doesn't belong to [0.0, 1.0] or init_sigma_range is non-positive")
        if gaussian_type != "full":
            raise NotImplementedError("v1t_b200: only gaussian_type='full' is in scope")
        self.init_mu_range, self.init_sigma, self.gaussian_type = init_mu_range, init_sigma, gaussian_type
        n = self.num_neurons
        self.grid_shape = (1, n, 1, 2)
        self._predicted_grid = False
        self._original_grid = True
        if args.disable_grid_predictor:
            self._mu = nn.Parameter(torch.empty(*self.grid_shape))
        else:
            self._init_grid_predictor(self.neuron_coordinates, args.grid_predictor_dim)
        self.sigma_shape = (1, n, 2, 2)
        self.sigma = nn.Parameter(torch.empty(*self.sigma_shape))
        c = self.input_shape[0]
        self._original_features = True
        self._shared_features = False
        self.features = nn.Parameter(torch.empty(1, c, 1, n))
        self.use_bias = use_bias
        self.bias_mode = args.bias_mode
        self._initialize(ds)

    def _init_grid_predictor(self, source_grid, input_dimensions: int = 2, hidden_features: int = 30):
        self._original_grid = False
        sg = np.asarray(source_grid)[:, :input_dimensions]
        self.mu_transform = nn.Sequential(nn.Linear(sg.shape[1], hidden_features), nn.ELU(),
                                          nn.Linear(hidden_features, 2), nn.Tanh())
        sg = sg - sg.mean(axis=0, keepdims=True)
        sg = sg / np.abs(sg).max()
        self.register_buffer("source_grid", torch.from_numpy(sg.astype(np.float32)))
        self._predicted_grid = True

    def _initialize(self, ds):  # gaussian2d.py:171-186
        if not self._predicted_grid or self._original_grid:
            self._mu.data.uniform_(-self.init_mu_range, self.init_mu_range)
        self.sigma.data.uniform_(-self.init_sigma, self.init_sigma)
        self.features.data.fill_(1 / self.input_shape[0])
        stats = ds.dataset.response_stats
        if self.use_bias:
            if self.bias_mode == 0:
                bias = torch.zeros(len(stats["mean"]))
            elif self.bias_mode == 1:
                bias = torch.from_numpy(np.asarray(stats["mean"], dtype=np.float32))
            elif self.bias_mode == 2:
                bias = torch.from_numpy(np.asarray(stats["mean"] / stats["std"], dtype=np.float32))
            else:
                raise NotImplementedError(f"Gaussian2dReadout: bias mode {self.bias_mode} has not been implemented.")
            self.bias = nn.Parameter(bias)
        else:
            self.bias = None

    def feature_l1(self, reduction="sum"):
        l1 = self.features.abs()
        return l1.sum() if reduction == "sum" else (l1.mean() if reduction == "mean" else l1)

    def regularizer(self, reduction="sum"):
        return self.reg_scale * self.feature_l1(reduction=reduction)

    @property
    def mu(self):
        if self._predicted_grid:  # Linear-ELU-Linear-Tanh over the neurons: one kernel each way (gaussian2d.py:188-193)
            mt = self.mu_transform
            return VF.small_mlp(self.source_grid, (mt[0], mt[2]), ("elu", "tanh")).view(*self.grid_shape)
        return self._mu

    def forward(self, inputs: torch.Tensor, sample: bool = None, shifts: torch.Tensor = None,
                noise: torch.Tensor = None):
        """inputs [B,C,h,w] -> pre-activation responses [B,N].  ``noise`` ([B,N,2] standard normal) may be
        injected for reproducible train-mode runs; by default it is drawn like the reference does
        (gaussian2d.py:219-235).  The position clamp/sampling/bilinear gather/feature dot run in one kernel."""
        b, c = inputs.shape[0], inputs.shape[1]
        n = self.num_neurons
        sample = self.training if sample is None else sample
        if noise is None and sample:
            noise = torch.empty((b, n, 1, 2), dtype=torch.float32, device=inputs.device).normal_()
        if noise is not None:
            noise = noise.reshape(b, n, 2)
        if not self._predicted_grid:
            # the reference clamps the `_mu` PARAMETER in place on every forward (gaussian2d.py:212-215; with the grid
            # predictor the same line acts on a temporary and is a no-op)
            with torch.no_grad():
                self._mu.clamp_(min=-1, max=1)
        return VF.readout_forward(inputs, self.mu.view(n, 2), self.sigma.view(n, 2, 2), noise, shifts,
                                  self.features.view(c, n), self.bias)


class Readouts(nn.ModuleDict):
    def __init__(self, args, model: str, input_shape, output_shapes, ds):
        super().__init__()
        if model not in _READOUTS:
            raise NotImplementedError(f"Readout {model} has not been implemented.")
        self.input_shape = input_shape
        self.output_shapes = output_shapes
        for mouse_id, output_shape in output_shapes.items():
            self.add_module(mouse_id, _READOUTS[model](args, input_shape=input_shape, output_shape=output_shape,
                                                       ds=ds[mouse_id], name=f"Mouse{mouse_id}Readout"))

    def regularizer(self, mouse_id, reduction: str = "sum"):
        return self[str(mouse_id)].regularizer(reduction=reduction)

    def forward(self, inputs, mouse_id: str, shifts=None, **kw):
        return self[mouse_id](inputs, shifts=shifts, **kw)


class ELU1(nn.Module):
    def __init__(self):
        super().__init__()
        self.elu = nn.ELU()
        self.register_buffer("one", torch.tensor(1.0))

    def forward(self, inputs):
        return VF.elu1(inputs)


# ------------------------------------------------------------------------------------------------------
# criterion
# ------------------------------------------------------------------------------------------------------
class Loss(nn.Module):
    def __init__(self, args, ds, reduction: str = "sum"):
        super().__init__()
        self.reduction = reduction
        self.ds_scale = args.ds_scale
        self._ds_sizes = {m: float(len(d.dataset)) for m, d in ds.items()}

    def scale_of(self, mouse_id: str, batch_size: int) -> float:
        return math.sqrt(self._ds_sizes[mouse_id] / batch_size) if self.ds_scale else 1.0


@register_criterion("poisson")
class PoissonLoss(Loss):
    def __init__(self, args, ds, eps: float = VF.EPS_F32, reduction: str = "sum"):
        super().__init__(args, ds=ds, reduction=reduction)
        self.register_buffer("eps", torch.tensor(eps))
        self._eps = float(eps)

    def forward(self, y_true, y_pred, mouse_id: str, batch_size: int = None):
        if batch_size is None:
            batch_size = y_true.size(0)
        return VF.poisson_loss(y_pred, y_true, self._eps, self.scale_of(mouse_id, batch_size))


def get_criterion(args, ds):
    assert args.criterion in _CRITERION, f"Criterion {args.criterion} not found."
    crit = _CRITERION[args.criterion](args, ds=ds)
    crit.to(args.device)
    return crit


# ------------------------------------------------------------------------------------------------------
# callers either side of the path: shifters and cropper, native through functional.small_mlp / crop_resize
# ------------------------------------------------------------------------------------------------------
class CoreShifter(nn.Module):
    def __init__(self, args, in_features: int, hidden_features: int, num_layers: int, name: str = "CoreShifter"):
        super().__init__()
        self.name = name
        self.register_buffer("reg_scale", torch.tensor(args.shifter_reg_scale))
        layers, width = [], in_features
        for _ in range(num_layers - 1):
            layers += [nn.Linear(width, hidden_features), nn.Tanh()]
            width = hidden_features
        layers += [nn.Linear(width, 2), nn.Tanh()]
        self.mlp = nn.Sequential(*layers)

    def regularizer(self):
        return self.reg_scale * sum(p.abs().sum() for p in self.parameters())

    def forward(self, pupil_center):  # Linear-Tanh stack (core_shifter.py:38-39): one kernel each way
        linears = [m for m in self.mlp if isinstance(m, nn.Linear)]
        return VF.small_mlp(pupil_center, linears, ("tanh",) * len(linears))


class CoreShifters(nn.ModuleDict):
    def __init__(self, args, mouse_ids, input_channels: int, hidden_features: int, num_layers: int):
        super().__init__()
        for m in mouse_ids:
            self.add_module(m, CoreShifter(args, input_channels, hidden_features, num_layers, name=f"Mouse{m}CoreShifter"))

    def regularizer(self, mouse_id: str):
        return self[mouse_id].regularizer()

    def forward(self, pupil_centers, mouse_id: str):
        return self[mouse_id](pupil_centers)


class ImageShifter(nn.Module):
    """image_cropper.py:10-48: Linear-Tanh stack on the pupil centre (+ behaviours for shift_mode 4), scaled by
    max_shift."""

    def __init__(self, args, max_shift: float, hidden_features: int = 10, num_layers: int = 1,
                 name: str = "ImageShifter"):
        super().__init__()
        assert 0 <= max_shift <= 1
        self.name = name
        self.shift_mode = args.shift_mode
        self.register_buffer("max_shift", torch.tensor(max_shift))
        self.register_buffer("reg_scale", torch.tensor(args.cropper_reg_scale))
        layers, width = [], 5 if self.shift_mode == 4 else 2
        for _ in range(num_layers - 1):
            layers += [nn.Linear(width, hidden_features), nn.Tanh()]
            width = hidden_features
        layers += [nn.Linear(width, 2), nn.Tanh()]
        self.mlp = nn.Sequential(*layers)

    def regularizer(self):
        return self.reg_scale * sum(p.abs().sum() for p in self.parameters())

    def forward(self, behaviors, pupil_centers):
        inputs = pupil_centers
        if self.shift_mode == 4:
            inputs = torch.cat((behaviors, pupil_centers), dim=-1)
        linears = [m for m in self.mlp if isinstance(m, nn.Linear)]
        return VF.small_mlp(inputs, linears, ("tanh",) * len(linears)) * self.max_shift


class ImageCropper(nn.Module):
    """The step before the hot path (image_cropper.py:51-140), all shift modes and crop scales: the crop (nearest
    sampling on the shifted grid), the 36x64 bilinear resize and the behaviour planes are one gather kernel
    (csrc/cropper.cu).  Same buffers / sub-module names as the reference so checkpoints load strictly."""

    def __init__(self, args, ds):
        super().__init__()
        self.shift_mode, self.input_shape, self.behavior_mode = args.shift_mode, args.input_shape, args.behavior_mode
        c, in_h, in_w = args.input_shape
        out_h, out_w = in_h, in_w
        if self.behavior_mode == 1:
            c += 3
        self.crop_scale = args.center_crop
        self.crop_h, self.crop_w = in_h, in_w
        if self.crop_scale < 1:
            out_h = self.crop_h = int(in_h * self.crop_scale)
            out_w = self.crop_w = int(in_w * self.crop_scale)
        ys = torch.linspace(-self.crop_scale, self.crop_scale, self.crop_h)
        xs = torch.linspace(-self.crop_scale, self.crop_scale, self.crop_w)
        my, mx = torch.meshgrid(ys, xs, indexing="ij")
        self.register_buffer("grid", torch.stack((mx, my), dim=2).unsqueeze(0))  # (x, y), image_cropper.py:104-112
        if self.shift_mode in (1, 3, 4):
            self.add_module("image_shifter", nn.ModuleDict({
                m: ImageShifter(args, max_shift=1 - self.crop_scale, num_layers=3, name=f"Mouse{m}ImageShifter")
                for m in list(ds.keys())}))
        else:
            self.image_shifter = None
        self.resize = None
        if getattr(args, "resize_image", 0) == 1 and getattr(args, "ds_name", "") != "franke2022":
            out_h, out_w = 36, 64
            self.resize = (out_h, out_w)
        self.output_shape = (c, out_h, out_w)

    def regularizer(self, mouse_id: str):
        return 0 if self.image_shifter is None else self.image_shifter[mouse_id].regularizer()

    def forward(self, inputs, mouse_id, behaviors, pupil_centers):
        grid = self.grid.expand(inputs.size(0), -1, -1, -1)
        shifts = None
        if self.image_shifter is not None:
            shifts = self.image_shifter[mouse_id](behaviors=behaviors, pupil_centers=pupil_centers)
            grid = grid + shifts[:, None, None, :]
        if shifts is None and self.resize is None and self.behavior_mode != 1 and self.crop_scale == 1:
            return inputs, grid  # nearest sampling on the identity grid returns the image itself
        outputs = VF.crop_resize(inputs, self.grid, None if shifts is None else shifts.detach(),
                                 self.output_shape[1:], behaviors if self.behavior_mode == 1 else None)
        return outputs, grid


class Model(nn.Module):
    """Same wiring and state-dict keys as the reference Model (models/model.py:50-177)."""

    def __init__(self, args, ds, name: str = "Model"):
        super().__init__()
        assert isinstance(args.output_shapes, dict)
        self.name = name
        self.input_shape, self.output_shapes, self.shift_mode = args.input_shape, args.output_shapes, args.shift_mode
        self.add_module("image_cropper", ImageCropper(args, ds=ds))
        self.add_module("core", get_core(args)(args, input_shape=self.image_cropper.output_shape))
        if self.shift_mode in (2, 3, 4):
            self.add_module("core_shifter", CoreShifters(args, list(ds.keys()), 2, 5, 3))
        else:
            self.core_shifter = None
        self.add_module("readouts", Readouts(args, model=args.readout, input_shape=self.core.output_shape,
                                             output_shapes=self.output_shapes, ds=ds))
        self.elu1 = ELU1()

    @property
    def device(self):
        return next(self.parameters()).device

    def get_parameters(self, core_lr: float):
        params = []
        if not self.core.frozen:
            params.append({"params": self.core.parameters(), "lr": core_lr, "name": "core"})
        params.append({"params": self.readouts.parameters(), "name": "readouts"})
        if self.image_cropper.image_shifter is not None:
            params.append({"params": self.image_cropper.parameters(), "name": "image_cropper"})
        if self.core_shifter is not None:
            params.append({"params": self.core_shifter.parameters(), "name": "core_shifter"})
        return params

    def regularizer(self, mouse_id: str):
        reg = 0
        if not self.core.frozen:
            reg = reg + self.core.regularizer()
        reg = reg + self.readouts.regularizer(mouse_id=mouse_id)
        reg = reg + self.image_cropper.regularizer(mouse_id=mouse_id)
        if self.core_shifter is not None:
            reg = reg + self.core_shifter.regularizer(mouse_id=mouse_id)
        return reg

    def forward(self, inputs, mouse_id: str, behaviors, pupil_centers, activate: bool = True, noise=None):
        images, grids = self.image_cropper(inputs, mouse_id=mouse_id, behaviors=behaviors, pupil_centers=pupil_centers)
        outputs = self.core(images, mouse_id=mouse_id, behaviors=behaviors, pupil_centers=pupil_centers)
        shifts = None
        if self.core_shifter is not None:
            shifts = self.core_shifter(pupil_centers, mouse_id=mouse_id)
        kw = {} if noise is None else {"noise": noise}
        outputs = self.readouts(outputs, mouse_id=mouse_id, shifts=shifts, **kw)
        if activate:
            outputs = self.elu1(outputs)
        return outputs, images, grids
