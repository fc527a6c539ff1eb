"""torch.autograd.Function wrappers over the C-ABI (the only place pointers cross the boundary).

Tensors stay owned by PyTorch (device memory, caching allocator, current stream); the native library
launches the kernels.  No function here has a CPU path: non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import BLOCK_FIELDS, CoreDims, CorePtrs, CoreShape, ReadoutShape

EPS_F32 = float(torch.finfo(torch.float32).eps)

# one reusable scratch arena per (device, stream): kernels of one stream run in order, so a stream can reuse its arena
# from call to call; members of an ensemble running on side streams each get their own (the library never allocates)
_SCRATCH = {}


def _scratch(device: torch.device, nbytes: int) -> torch.Tensor:
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0)
    buf = _SCRATCH.get(key)
    if buf is None or buf.numel() < nbytes:
        _SCRATCH.pop(key, None)
        buf = None
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _SCRATCH[key] = buf
    return buf


def release_scratch():
    _SCRATCH.clear()


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("v1t_b200 is CUDA-only (sm_100a kernels, no CPU path): got a tensor on "
                               f"{t.device}; move the model and inputs to a B200 first")


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


@dataclass(frozen=True)
class CoreSpec:
    """Static description of a ViT core (the shape-defining args of vit.py:374-405)."""

    in_ch: int
    in_h: int
    in_w: int
    patch: int
    stride: int
    emb: int
    heads: int
    mlp: int
    blocks: int
    bdim: int
    impl: int = _lib.IMPL_FP32

    def shape(self, batch: int, p_tokens: float = 0.0, p_block: float = 0.0, seed: int = 0) -> CoreShape:
        return CoreShape(batch=batch, in_ch=self.in_ch, in_h=self.in_h, in_w=self.in_w, patch=self.patch,
                         stride=self.stride, emb=self.emb, heads=self.heads, mlp=self.mlp, blocks=self.blocks,
                         bdim=self.bdim, impl=self.impl, p_drop_tokens=p_tokens, p_drop_block=p_block,
                         seed=seed & 0xFFFFFFFFFFFFFFFF)

    def dims(self) -> CoreDims:
        d = CoreDims()
        _lib.check(_lib.load().v1t_core_dims_of(C.byref(self.shape(1)), C.byref(d)), "core_dims_of")
        return d


N_HEAD_PARAMS = 4
N_BLOCK_PARAMS = len(BLOCK_FIELDS)


def _fill_ptrs(tensors: Sequence[Optional[torch.Tensor]], blocks: int) -> CorePtrs:
    p = CorePtrs()
    p.cls, p.pos, p.wpe, p.bpe = (_ptr(t) for t in tensors[:N_HEAD_PARAMS])
    for i in range(blocks):
        chunk = tensors[N_HEAD_PARAMS + i * N_BLOCK_PARAMS: N_HEAD_PARAMS + (i + 1) * N_BLOCK_PARAMS]
        for name, t in zip(BLOCK_FIELDS, chunk):
            setattr(p.blk[i], name, _ptr(t))
    return p


class GradSink:
    """One flat fp32 buffer holding the gradients of a set of parameters, whose ``.grad`` are views into it.

    Autograd accumulates a parameter used by several backward passes with one small ``add`` kernel per parameter per
    pass (the shared core: 52 tensors x 7 mice per optimizer step, train.py:97-111).  With a sink armed, the core's
    backward writes all its parameter gradients into one temporary flat buffer, adds it to the sink with ONE kernel
    and reports no gradient to autograd; ``p.grad`` (a view of the sink) sees the sum.  Used by the repo's own step
    function (parallel.sweep(fused_accumulate=True)); plain ``loss.backward()`` on a model is unchanged."""

    ALIGN = 4  # floats: keeps every view 16-byte aligned for the fused optimizer's 128-bit accesses

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        self.offsets = {}
        off = 0
        for p in self.params:
            self.offsets[id(p)] = (off, p.numel())
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = off
        self.flat: Optional[torch.Tensor] = None
        self.armed = False

    def view_of(self, p: torch.Tensor, flat: torch.Tensor) -> torch.Tensor:
        off, n = self.offsets[id(p)]
        return flat[off:off + n].view(p.shape)

    def arm(self):
        """Make every parameter's .grad the sink's view (existing gradients are carried over, None becomes 0)."""
        if not self.params:
            return
        dev = self.params[0].device
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
            fresh = True
        else:
            fresh = False
        if not fresh and all(p.grad is None for p in self.params):
            self.flat.zero_()
            fresh = True
        for p in self.params:
            view = self.view_of(p, self.flat)
            if p.grad is None:
                if not fresh:
                    view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view
        self.armed = True

    def disarm(self):
        self.armed = False

    def slots(self, tensors: Sequence[Optional[torch.Tensor]]):
        """(offset, numel) of each tensor of a call's parameter list (None where absent / not in the sink)."""
        return [self.offsets.get(id(t)) if t is not None else None for t in tensors]


class _CoreFunction(torch.autograd.Function):
    """tokens[B,T,emb_ld] = ViTCore(images, behaviors; params)   (vit.py:423-436 without the final view)."""

    @staticmethod
    def forward(ctx, spec: CoreSpec, p_tokens: float, p_block: float, seed: int, keep_saved, sink, want_grad: bool,
                images, behaviors, *params):
        lib = _lib.load()
        _need_cuda(images, behaviors, *params)
        images = _f32c(images)
        behaviors = _f32c(behaviors)
        params = [_f32c(p) for p in params]
        B = images.shape[0]
        shape = spec.shape(B, p_tokens, p_block, seed)
        dims = spec.dims()
        dev = images.device
        # `want_grad` is evaluated by core_forward() OUTSIDE this Function: ctx.needs_input_grad ignores
        # torch.no_grad() and grad mode always reads as disabled in here.  `keep_saved` (a dict, possibly empty) asks
        # for the saved activations regardless (attention-map hooks on a frozen / no_grad model).
        keep = bool(want_grad) or keep_saved is not None
        saved = None
        if keep:
            saved = torch.empty(lib.v1t_core_saved_bytes(C.byref(shape)), dtype=torch.uint8, device=dev)
        scratch = _scratch(dev, lib.v1t_core_scratch_bytes(C.byref(shape)))
        tokens = torch.empty((B, dims.tokens, dims.emb_ld), dtype=torch.float32, device=dev)
        ptrs = _fill_ptrs(params, spec.blocks)
        with torch.cuda.device(dev):
            rc = lib.v1t_core_forward(C.byref(shape), C.byref(ptrs), images.data_ptr(), _ptr(behaviors),
                                      tokens.data_ptr(), _ptr(saved), scratch.data_ptr(), int(keep),
                                      _stream_ptr(dev))
        _lib.check(rc, "core_forward")
        ctx.spec, ctx.shape_args = spec, (B, p_tokens, p_block, seed)
        ctx.saved_buf = saved
        ctx.has_beh = behaviors is not None
        ctx.save_for_backward(images, *([behaviors] if behaviors is not None else []),
                              *[p for p in params if p is not None])
        ctx.param_present = [p is not None for p in params]
        ctx.sink = sink  # (GradSink, slots) or None
        if isinstance(keep_saved, dict):
            keep_saved["saved"], keep_saved["shape"] = saved, shape
        return tokens

    @staticmethod
    def backward(ctx, d_tokens):
        lib = _lib.load()
        spec = ctx.spec
        saved_t = list(ctx.saved_tensors)
        images = saved_t.pop(0)
        behaviors = saved_t.pop(0) if ctx.has_beh else None
        params: List[Optional[torch.Tensor]] = [saved_t.pop(0) if present else None for present in ctx.param_present]
        shape = spec.shape(*ctx.shape_args)
        dev = images.device
        d_tokens = d_tokens.contiguous().clone()  # clobbered by the library
        needs = ctx.needs_input_grad[9:]
        flat_tmp = None
        if ctx.sink is not None and ctx.sink[0].armed:
            sink, slots = ctx.sink
            if any(p is not None and need and slot is None for p, need, slot in zip(params, needs, slots)):
                raise RuntimeError("core backward: a parameter that needs a gradient is not in the armed GradSink")
            flat_tmp = torch.zeros_like(sink.flat)  # absent parameters (other mice's b-mlp) contribute zero
            grads = [flat_tmp[slot[0]:slot[0] + slot[1]].view(p.shape) if (p is not None and need) else None
                     for p, need, slot in zip(params, needs, slots)]
        else:
            grads = [torch.empty_like(p) if (p is not None and need) else None for p, need in zip(params, needs)]
        d_images = torch.empty_like(images) if ctx.needs_input_grad[7] else None
        scratch = _scratch(dev, lib.v1t_core_scratch_bytes(C.byref(shape)))
        pptr, gptr = _fill_ptrs(params, spec.blocks), _fill_ptrs(grads, spec.blocks)
        with torch.cuda.device(dev):
            rc = lib.v1t_core_backward(C.byref(shape), C.byref(pptr), images.data_ptr(), _ptr(behaviors),
                                       d_tokens.data_ptr(), ctx.saved_buf.data_ptr(), scratch.data_ptr(),
                                       C.byref(gptr), _ptr(d_images), _stream_ptr(dev))
        _lib.check(rc, "core_backward")
        if flat_tmp is not None:
            ctx.sink[0].flat.add_(flat_tmp)  # one kernel for every core parameter; autograd sees no gradient
            grads = [None] * len(grads)
        return (None, None, None, None, None, None, None, d_images, None, *grads)


def core_forward(spec: CoreSpec, images, behaviors, params, p_tokens=0.0, p_block=0.0, seed=0, keep_saved=None,
                 sink: Optional[GradSink] = None):
    """Returns tokens [B, T, emb_ld] (fp32).  `params`: list in the order cls,pos,wpe,bpe + BLOCK_FIELDS per block.
    ``sink``: an armed GradSink receives the parameter gradients in one add instead of autograd's per-tensor adds."""
    sink_arg = (sink, sink.slots(params)) if (sink is not None and sink.armed) else None
    want_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (images, behaviors, *params))
    return _CoreFunction.apply(spec, float(p_tokens), float(p_block), int(seed), keep_saved, sink_arg, want_grad,
                               images, behaviors, *params)


def attention_probs(spec: CoreSpec, keep: dict, block: int) -> torch.Tensor:
    """softmax(QK^T/sqrt(E)) [B,H,T,T] of one block, from the saved activations of a core_forward call."""
    lib = _lib.load()
    shape, saved = keep["shape"], keep["saved"]
    d = spec.dims()
    out = torch.empty((shape.batch, spec.heads, d.tokens, d.tokens), dtype=torch.float32, device=saved.device)
    with torch.cuda.device(saved.device):
        rc = lib.v1t_attention_probs(C.byref(shape), saved.data_ptr(), block, out.data_ptr(),
                                     _stream_ptr(saved.device))
    _lib.check(rc, "attention_probs")
    return out


# ------------------------------------------------------------------------------------------------------
# readout
# ------------------------------------------------------------------------------------------------------
def _channel_last(fmap: torch.Tensor) -> torch.Tensor:
    """[B,C,h,w] tensor whose channel stride is 1 (the core emits exactly that view)."""
    if fmap.dtype != torch.float32:
        fmap = fmap.float()
    if fmap.stride(1) != 1:
        fmap = fmap.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    return fmap


def _readout_shape(fmap: torch.Tensor, n: int) -> ReadoutShape:
    B, Cc, h, w = fmap.shape
    return ReadoutShape(batch=B, neurons=n, channels=Cc, gh=h, gw=w, fs_b=fmap.stride(0), fs_y=fmap.stride(2),
                        fs_x=fmap.stride(3))


class _ReadoutFunction(torch.autograd.Function):
    """z[B,N] = Gaussian2DReadout(fmap; mu, sigma, noise, shifts, features, bias)  (gaussian2d.py:237-278)."""

    @staticmethod
    def forward(ctx, fmap, mu, sigma, noise, shifts, features, bias):
        lib = _lib.load()
        _need_cuda(fmap, mu, sigma, noise, shifts, features, bias)
        fmap = _channel_last(fmap)
        mu, sigma, noise, shifts, features, bias = map(_f32c, (mu, sigma, noise, shifts, features, bias))
        N = mu.shape[0]
        rs = _readout_shape(fmap, N)
        dev = fmap.device
        z = torch.empty((fmap.shape[0], N), dtype=torch.float32, device=dev)
        scratch = _scratch(dev, lib.v1t_readout_scratch_bytes(C.byref(rs)))
        with torch.cuda.device(dev):
            rc = lib.v1t_readout_forward(C.byref(rs), fmap.data_ptr(), mu.data_ptr(), _ptr(sigma), _ptr(noise),
                                         _ptr(shifts), features.data_ptr(), _ptr(bias), None, 1.0, z.data_ptr(),
                                         None, None, scratch.data_ptr(), _stream_ptr(dev))
        _lib.check(rc, "readout_forward")
        ctx.flags = (sigma is not None, noise is not None, shifts is not None, bias is not None)
        ctx.save_for_backward(*[t for t in (fmap, mu, sigma, noise, shifts, features) if t is not None])
        return z

    @staticmethod
    def backward(ctx, dz):
        lib = _lib.load()
        has_sigma, has_noise, has_shifts, has_bias = ctx.flags
        sv = list(ctx.saved_tensors)
        fmap, mu = sv.pop(0), sv.pop(0)
        sigma = sv.pop(0) if has_sigma else None
        noise = sv.pop(0) if has_noise else None
        shifts = sv.pop(0) if has_shifts else None
        features = sv.pop(0)
        N = mu.shape[0]
        rs = _readout_shape(fmap, N)
        dev = fmap.device
        dz = _f32c(dz)
        need = ctx.needs_input_grad
        d_fmap = None
        if need[0]:
            # same strides as fmap so the channel-last gather/scatter pattern is shared by fwd and bwd
            d_fmap = torch.empty_strided(fmap.shape, fmap.stride(), dtype=torch.float32, device=dev)
            d_fmap.zero_()
        d_mu = torch.empty_like(mu) if need[1] else None
        d_sigma = torch.empty_like(sigma) if (need[2] and sigma is not None) else None
        d_shifts = torch.empty_like(shifts) if (need[4] and shifts is not None) else None
        d_feat = torch.empty_like(features) if need[5] else None
        d_bias = torch.empty((N,), dtype=torch.float32, device=dev) if (need[6] and has_bias) else None
        scratch = _scratch(dev, lib.v1t_readout_scratch_bytes(C.byref(rs)))
        with torch.cuda.device(dev):
            rc = lib.v1t_readout_backward(C.byref(rs), fmap.data_ptr(), mu.data_ptr(), _ptr(sigma), _ptr(noise),
                                          _ptr(shifts), features.data_ptr(), None, dz.data_ptr(), None, 1.0, 1.0,
                                          _ptr(d_fmap), _ptr(d_mu), _ptr(d_sigma), _ptr(d_shifts), _ptr(d_feat),
                                          _ptr(d_bias), scratch.data_ptr(), _stream_ptr(dev))
        _lib.check(rc, "readout_backward")
        if d_sigma is not None and noise is None:
            d_sigma.zero_()
        return d_fmap, d_mu, d_sigma, None, d_shifts, d_feat, d_bias


def readout_forward(fmap, mu, sigma, noise, shifts, features, bias):
    """fmap [B,C,h,w] (channel stride 1), mu [N,2], sigma [N,2,2], noise [B,N,2]|None, shifts [B,2]|None,
    features [C,N], bias [N]|None  ->  z [B,N]."""
    return _ReadoutFunction.apply(fmap, mu, sigma, noise, shifts, features, bias)


class _Elu1Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        lib = _lib.load()
        _need_cuda(z)
        z = _f32c(z)
        y = torch.empty_like(z)
        with torch.cuda.device(z.device):
            _lib.check(lib.v1t_elu1_forward(z.data_ptr(), y.data_ptr(), z.numel(), _stream_ptr(z.device)), "elu1_forward")
        ctx.save_for_backward(z)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        (z,) = ctx.saved_tensors
        dy = _f32c(dy)
        dz = torch.empty_like(z)
        with torch.cuda.device(z.device):
            _lib.check(lib.v1t_elu1_backward(z.data_ptr(), dy.data_ptr(), dz.data_ptr(), z.numel(),
                                             _stream_ptr(z.device)), "elu1_backward")
        return dz


def elu1(z):
    return _Elu1Function.apply(z)


class _PoissonFunction(torch.autograd.Function):
    """loss = scale * sum((y+eps) - (t+eps) log(y+eps))   (losses.py:153-166 + :114-119)."""

    @staticmethod
    def forward(ctx, y_pred, y_true, eps, scale):
        lib = _lib.load()
        _need_cuda(y_pred, y_true)
        y_pred, y_true = _f32c(y_pred), _f32c(y_true)
        if y_pred.shape != y_true.shape:
            raise RuntimeError(f"poisson loss: shape mismatch {tuple(y_pred.shape)} vs {tuple(y_true.shape)}")
        dev = y_pred.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        n = y_pred.numel()
        scratch = _scratch(dev, lib.v1t_poisson_scratch_bytes(n))
        with torch.cuda.device(dev):
            _lib.check(lib.v1t_poisson_forward(y_pred.data_ptr(), y_true.data_ptr(), n, eps, scale, loss.data_ptr(),
                                               scratch.data_ptr(), _stream_ptr(dev)), "poisson_forward")
        ctx.save_for_backward(y_pred, y_true)
        ctx.consts = (eps, scale)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        lib = _lib.load()
        y_pred, y_true = ctx.saved_tensors
        eps, scale = ctx.consts
        dloss = _f32c(dloss)
        dy = torch.empty_like(y_pred)
        with torch.cuda.device(y_pred.device):
            _lib.check(lib.v1t_poisson_backward(y_pred.data_ptr(), y_true.data_ptr(), y_pred.numel(), eps, scale,
                                                dloss.data_ptr(), dy.data_ptr(), _stream_ptr(y_pred.device)),
                       "poisson_backward")
        return dy, None, None, None


def poisson_loss(y_pred, y_true, eps: float = EPS_F32, scale: float = 1.0):
    return _PoissonFunction.apply(y_pred, y_true, float(eps), float(scale))


def dropout_mask(n: int, seed: int, site: int, p: float, device) -> torch.Tensor:
    """The multipliers the kernels apply at a dropout site (test / replay helper)."""
    lib = _lib.load()
    out = torch.empty((n,), dtype=torch.float32, device=device)
    with torch.cuda.device(out.device):
        _lib.check(lib.v1t_dropout_mask(out.data_ptr(), n, seed & 0xFFFFFFFFFFFFFFFF, site, p,
                                        _stream_ptr(out.device)), "dropout_mask")
    return out


def loss_scale(ds_size: float, batch_size: int, ds_scale: bool = True) -> float:
    return math.sqrt(ds_size / batch_size) if ds_scale else 1.0


# ------------------------------------------------------------------------------------------------------
# callers either side of the path: small MLPs, image cropper, attention rollout
# ------------------------------------------------------------------------------------------------------
_ACTS = {None: _lib.ACT_NONE, "none": _lib.ACT_NONE, "tanh": _lib.ACT_TANH, "elu": _lib.ACT_ELU}


def _mlp_spec(x: torch.Tensor, weights: Sequence[torch.Tensor], acts: Sequence[Optional[str]]) -> _lib.MlpSpec:
    if not 1 <= len(weights) <= _lib.MLP_MAX_LAYERS or len(acts) != len(weights):
        raise RuntimeError(f"small_mlp: {len(weights)} layers / {len(acts)} activations unsupported")
    spec = _lib.MlpSpec(rows=x.shape[0], layers=len(weights), x_ld=x.stride(0))
    spec.width[0] = weights[0].shape[1]
    for i, w in enumerate(weights):
        if w.shape[1] != spec.width[i]:
            raise RuntimeError(f"small_mlp: layer {i} expects {w.shape[1]} inputs, got {spec.width[i]}")
        spec.width[i + 1] = w.shape[0]
        spec.act[i] = _ACTS[acts[i]]
    if max(spec.width[: len(weights) + 1]) > _lib.MLP_MAX_WIDTH:
        raise NotImplementedError(f"small_mlp: widths {list(spec.width)} exceed {_lib.MLP_MAX_WIDTH}")
    if x.shape[1] < spec.width[0] or x.stride(1) != 1:
        raise RuntimeError("small_mlp: x must be [rows, >= in_features] with unit column stride")
    return spec


def _mlp_ptrs(weights, biases) -> _lib.MlpPtrs:
    p = _lib.MlpPtrs()
    for i, (w, b) in enumerate(zip(weights, biases)):
        p.w[i] = _ptr(w)
        p.b[i] = _ptr(b)
    return p


class _SmallMlpFunction(torch.autograd.Function):
    """y = act_L(Linear_L(... act_1(Linear_1(x)))) for the grid predictor (gaussian2d.py:102-136) and the shifters
    (core_shifter.py:24-40): one kernel forward, one backward that recomputes the activations."""

    @staticmethod
    def forward(ctx, x, acts, *params):
        lib = _lib.load()
        weights, biases = list(params[0::2]), list(params[1::2])
        _need_cuda(x, *weights, *biases)
        if x.dtype != torch.float32:
            x = x.float()
        if x.dim() != 2 or x.stride(1) != 1:
            x = x.reshape(x.shape[0], -1).contiguous()
        weights = [_f32c(w) for w in weights]
        biases = [_f32c(b) for b in biases]
        spec = _mlp_spec(x, weights, acts)
        dev = x.device
        y = torch.empty((x.shape[0], weights[-1].shape[0]), dtype=torch.float32, device=dev)
        ptrs = _mlp_ptrs(weights, biases)
        with torch.cuda.device(dev):
            _lib.check(lib.v1t_small_mlp_forward(C.byref(spec), C.byref(ptrs), x.data_ptr(), y.data_ptr(),
                                                 _stream_ptr(dev)), "small_mlp_forward")
        ctx.acts = tuple(acts)
        ctx.n_layers = len(weights)
        ctx.save_for_backward(x, *weights, *[b for b in biases if b is not None])
        ctx.has_bias = [b is not None for b in biases]
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        sv = list(ctx.saved_tensors)
        x = sv.pop(0)
        weights = [sv.pop(0) for _ in range(ctx.n_layers)]
        biases = [sv.pop(0) if h else None for h in ctx.has_bias]
        spec = _mlp_spec(x, weights, ctx.acts)
        dev = x.device
        dy = _f32c(dy)
        gw = [torch.empty_like(w) for w in weights]
        gb = [torch.empty_like(b) if b is not None else None for b in biases]
        if x.shape[0] == 0:
            for g in gw + [g for g in gb if g is not None]:
                g.zero_()
        else:
            scratch = _scratch(dev, lib.v1t_small_mlp_scratch_bytes(C.byref(spec)))
            pp, gp = _mlp_ptrs(weights, biases), _mlp_ptrs(gw, gb)
            with torch.cuda.device(dev):
                _lib.check(lib.v1t_small_mlp_backward(C.byref(spec), C.byref(pp), x.data_ptr(), dy.data_ptr(),
                                                      C.byref(gp), scratch.data_ptr(), _stream_ptr(dev)),
                           "small_mlp_backward")
        out = [None, None]
        for w, b in zip(gw, gb):
            out += [w, b]
        return tuple(out)


def small_mlp(x: torch.Tensor, layers: Sequence, acts: Sequence[Optional[str]]) -> torch.Tensor:
    """``layers``: nn.Linear modules (or (weight, bias) pairs); ``acts``: "tanh" | "elu" | None after each layer.
    x [rows, in] is data (no gradient flows into it)."""
    if x.requires_grad:
        raise NotImplementedError("small_mlp: gradients with respect to the input are not computed")
    params = []
    for l in layers:
        w, b = (l.weight, l.bias) if hasattr(l, "weight") else l
        params += [w, b]
    return _SmallMlpFunction.apply(x, tuple(acts), *params)


def crop_resize(images: torch.Tensor, grid: torch.Tensor, shifts: Optional[torch.Tensor], out_hw,
                behaviors: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ImageCropper.forward's tensor work (image_cropper.py:120-140): nearest crop on ``grid`` [1|-,crop_h,crop_w,2]
    (+ per-sample ``shifts`` [B,2]), bilinear resize to ``out_hw``, behaviour planes appended when given."""
    lib = _lib.load()
    _need_cuda(images, grid, shifts, behaviors)
    images, grid, shifts, behaviors = _f32c(images), _f32c(grid), _f32c(shifts), _f32c(behaviors)
    b, c, in_h, in_w = images.shape
    crop_h, crop_w = grid.shape[-3], grid.shape[-2]
    planes = 0 if behaviors is None else behaviors.shape[1]
    cs = _lib.CropShape(batch=b, channels=c, in_h=in_h, in_w=in_w, crop_h=crop_h, crop_w=crop_w, out_h=out_hw[0],
                        out_w=out_hw[1], behavior_planes=planes)
    out = torch.empty((b, c + planes, out_hw[0], out_hw[1]), dtype=torch.float32, device=images.device)
    with torch.cuda.device(images.device):
        _lib.check(lib.v1t_crop_resize(C.byref(cs), images.data_ptr(), grid.data_ptr(), _ptr(shifts), _ptr(behaviors),
                                       out.data_ptr(), _stream_ptr(images.device)), "crop_resize")
    return out


def attention_rollouts(attentions: torch.Tensor, image_shape, grid_hw) -> torch.Tensor:
    """attentions [B,L,H,T,T] -> heatmaps [B,*image_shape] (attention_rollout.py:92-133); grid_hw = (gh, gw) with
    gh*gw = T-1 (the reference's find_shape)."""
    lib = _lib.load()
    _need_cuda(attentions)
    if attentions.dim() != 5 or attentions.shape[-1] != attentions.shape[-2]:
        raise RuntimeError(f"attention_rollouts: expected [B,L,H,T,T], got {tuple(attentions.shape)}")
    a = _f32c(attentions)
    b, l, h, t, _ = a.shape
    dev = a.device
    out = torch.empty((b, int(image_shape[0]), int(image_shape[1])), dtype=torch.float32, device=dev)
    if b == 0:
        return out
    scratch = _scratch(dev, lib.v1t_rollout_scratch_bytes(b, t))
    with torch.cuda.device(dev):
        _lib.check(lib.v1t_attention_rollout(a.data_ptr(), b, l, h, t, int(grid_hw[0]), int(grid_hw[1]),
                                             out.shape[1], out.shape[2], out.data_ptr(), scratch.data_ptr(),
                                             _stream_ptr(dev)), "attention_rollout")
    return out


class _EnsembleFunction(torch.autograd.Function):
    """y = elu(sum_k w[k] x_k + b) + 1, or elu(mean_k x_k) + 1 when w is None (ensemble.py:30-80,131-151)."""

    @staticmethod
    def forward(ctx, weight, bias, *members):
        lib = _lib.load()
        _need_cuda(weight, bias, *members)
        if not 1 <= len(members) <= _lib.ENSEMBLE_MAX:
            raise NotImplementedError(f"ensemble: {len(members)} members outside 1..{_lib.ENSEMBLE_MAX}")
        members = [_f32c(m) for m in members]
        if any(m.shape != members[0].shape for m in members):
            raise RuntimeError("ensemble: members disagree on the output shape")
        weight, bias = _f32c(weight), _f32c(bias)
        if weight is not None and weight.numel() != len(members):
            raise RuntimeError(f"ensemble: weight has {weight.numel()} entries for {len(members)} members")
        tab = _lib.EnsembleMembers(count=len(members))
        for k, m in enumerate(members):
            tab.x[k] = m.data_ptr()
        y = torch.empty_like(members[0])
        dev = y.device
        with torch.cuda.device(dev):
            _lib.check(lib.v1t_ensemble_forward(C.byref(tab), _ptr(weight), _ptr(bias), y.numel(), y.data_ptr(),
                                                _stream_ptr(dev)), "ensemble_forward")
        ctx.has = (weight is not None, bias is not None)
        ctx.save_for_backward(*[t for t in (weight, bias) if t is not None], *members)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        sv = list(ctx.saved_tensors)
        weight = sv.pop(0) if ctx.has[0] else None
        bias = sv.pop(0) if ctx.has[1] else None
        members = sv
        if any(ctx.needs_input_grad[2:]):
            raise NotImplementedError("ensemble: the members are frozen (ensemble.py:106); no gradient flows into them")
        if weight is None:
            return (None, None) + (None,) * len(members)
        tab = _lib.EnsembleMembers(count=len(members))
        for k, m in enumerate(members):
            tab.x[k] = m.data_ptr()
        dev = dy.device
        dy = _f32c(dy)
        dw = torch.empty_like(weight)
        db = torch.empty_like(bias) if bias is not None else None
        scratch = _scratch(dev, lib.v1t_ensemble_scratch_bytes(dy.numel(), len(members)))
        with torch.cuda.device(dev):
            _lib.check(lib.v1t_ensemble_backward(C.byref(tab), weight.data_ptr(), _ptr(bias), dy.data_ptr(), dy.numel(),
                                                 dw.data_ptr(), _ptr(db), scratch.data_ptr(), _stream_ptr(dev)),
                       "ensemble_backward")
        return (dw, db) + (None,) * len(members)


def ensemble_combine(members: Sequence[torch.Tensor], weight: Optional[torch.Tensor] = None,
                     bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """members: K pre-activation responses [B,N]; weight [1,K] / bias [1] of the nn.Linear(K,1) output module, or
    None for the mean (ensemble_mode 0).  Returns the activated ensemble response [B,N]."""
    return _EnsembleFunction.apply(weight, bias, *members)
