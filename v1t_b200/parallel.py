"""Single-node data parallelism for the hot path: one process per GPU, torch.distributed (NCCL over NVLink).

The reference has no distributed code; its training loop accumulates gradients over all mice
before one optimizer step (train.py:97-111), so every layout below is mathematically identical to it:

* ``mouse2d`` (bench default) the step's samples, listed mouse by mouse, are cut into ``world`` equal contiguous
             ranges (SURVEY.md §8e "2-D layout"): a rank works on one or two mice, no rank idles at 7 mice on 8
             GPUs, and a readout lives only on the ranks that hold rows of its mouse.  The shared core's gradients
             are all-reduced over all ranks (in place in the flat gradient sink, one call); a readout's gradients are
             reduced only inside its mouse's rank group (a singleton for most mice: they never leave their rank).
* ``batch``  every rank sweeps all mice on its own shard of each mouse's batch; ALL gradients are sum-all-reduced
             (core 9.9 MB + readouts ~5.1 MB per mouse, fp32; readouts replicated).
* ``mouse``  rank r owns mice r, r+W, ... with the per-mouse batch unchanged (strong scaling of one sweep; ranks
             idle when W > mice).

The criterion must be given the GLOBAL batch size (losses.py:114-119 scales by sqrt(ds_size / batch_size)).
In the mouse-sharded layouts a rank only ever updates the readouts of its own mice; the copies on other ranks go
stale by design ("readouts sharded by mouse") and a checkpoint gathers each readout from the first rank of its group.
"""
from __future__ import annotations

import os
import typing as t
from dataclasses import dataclass, field

import torch
import torch.distributed as dist


def init_from_env(backend: t.Optional[str] = None):
    """Initialise from torchrun's RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def seed_rank_streams(seed: int, rank: int):
    """Give every rank its own dropout-seed (CPU generator, modules.ViTCore.forward) and position-noise (device
    generator) streams.  Call AFTER the model was constructed under the common seed: with identical streams all
    ranks would apply bit-identical dropout masks and readout noise to their different shards."""
    torch.manual_seed((int(seed) * 1000003 + 7919 * (rank + 1)) & 0x7FFFFFFFFFFFFFFF)


def mice_of_rank(mouse_ids: t.Sequence[str], rank: int, world: int, mode: str) -> t.List[str]:
    return list(make_plan(mouse_ids, rank, world, mode, 1).my_slices)


@dataclass
class Plan:
    """Which rows of which mouse's global batch every rank processes in one step."""

    mode: str
    rank: int
    world: int
    mice: t.List[str]
    global_batch: t.Dict[str, int]                                  # what the criterion is scaled with
    slices: t.List[t.Dict[str, t.Tuple[int, int]]] = field(default_factory=list)  # per rank: mouse -> [lo, hi)

    @property
    def my_slices(self) -> t.Dict[str, t.Tuple[int, int]]:
        return self.slices[self.rank]

    @property
    def scaling(self) -> str:
        return "strong" if self.mode == "mouse" else "weak"

    def group_of(self, mouse: str) -> t.Tuple[int, ...]:
        """Ranks that hold rows of ``mouse`` (its readout's gradient exchange group)."""
        return tuple(r for r in range(self.world) if mouse in self.slices[r])


def make_plan(mouse_ids: t.Sequence[str], rank: int, world: int, mode: str, batch: int) -> Plan:
    """``batch`` = rows per mouse per GPU of the single-GPU step (weak scaling: the global batch of every mouse is
    batch x world in the ``batch`` and ``mouse2d`` layouts)."""
    mice = list(mouse_ids)
    if world == 1:
        mode_eff = mode if mode in ("batch", "mouse", "mouse2d") else None
        if mode_eff is None:
            raise ValueError(f"unknown dp mode {mode}")
        return Plan(mode, rank, world, mice, {m: batch for m in mice}, [{m: (0, batch) for m in mice}])
    if mode == "batch":
        return Plan(mode, rank, world, mice, {m: batch * world for m in mice},
                    [{m: (r * batch, (r + 1) * batch) for m in mice} for r in range(world)])
    if mode == "mouse":
        return Plan(mode, rank, world, mice, {m: batch for m in mice},
                    [{m: (0, batch) for i, m in enumerate(mice) if i % world == r} for r in range(world)])
    if mode == "mouse2d":
        gb = batch * world                      # rows of one mouse in the step
        per = len(mice) * batch                 # rows per rank: the concatenated list has len(mice) * gb rows
        slices = []
        for r in range(world):
            lo, hi = r * per, (r + 1) * per
            mine = {}
            for i, m in enumerate(mice):
                a, b = max(lo, i * gb), min(hi, (i + 1) * gb)
                if b > a:
                    mine[m] = (a - i * gb, b - i * gb)
            slices.append(mine)
        return Plan(mode, rank, world, mice, {m: gb for m in mice}, slices)
    raise ValueError(f"unknown dp mode {mode}")


class FlatGrads:
    """``.grad`` of a set of parameters as views into ONE flat fp32 buffer, so that the set is exchanged with a
    single in-place all-reduce (no per-parameter copy in / copy out).  autograd accumulates into an existing
    ``.grad`` in place, so the views survive ``backward()``; ``arm()`` must be called again after
    ``zero_grad(set_to_none=True)``."""

    def __init__(self, params: t.Sequence[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        self.offsets, off = {}, 0
        for p in self.params:
            self.offsets[id(p)] = (off, p.numel())
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.flat: t.Optional[torch.Tensor] = None

    def view_of(self, p):
        off, n = self.offsets[id(p)]
        return self.flat[off:off + n].view(p.shape)

    def arm(self):
        if not self.params:
            return
        dev = self.params[0].device
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
            fresh = True
        elif all(p.grad is None for p in self.params):
            self.flat.zero_()
            fresh = True
        else:
            fresh = False
        for p in self.params:
            v = self.view_of(p)
            if p.grad is None:
                if not fresh:
                    v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v

    def is_live(self) -> bool:
        return self.flat is not None and all(
            p.grad is not None and p.grad.data_ptr() == self.view_of(p).data_ptr() for p in self.params)


class GradSync:
    """Gradient exchange of one step.

    ``GradSync(model, plan)``: the shared core over all ranks (in place in its flat gradient sink when armed), every
    mouse's private parameters (readout, core shifter, image shifter) inside the mouse's rank group only, through a
    FlatGrads buffer (one in-place call per shared mouse, none for a mouse that lives on one rank).
    ``GradSync(params)``: plain flat-bucket sum-all-reduce of an arbitrary parameter list over all ranks."""

    def __init__(self, model_or_params, plan: t.Optional[Plan] = None, bucket_mb: float = 64.0):
        self.plan = plan
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self._flat: t.Dict[int, torch.Tensor] = {}
        self.mouse_flat: t.Dict[str, FlatGrads] = {}
        self.mouse_group: t.Dict[str, t.Any] = {}
        self.side_stream = None
        if plan is None or not isinstance(model_or_params, torch.nn.Module):
            self.params = [p for p in model_or_params if p.requires_grad]
            return
        model = model_or_params
        private: t.Dict[int, str] = {}
        for m in plan.mice:
            ps = self._mouse_params(model, m)
            for p in ps:
                private[id(p)] = m
            grp = plan.group_of(m)
            if len(grp) > 1 and plan.world > 1 and dist.is_initialized():
                # every rank creates every multi-rank group, in the same order (torch.distributed requirement)
                pg = dist.group.WORLD if len(grp) == plan.world else dist.new_group(ranks=list(grp))
                if plan.rank in grp:
                    self.mouse_group[m] = pg
                    self.mouse_flat[m] = FlatGrads(ps)
        # everything that is not private to a mouse (the core; anything else a caller added) goes over all ranks
        self.params = [p for p in model.parameters() if p.requires_grad and id(p) not in private]

    @staticmethod
    def _mouse_params(model, mouse: str) -> t.List[torch.nn.Parameter]:
        ps = list(model.readouts[mouse].parameters())
        shifter = getattr(model, "core_shifter", None)
        if shifter is not None and mouse in shifter:
            ps += list(shifter[mouse].parameters())
        cropper = getattr(model, "image_cropper", None)
        ish = getattr(cropper, "image_shifter", None) if cropper is not None else None
        if ish is not None and mouse in ish:
            ps += list(ish[mouse].parameters())
        return ps

    def arm(self):
        """Point the shared mice's gradients at their flat exchange buffers (call before the step's backward passes)."""
        for fg in self.mouse_flat.values():
            fg.arm()

    def buckets(self):
        cur, n = [], 0
        for p in self.params:
            if cur and n + p.numel() > self.bucket_elems:
                yield cur
                cur, n = [], 0
            cur.append(p)
            n += p.numel()
        if cur:
            yield cur

    @staticmethod
    def _sink_is_live(sink) -> bool:
        """True when every parameter of the sink has its .grad viewing the sink's flat buffer (GradSink.arm)."""
        if sink is None or sink.flat is None or not sink.params:
            return False
        return all(p.grad is not None and p.grad.data_ptr() == sink.view_of(p, sink.flat).data_ptr()
                   for p in sink.params)

    @torch.no_grad()
    def all_reduce(self, sinks: t.Sequence = ()):
        """``sinks``: GradSinks (functional.GradSink) whose flat buffers already hold their parameters' gradients
        contiguously: those are reduced in place with one call each and their parameters skip the bucket copies.
        Every rank must pass the same sinks (the collectives are issued in the same order on all ranks)."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        skip = set()
        for sink in sinks:
            if self._sink_is_live(sink):
                dist.all_reduce(sink.flat, op=dist.ReduceOp.SUM)
                skip.update(id(p) for p in sink.params)
        # mouse-private parameters: inside the mouse's group, in plan order (identical on every member)
        if self.plan is not None:
            for m in self.plan.mice:
                fg = self.mouse_flat.get(m)
                if fg is None:
                    continue
                if not fg.is_live():
                    fg.arm()
                dist.all_reduce(fg.flat, op=dist.ReduceOp.SUM, group=self.mouse_group[m])
        for i, bucket in enumerate(self.buckets()):
            bucket = [p for p in bucket if id(p) not in skip]
            if not bucket:
                continue
            n = sum(p.numel() for p in bucket)
            flat = self._flat.get(i)
            dev = bucket[0].device
            if flat is None or flat.numel() != n or flat.device != dev:
                flat = torch.empty(n, dtype=torch.float32, device=dev)
                self._flat[i] = flat
            o = 0
            for p in bucket:
                k = p.numel()
                if p.grad is None:
                    flat[o:o + k].zero_()
                else:
                    flat[o:o + k].copy_(p.grad.reshape(-1))
                o += k
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            o = 0
            for p in bucket:
                k = p.numel()
                if p.grad is None:
                    p.grad = flat[o:o + k].reshape(p.shape).clone()
                else:
                    p.grad.copy_(flat[o:o + k].reshape(p.shape))
                o += k


def _can_fuse_core(model, batches, micro_batch: int) -> bool:
    """One core pass over all mice of the step is possible when the core's parameters do not depend on the mouse
    (behavior_mode 4 has per-mouse behaviour MLPs), the step is not micro-batched and the model has the repo's wiring."""
    core = getattr(model, "core", None)
    return (len(batches) > 1 and not (micro_batch and micro_batch > 0) and getattr(core, "behavior_mode", 4) != 4
            and all(hasattr(model, a) for a in ("image_cropper", "readouts", "elu1", "core_shifter")))


def _fused_core_pass(model, criterion, batches, global_batch):
    """The shared core runs ONCE on the concatenated batches of all mice (same samples, same math: the reference feeds
    the mice one after the other only because Model.forward takes one mouse_id, model.py:151-177); readout, ELU1 and
    the loss stay per mouse on row slices of the core's output view, and one backward through the summed loss gives the
    gradients train.py:84-111 accumulates step by step."""
    mice = list(batches)
    images, behaviors, pupils, sizes = [], [], [], []
    for m in mice:
        b = batches[m]
        im, _ = model.image_cropper(b["image"], mouse_id=m, behaviors=b["behavior"], pupil_centers=b["pupil_center"])
        images.append(im)
        behaviors.append(b["behavior"])
        pupils.append(b["pupil_center"])
        sizes.append(im.shape[0])
    fmap = model.core(torch.cat(images), mouse_id=mice[0], behaviors=torch.cat(behaviors),
                      pupil_centers=torch.cat(pupils))
    total = None
    for m, f in zip(mice, fmap.split(sizes)):  # row slices of the channel-last view (strides preserved)
        b = batches[m]
        shifts = model.core_shifter(b["pupil_center"], mouse_id=m) if model.core_shifter is not None else None
        y = model.elu1(model.readouts(f, mouse_id=m, shifts=shifts))
        loss = criterion(y_true=b["response"], y_pred=y, mouse_id=m, batch_size=global_batch[m])
        total = loss if total is None else total + loss
    total.backward()
    return total.detach()


def sweep(model, criterion, batches: t.Dict[str, t.Dict[str, torch.Tensor]], global_batch: t.Dict[str, int],
          sync: t.Optional[GradSync] = None, fused_accumulate: bool = False, micro_batch: int = 0,
          fuse_core: bool = False, fuse_rows: int = 128):
    """One optimizer step's worth of forward/backward: every mouse batch of this rank once (in micro-batches of
    ``micro_batch`` rows when > 0, like data.micro_batching / train.py:55), gradients accumulated (train.py:84-111
    without the optimizer), then the gradient exchange.  ``fuse_core``: run the shared core once over all mice of the
    step (see _fused_core_pass) when the model allows it.  Returns the summed loss (device scalar)."""
    total = None
    if fused_accumulate:  # one add per backward for all shared-core gradients instead of one per parameter
        model.core.fused_grad_accumulation(True)
    if sync is not None:
        sync.arm()
    if fuse_core and _can_fuse_core(model, batches, micro_batch):
        # consecutive mice are grouped while a core pass stays within ``fuse_rows`` samples (saved activations: ~150 MB
        # per sample of the default core); a group of one mouse is the ordinary per-mouse pass
        group, rows = {}, 0
        for mouse_id, b in list(batches.items()) + [(None, None)]:
            n = b["image"].shape[0] if b is not None else 0
            if group and (b is None or rows + n > fuse_rows):
                part = (_fused_core_pass(model, criterion, group, global_batch) if len(group) > 1
                        else sweep(model, criterion, group, global_batch, None, False, 0, False))
                total = part if total is None else total + part
                group, rows = {}, 0
            if b is not None:
                group[mouse_id] = b
                rows += n
        batches = {}
    for mouse_id, b in batches.items():
        rows = b["image"].shape[0]
        step = micro_batch if micro_batch and micro_batch > 0 else rows
        for lo in range(0, rows, step):
            mb = b if step >= rows else {k: (v[lo:lo + step] if v is not None else None) for k, v in b.items()}
            y, _, _ = model(inputs=mb["image"], mouse_id=mouse_id, behaviors=mb["behavior"],
                            pupil_centers=mb["pupil_center"])
            loss = criterion(y_true=mb["response"], y_pred=y, mouse_id=mouse_id, batch_size=global_batch[mouse_id])
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
    sinks = ()
    if fused_accumulate:
        model.core.fused_grad_accumulation(False)
        sinks = (model.core.grad_sink,)  # its flat buffer IS the core gradients: reduced in place, no bucket copies
    if sync is not None:
        sync.all_reduce(sinks=sinks)
    return total
