"""Single-node data parallelism for the hot path: one process per GPU, torch.distributed (NCCL over NVLink).

The reference has no distributed code; its training loop accumulates gradients over all mice
before one optimizer step (train.py:97-111), so both layouts below are mathematically identical to it:

* ``batch``  every rank sweeps all mice on its own shard of each mouse's batch; ALL gradients are sum-all-reduced
             (core 9.9 MB + readouts ~5.1 MB per mouse, fp32).  Weak-scaling layout used by bench.py.
* ``mouse``  rank r owns mice r, r+W, ... (readouts "sharded by mouse"); only the shared core's gradients are
             all-reduced, readout / shifter gradients never leave their rank.

The criterion must be given the GLOBAL batch size (losses.py:114-119 scales by sqrt(ds_size / batch_size)).
"""
from __future__ import annotations

import os
import typing as t

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


def mice_of_rank(mouse_ids: t.Sequence[str], rank: int, world: int, mode: str) -> t.List[str]:
    if mode == "batch" or world == 1:
        return list(mouse_ids)
    if mode == "mouse":
        return [m for i, m in enumerate(mouse_ids) if i % world == rank]
    raise ValueError(f"unknown dp mode {mode}")


class GradSync:
    """Flat-bucket sum-all-reduce of gradients (one NCCL call per bucket; NVSwitch makes one big bucket best)."""

    def __init__(self, params: t.Iterable[torch.nn.Parameter], bucket_mb: float = 64.0):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self._flat: t.Dict[int, torch.Tensor] = {}

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


def sweep(model, criterion, batches: t.Dict[str, t.Dict[str, torch.Tensor]], global_batch: t.Dict[str, int],
          sync: t.Optional[GradSync] = None, fused_accumulate: bool = False):
    """One optimizer step's worth of forward/backward: every mouse batch once, gradients accumulated
    (train.py:84-111 without the optimizer), then the gradient exchange.  Returns the summed loss (device scalar)."""
    total = None
    if fused_accumulate:  # one add per backward for all shared-core gradients instead of one per parameter
        model.core.fused_grad_accumulation(True)
    for mouse_id, b in batches.items():
        y, _, _ = model(inputs=b["image"], mouse_id=mouse_id, behaviors=b["behavior"],
                        pupil_centers=b["pupil_center"])
        loss = criterion(y_true=b["response"], y_pred=y, mouse_id=mouse_id, batch_size=global_batch[mouse_id])
        loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    sinks = ()
    if fused_accumulate:
        model.core.fused_grad_accumulation(False)
        sinks = (model.core.grad_sink,)  # its flat buffer IS the core gradients: reduced in place, no bucket copies
    if sync is not None:
        sync.all_reduce(sinks=sinks)
    return total
