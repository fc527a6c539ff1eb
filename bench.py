#!/usr/bin/env python
"""Benchmark of the V1T hot path (ViT core + per-mouse Gaussian2d readout + Poisson loss, forward + backward).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    torchrun ... bench.py --gpus N ...                        # one rank per GPU, NCCL
    python bench.py --impl reference ...                      # the reference's own CPU path on the host cores
    python bench.py --config franke|scaled|ensemble ...       # the other BASELINE.json configs (not the driver line)

Default workload = BASELINE.json configs[1]: Sensorium+ 7-mouse shared V1T core (default vit args, 1x36x64 frames,
1654 tokens), per-mouse gaussian2d readouts (~8k neurons each), behavior_mode 3, train mode (dropout + readout
position sampling), batch 16 per mouse per GPU.  One STEP = one optimizer step's worth of forward/backward: every
mouse batch once with gradient accumulation (train.py:84-111) + (N>1) the gradient all-reduce.  The optimizer
update, the L1 regulariser and the image cropper are outside the metric.  Synthetic data, seeded.

Baselines printed beside it (N=1): ``cpu_baseline`` = the UNMODIFIED reference (oracle/_ref, staged by
oracle/make_ref.py; the oracle port when that is absent) on the host cores, train mode and eval-with-grads;
``gpu_eager_baseline`` = the same reference modules on this B200 in PyTorch eager with TF32 as the reference enables
it (utils/utils.py:38-43) — the "bar to beat" of SURVEY.md §2.2.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train samples/sec (core+readout fwd/bwd)"
SEED = 1234

# BASELINE.json configs[1..4]; "core" overrides the default vit args (train.py:543-573)
CONFIGS = {
    "sensorium": dict(index=1, mice=7, batch=16, neurons=8000, in_shape=(1, 36, 64), ds_name="sensorium", core={},
                      impl="bf16x3", micro=0,
                      label="Sensorium+ 7-mouse shared V1T core (default vit args, 1x36x64, 1654 tokens) + per-mouse "
                            "gaussian2d readouts (~8k neurons), behavior_mode 3, train mode (dropout + position sampling)"),
    "franke": dict(index=2, mice=10, batch=64, neurons=1100, in_shape=(2, 36, 64), ds_name="franke2022", core={},
                   impl="bf16x3", micro=0,
                   label="Franke2022 10-mouse colour (2-channel 36x64) stimuli, shared default V1T core, per-mouse "
                         "gaussian2d readouts (~1.1k neurons), behavior_mode 3, train mode, batch 64 per mouse per GPU"),
    "scaled": dict(index=3, mice=7, batch=256, neurons=8000, in_shape=(1, 36, 64), ds_name="sensorium",
                   core=dict(emb_dim=512, num_blocks=8, num_heads=8), impl="bf16", micro=16,
                   label="scaled V1T core (emb 512 = head dim 512, 8 blocks, 8 heads, mlp 488) on 7 mice, batch 256 per "
                         "mouse in micro-batches (data.micro_batching, train.py:55), bf16 tensor-core operands"),
    "ensemble": dict(index=4, mice=1, batch=16, neurons=8000, in_shape=(1, 36, 64), ds_name="sensorium", core={},
                     impl="bf16x3", micro=0, members=5,
                     label="5-model ensemble inference + attention-rollout extraction (Recorder on every block of one "
                           "member, emit-P, rollout), Sensorium+ test-shape batch of 16, eval mode"),
}


def neuron_counts(n_mice: int, base: int):
    rng = np.random.default_rng(SEED)
    return {chr(ord("A") + i): int(base * (1 + 0.1 * (2 * rng.random() - 1))) // 8 * 8 for i in range(n_mice)}


def make_args(neurons, device, impl="bf16x3", in_shape=(1, 36, 64), ds_name="sensorium", **over):
    a = dict(input_shape=tuple(in_shape), output_shapes={k: (n,) for k, n in neurons.items()}, device=device,
             core="vit", readout="gaussian2d", behavior_mode=3, shift_mode=2, center_crop=1.0, resize_image=0,
             ds_name=ds_name, patch_mode=0, patch_size=8, patch_stride=1, emb_dim=155, num_blocks=4, num_heads=4,
             mlp_dim=488, p_dropout=0.0229, t_dropout=0.2544, drop_path=0.0, use_lsa=False, disable_bias=False,
             grad_checkpointing=0, core_reg_scale=0.5379, readout_reg_scale=0.0076, disable_grid_predictor=False,
             grid_predictor_dim=2, bias_mode=0, shifter_reg_scale=0.0, cropper_reg_scale=0.0, criterion="poisson",
             ds_scale=1, verbose=0, b200_impl=impl, gray_scale=False)
    a.update(over)
    return SimpleNamespace(**a)


class _DS:
    def __init__(self, n, rng):
        self.coordinates = rng.standard_normal((n, 3)).astype(np.float32)
        self.response_stats = {"mean": np.ones(n, np.float32), "std": np.ones(n, np.float32)}

    def __len__(self):
        return 4500


def make_ds(neurons):
    rng = np.random.default_rng(SEED + 1)
    return {k: SimpleNamespace(dataset=_DS(n, rng)) for k, n in neurons.items()}


def host_batches(neurons, batch, rank, pin, in_shape=(1, 36, 64)):
    """Synthetic Sensorium-shaped inputs: images~N(0,1), behaviours/pupil~U(0,1), responses~U(0,2)."""
    g = torch.Generator().manual_seed(SEED + 17 * (rank + 1))
    out = {}
    for m, n in neurons.items():
        b = {"image": torch.randn((batch,) + tuple(in_shape), generator=g), "behavior": torch.rand((batch, 3), generator=g),
             "pupil_center": torch.rand((batch, 2), generator=g), "response": torch.rand((batch, n), generator=g) * 2}
        out[m] = {k: (v.pin_memory() if pin else v) for k, v in b.items()}
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def core_dims(cfg, margs):
    c, h, w = cfg["in_shape"]
    gh = (h - margs.patch_size) // margs.patch_stride + 1
    gw = (w - margs.patch_size) // margs.patch_stride + 1
    return dict(T=gh * gw + 1, E=margs.emb_dim, H=margs.num_heads, M=margs.mlp_dim, blocks=margs.num_blocks, C=c,
                pd=c * margs.patch_size ** 2)


# per-sample algorithmic forward FLOPs; head dim == emb dim (vit.py:218)
def attn_flops_fwd(T=1654, E=155, H=4, blocks=4, **_):
    return blocks * 2 * (2 * H * T * T * E)


def core_flops_fwd(T=1654, E=155, H=4, M=488, blocks=4, pd=64, **_):
    L = T - 1
    return 2 * L * pd * E + blocks * (2 * T * E * 3 * H * E + 4 * H * T * T * E + 2 * T * H * E * E + 4 * T * E * M)


# ------------------------------------------------------------------------------------------------------
# reference legs (oracle/_ref = the unmodified reference; oracle port as the fallback) — never the product path
# ------------------------------------------------------------------------------------------------------
def build_reference(cfg, neurons, device, grad_checkpointing=0):
    """(model, criterion) of the UNMODIFIED reference with the bench's trained-like weights, or None if absent."""
    from oracle import ref_harness as rh

    if not rh.reference_available():
        return None
    over = dict(cfg["core"])
    args = rh.make_args(neurons, in_shape=cfg["in_shape"], device=torch.device(device), ds_name=cfg["ds_name"],
                        grad_checkpointing=grad_checkpointing, **over)
    ds = rh.make_fake_ds(neurons, ds_size=4500, seed=SEED + 1)
    model, crit = rh.build_reference_model(args, ds, seed=SEED, trained_like=True)
    return model.to(device), crit.to(device)


def reference_step(model, crit, mouse, b, global_batch):
    """fwd + Poisson loss + backward of one mouse batch through the reference's own Model (model.py:151-177)."""
    y, _, _ = model(inputs=b["image"], mouse_id=mouse, behaviors=b["behavior"], pupil_centers=b["pupil_center"])
    loss = crit(y_true=b["response"], y_pred=y, mouse_id=mouse, batch_size=global_batch)
    loss.backward()
    return loss


def cpu_reference_sample(cfg, batch, steps, warmup, cores, train_mode=True):
    """The reference's CPU path on a bounded sample of the workload: each step = ONE mouse batch of ``batch`` samples
    (the mice of the config taken in turn), fwd + loss + bwd, fp32, all host threads."""
    torch.set_num_threads(cores)
    neurons = neuron_counts(cfg["mice"], cfg["neurons"])
    built = build_reference(cfg, neurons, "cpu")
    mice = list(neurons)
    host = host_batches(neurons, batch, 0, pin=False, in_shape=cfg["in_shape"])
    times = []
    if built is not None:
        model, crit = built
        model.train(train_mode)
        kind = "reference"
        for i in range(warmup + steps):
            m = mice[i % len(mice)]
            model.zero_grad(set_to_none=True)
            t0 = time.perf_counter()
            reference_step(model, crit, m, host[m], batch)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    else:  # oracle port (same ATen ops), default core only
        from oracle import torch_port as TP
        from oracle.v1t_oracle import CoreConfig
        import v1t_b200

        kind = "port"
        one = {mice[0]: neurons[mice[0]]}
        margs = make_args(one, torch.device("cpu"))
        torch.manual_seed(SEED)
        sd = {k: v.detach().clone() for k, v in v1t_b200.Model(margs, ds=make_ds(one)).state_dict().items()}
        for k, v in sd.items():
            if v.is_floating_point() and v.dim() > 0 and not k.endswith(("source_grid", "grid")):
                v.requires_grad_(True)
        b = host[mice[0]]
        for i in range(warmup + steps):
            for v in sd.values():
                v.grad = None
            t0 = time.perf_counter()
            TP.step(sd, CoreConfig(), mice[0], b["image"], b["behavior"], b["pupil_center"], b["response"],
                    ds_size=4500, p_drop=margs.p_dropout, t_drop=margs.t_dropout, training=train_mode)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    return {"value": batch / (ms / 1e3), "unit": "samples/s", "cores": cores, "kind": kind, "ms_per_step": ms,
            "sample": f"per step ONE mouse batch of {batch} samples (the config's {len(mice)} mice in turn, "
                      f"N~{cfg['neurons']} neurons each), {'train' if train_mode else 'eval-with-grads'} mode fwd+loss+bwd, "
                      f"fp32 torch CPU, {steps} timed steps after {warmup} warm-up"}


def run_reference(args, cfg):
    """`--impl reference`: the reference's own CPU implementation of the path, all host threads, a bounded sample of
    the workload per step (see cpu_reference_sample).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    res = cpu_reference_sample(cfg, args.ref_batch, max(args.steps, 1), max(args.warmup, 0), cores, train_mode=True)
    ev = cpu_reference_sample(cfg, args.ref_batch, min(max(args.steps, 1), 3), 1, cores, train_mode=False)
    config = workload_config(args, cfg, 1)
    config["reference_sample"] = res["sample"]
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "cpu_baseline_eval": {k: ev[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(cfg, dev, steps=3, warmup=2, micro=0, max_mice=None):
    """The reference's own modules on this GPU in PyTorch eager, TF32 as utils/utils.py:42-43 enables it, the same
    sweep as the timed workload (every mouse batch once, fwd + loss + bwd, train mode), both with the reference's
    CUDA default (gradient checkpointing of the attention, vit.py:377-380) and without."""
    neurons = neuron_counts(cfg["mice"], cfg["neurons"])
    if max_mice:  # bounded sample for the big configs: the first mice only (the per-sample rate is what is reported)
        neurons = dict(list(neurons.items())[:max_mice])
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    out = {}
    try:
        for name, ckpt in (("checkpointing", 1), ("no_checkpointing", 0)):
            built = build_reference(cfg, neurons, dev, grad_checkpointing=ckpt)
            if built is None:
                return {"unavailable": "oracle/_ref absent (run oracle/make_ref.py where /root/reference exists)"}
            model, crit = built
            model.train(True)
            host = host_batches(neurons, cfg["batch"], 0, pin=False, in_shape=cfg["in_shape"])
            resident = {m: {k: v.to(dev) for k, v in b.items()} for m, b in host.items()}

            def sweep():
                model.zero_grad(set_to_none=True)
                for m, b in resident.items():
                    rows = b["image"].shape[0]
                    step = micro or rows
                    for lo in range(0, rows, step):  # data.micro_batching (train.py:55)
                        reference_step(model, crit, m, {k: v[lo:lo + step] for k, v in b.items()}, cfg["batch"])

            for _ in range(warmup):
                sweep()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                sweep()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "value": cfg["batch"] * len(neurons) / (ms / 1e3)}
            del model, crit, resident, built
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    best = max(out.values(), key=lambda r: r["value"])
    return {"value": best["value"], "unit": "samples/s", "ms_per_step": best["ms_per_step"], "steps": steps,
            "warmup": warmup, "kind": "reference modules (oracle/_ref), PyTorch eager on this GPU, TF32 matmul",
            "mice": len(neurons), "micro_batch": micro or cfg["batch"], "variants": out,
            "note": "value = the faster variant"}


def workload_config(args, cfg, world):
    return {"workload": cfg["label"], "baseline_config_index": cfg["index"], "mice": args.mice,
            "batch_per_mouse_per_gpu": args.batch, "global_batch": args.mice * args.batch * world,
            "neurons_per_mouse": args.neurons, "parallelism": f"dp{world} ({args.dp_mode})", "impl": args.b200_impl,
            "micro_batch": args.micro or args.batch,
            "core_passes": "one per mouse" if (getattr(args, "no_fuse_core", False) or args.micro) else
                           "one over all mice of the step (shared core, per-mouse readouts on row slices)",
            "l2": "per-step working set (saved activations ~1.4 GB per mouse batch) >> 126 MB L2; no explicit flush"}


def run_b200(args, cfg):
    import v1t_b200
    from v1t_b200 import _lib, parallel

    rank, local, world = parallel.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU path); use --impl reference for the CPU baseline"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    neurons = neuron_counts(args.mice, args.neurons)
    margs = make_args(neurons, dev, impl=args.b200_impl, in_shape=cfg["in_shape"], ds_name=cfg["ds_name"], **cfg["core"])
    torch.manual_seed(SEED)
    model = v1t_b200.Model(margs, ds=make_ds(neurons)).to(dev)
    crit = v1t_b200.get_criterion(margs, ds=make_ds(neurons))
    with torch.no_grad():  # "trained-like" readout weights so responses are not all ~1
        g = torch.Generator(device=dev).manual_seed(SEED + 2)
        for r in model.readouts.values():
            r.features.add_(torch.randn(r.features.shape, device=dev, generator=g) * 0.05)
            r.bias.add_(torch.randn(r.bias.shape, device=dev, generator=g) * 0.3)
    model.train(True)
    parallel.seed_rank_streams(SEED, rank)  # dropout / position-noise streams differ across ranks
    plan = parallel.make_plan(list(neurons), rank, world, args.dp_mode, args.batch)
    sync = parallel.GradSync(model, plan) if world > 1 else None
    host = {}
    for m, (lo, hi) in plan.my_slices.items():  # this rank's rows of each mouse's GLOBAL batch (same on every rank)
        full = host_batches({m: neurons[m]}, plan.global_batch[m], list(neurons).index(m), pin=False,
                            in_shape=cfg["in_shape"])[m]
        host[m] = {k: v[lo:hi].contiguous().pin_memory() for k, v in full.items()}
    resident = {m: {k: v.to(dev) for k, v in b.items()} for m, b in host.items()}
    samples_per_step_rank = sum(hi - lo for lo, hi in plan.my_slices.values())
    h2d = sum(v.numel() * 4 for b in host.values() for v in b.values())
    micro = args.micro
    fuse = not args.no_fuse_core

    def step_resident():
        model.zero_grad(set_to_none=True)
        return parallel.sweep(model, crit, resident, plan.global_batch, sync, fused_accumulate=True, micro_batch=micro,
                              fuse_core=fuse)

    def step_e2e():
        model.zero_grad(set_to_none=True)
        dev_b = {m: {k: v.to(dev, non_blocking=True) for k, v in b.items()} for m, b in host.items()}
        loss = parallel.sweep(model, crit, dev_b, plan.global_batch, sync, fused_accumulate=True, micro_batch=micro,
                              fuse_core=fuse)
        return float(loss.item()) if loss is not None else 0.0  # device->host read of the step's loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    launches0 = lib.v1t_launch_count()
    with ClockSampler(local) as clocks:
        ms_total = timed(step_resident, args.steps)
    launches = int(lib.v1t_launch_count() - launches0)
    total_samples = torch.tensor([samples_per_step_rank], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(total_samples)
    samples_per_step = float(total_samples.item())
    ms_per_step = ms_total / args.steps
    value = samples_per_step / (ms_per_step / 1e3)

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e = {"value": samples_per_step / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e}

    # per-phase device time (CUDA events on the launching stream) of one more timed window
    lib.v1t_prof_enable(1)
    lib.v1t_prof_reset()
    psteps = min(args.steps, 3)
    for _ in range(psteps):
        step_resident()
    torch.cuda.synchronize(dev)
    phases = {}
    for i, name in enumerate(_lib.PHASES):
        tot, cnt = ctypes.c_float(0), ctypes.c_int(0)
        lib.v1t_prof_read(i, ctypes.byref(tot), ctypes.byref(cnt))
        phases[name] = {"ms_per_step": tot.value / psteps, "scopes_per_step": cnt.value / psteps}
    lib.v1t_prof_enable(0)
    lib.v1t_prof_reset()
    phase_sum = sum(p["ms_per_step"] for k, p in phases.items() if k not in _lib.NESTED_PHASES)

    peaks = load_peaks()
    D = core_dims(cfg, margs)
    n_local = samples_per_step_rank
    # Dominant KERNEL: the fused attention backward (one launch per block per micro-batch).  Algorithmic FLOPs per
    # launch = 2 x the forward's 4*B*H*T^2*E (no recompute credit); duration = CUDA events recorded by
    # the library on the launching stream right around the launch, averaged over the timed window.
    def kernel_roofline(name, phase, flops_per_sample_block):
        p = phases[phase]
        if p["scopes_per_step"] <= 0 or p["ms_per_step"] <= 0:
            return None
        us = 1e3 * p["ms_per_step"] / p["scopes_per_step"]
        b_launch = n_local / (p["scopes_per_step"] / D["blocks"])  # samples per launch
        flops = flops_per_sample_block * b_launch
        ach = flops / (us * 1e-6) / 1e12
        return {"kernel": name, "launches_per_step": p["scopes_per_step"], "us_per_launch": us,
                "flops_per_launch": flops, "samples_per_launch": b_launch, "achieved": ach,
                "frac": ach / peaks["bf16_tflops_sustained"], "share_of_step": p["ms_per_step"] / max(phase_sum, 1e-9)}

    per_block_fwd = 4 * D["H"] * D["T"] ** 2 * D["E"]
    dq_gemm = os.environ.get("V1T_ATTN_DQ", "gemm")[:1] != "p" and os.environ.get("V1T_ATTN_BWD", "pair")[:1] != "t"
    dq_kernel = "tc_gemm_kernel<8," if dq_gemm else "attn_bwd2_kernel"
    kb = kernel_roofline("attention backward = attn_bwd_pair_kernel (dV + dK + dS' planes, persistent 2-CTA clusters) + " +
                         ("tc_gemm_kernel<8,0,0> (dQ = dS K, batched plane GEMM)" if dq_gemm else "attn_bwd2_kernel (dQ pass)"),
                         "attn_bwd_kernel", 2 * per_block_fwd)
    if kb:
        for key, ph in (("pair_us_per_launch", "attn_bwd_pair"), ("dq_us_per_launch", "attn_bwd_dq")):
            p = phases.get(ph)
            if p and p["scopes_per_step"] > 0:
                kb[key] = 1e3 * p["ms_per_step"] / p["scopes_per_step"]
    kf = kernel_roofline("attn_fwd2_kernel", "attn_fwd_kernel", per_block_fwd)
    fused_path = kb is not None
    if kb is None:  # head dim > 160: no fused kernel; the scope holds the batched plane GEMMs + softmax-backward rows kernel
        kb = kernel_roofline("materialised attention backward (tc_gemm_kernel x6 over operand planes + "
                             "softmax_bwd_rows_planes_kernel + bh_planes_kernel x4, per chunk of samples)", "attn_bwd",
                             2 * per_block_fwd)
        kf = kernel_roofline("materialised attention forward (tc_gemm_kernel x2 + softmax_rows_planes_kernel + "
                             "bh_planes_kernel x3)", "attn_fwd", per_block_fwd)
    dom = kb or {"kernel": "attention backward", "achieved": 0.0, "frac": 0.0}
    roofline = {"bound": "tensor", "kernel": dom["kernel"], "achieved": dom["achieved"],
                "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": dom["frac"],
                "traffic": ((kernel_traffic("attn_bwd_pair_kernel") or 0) + (kernel_traffic(dq_kernel) or 0) or None)
                if fused_path else None,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the two launches (ncu launch list under profiles/; "
                                "captured at the batch that list names)",
                "peak_source": peaks["source"] + " (sustained cuBLAS bf16)", "detail": kb, "attn_fwd2_kernel": kf,
                "flops_note": "algorithmic FLOPs: fwd 4*B*H*T^2*E per launch, bwd 2x that; the bf16x3 mode executes 3 "
                              "MMAs per product and the backward recomputes S, so the tensor pipe does ~8x this"}
    ro_ms = phases["readout_fwd"]["ms_per_step"] + phases["readout_bwd"]["ms_per_step"]
    L, E = D["T"] - 1, D["E"]
    ro_bytes = 0
    for m, (lo, hi) in plan.my_slices.items():
        n, B = neurons[m], hi - lo
        fwd = B * L * E * 4 + E * n * 4 + 7 * n * 4 + 2 * B * n * 4 + 2 * B * n * 4
        ro_bytes += fwd + (fwd + B * L * E * 4 + E * n * 4)  # forward + backward
    readout = {"bound": "hbm", "achieved": ro_bytes / (ro_ms / 1e3) / 1e9 if ro_ms > 0 else 0.0,
               "peak": peaks["hbm_gbs"], "unit": "GB/s", "batch": args.batch}
    readout["frac"] = readout["achieved"] / readout["peak"]
    total_flops = 3 * core_flops_fwd(**D) * n_local

    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": plan.scaling, "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (fp32 accumulate)", "bf16": "bf16"}.get(args.b200_impl, args.b200_impl),
            "data": "synthetic", "config": workload_config(args, cfg, world), "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks.summary(), "roofline": roofline, "roofline_readout": readout, "phases": phases,
            "model_tflops": total_flops / (ms_per_step / 1e3) / 1e12}
    if world == 1 and not args.no_extras:
        line["extras"] = extras_rooflines(model, neurons, dev, peaks, cfg)
    del resident
    if world == 1 and not args.no_eager_baseline:
        from v1t_b200 import functional as VF

        model.zero_grad(set_to_none=True)
        VF.release_scratch()
        torch.cuda.empty_cache()
        try:
            big = bool(args.micro)
            eg = gpu_eager_baseline(cfg, dev, steps=1 if big else 3, warmup=1 if big else 2, micro=args.micro,
                                    max_mice=1 if big else None)
        except Exception as e:  # keep the headline line; report the failure
            eg = {"error": repr(e)[:300]}
        line["gpu_eager_baseline"] = eg
        if eg.get("value"):
            line["vs_gpu_eager"] = value / eg["value"]
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cb = cpu_reference_sample(cfg, args.ref_batch, 3, 1, cores, train_mode=True)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            ce = cpu_reference_sample(cfg, args.ref_batch, 3, 1, cores, train_mode=False)
            line["cpu_baseline_eval"] = {k: ce[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def extras_rooflines(model, neurons, dev, peaks, cfg):
    """HBM-bound kernels timed alone with CUDA events after warm-up, against the measured copy bandwidth: the fused
    L1 + AdamW pass over every parameter (28 B per element), the attention rollout of a recorded stack (each block's
    [H,T,T] read once) and the readout + ELU1 + Poisson kernels at batch 256 (SURVEY.md §8d byte formula).
    Not part of `value` (the metric excludes the optimizer)."""
    from v1t_b200 import _lib, functional as VF
    from v1t_b200.optim import FusedAdamWL1, l1_coefficients

    out = {}

    def timed_ms(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    try:
        opt = FusedAdamWL1(model.get_parameters(core_lr=1e-3), lr=1e-3, betas=(0.9, 0.9999), eps=1e-8, weight_decay=0,
                           l1=l1_coefficients(model, list(neurons)))
        numel = 0
        for p in model.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            numel += p.numel()
        ms_api = timed_ms(lambda: opt.step(), 20)  # through the torch.optim API (host-side table check included)
        table, prefix, n_tensors, n_chunks, scratch, sums = opt._table
        lib = _lib.load()
        stream = torch.cuda.current_stream(dev).cuda_stream

        def kernel_only():  # the same launch, back to back, so the device time is what is measured
            _lib.check(lib.v1t_adamw_l1_step(table.data_ptr(), prefix.data_ptr(), n_tensors, n_chunks, 0.9, 0.9999,
                                             1e-8, 0.5, 0.5, 1.0, 0, sums.data_ptr(), opt.n_l1_groups,
                                             scratch.data_ptr(), stream), "adamw_l1_step")

        ms = timed_ms(kernel_only, 50)
        gbs = 28.0 * numel / (ms / 1e3) / 1e9
        out["adamw_l1"] = {"ms_per_step": ms, "ms_per_step_api": ms_api, "params": numel, "bound": "hbm",
                           "traffic": kernel_traffic("adamw_l1_kernel"), "achieved": gbs,
                           "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                           "bytes_per_element": 28}
    except Exception as e:  # keep the headline line; report the failure
        out["adamw_l1"] = {"error": repr(e)}
    try:
        B, L, H, T = 16, 4, 4, 1654  # BASELINE configs[4]: Sensorium+ test-shape batch, default core
        g = torch.Generator(device=dev).manual_seed(SEED + 5)
        attn = torch.softmax(torch.randn((B, L, H, T, T), device=dev, generator=g) * 2.0, dim=-1)
        ms = timed_ms(lambda: VF.attention_rollouts(attn, (36, 64), (29, 57)), 10)
        nbytes = B * (L - 1) * H * T * T * 4
        gbs = nbytes / (ms / 1e3) / 1e9
        out["attention_rollout"] = {"ms": ms, "shape": [B, L, H, T, T], "bound": "hbm", "achieved": gbs,
                                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                                    "launches_per_call": L + 1, "traffic_per_step_launch_b4": kernel_traffic("rollout_step_kernel"),
                                    "note": "2.8 GB stack, 2.1 GB read per call (blocks 0..L-2 once, block L-1 row 0)"}
        del attn
    except Exception as e:
        out["attention_rollout"] = {"error": repr(e)}
    try:
        out["readout_b256"] = readout_roofline(model, neurons, dev, peaks, 256, timed_ms)
    except Exception as e:
        out["readout_b256"] = {"error": repr(e)}
    return out


def readout_roofline(model, neurons, dev, peaks, B, timed_ms):
    """Readout + ELU1 + Poisson forward and backward of ONE mouse at batch ``B`` on a synthetic core map, timed alone
    (CUDA events, back-to-back calls; the 262 MB map at B=256 exceeds L2).  Bytes = SURVEY.md §8d's formula."""
    import v1t_b200
    from v1t_b200 import _lib

    lib = _lib.load()
    m = next(iter(neurons))
    n = neurons[m]
    ro = model.readouts[m]
    E, gh, gw = model.core.output_shape
    L = gh * gw
    g = torch.Generator(device=dev).manual_seed(SEED + 9)
    tokens = torch.randn((B, L + 1, (E + 31) // 32 * 32), device=dev, generator=g)
    fmap = tokens[:, 1:, :E].unflatten(1, (gh, gw)).permute(0, 3, 1, 2).requires_grad_(True)
    y_true = torch.rand((B, n), device=dev, generator=g) * 2
    crit = v1t_b200.PoissonLoss(SimpleNamespace(ds_scale=1), ds={m: SimpleNamespace(dataset=[0] * 4500)}).to(dev)
    model.train(True)

    def fwd_bwd():
        z = ro(fmap, shifts=None)
        loss = crit(y_true=y_true, y_pred=model.elu1(z), mouse_id=m, batch_size=B)
        loss.backward()

    lib.v1t_prof_enable(1)
    lib.v1t_prof_reset()
    reps = 5
    ms_total = timed_ms(fwd_bwd, reps)
    torch.cuda.synchronize(dev)
    res = {}
    for name in ("readout_fwd", "readout_bwd"):
        tot, cnt = ctypes.c_float(0), ctypes.c_int(0)
        lib.v1t_prof_read(_lib.PHASES.index(name), ctypes.byref(tot), ctypes.byref(cnt))
        res[name] = tot.value / max(cnt.value, 1)
    lib.v1t_prof_enable(0)
    lib.v1t_prof_reset()
    fwd = B * L * E * 4 + E * n * 4 + 7 * n * 4 + 2 * B * n * 4 + 2 * B * n * 4
    bwd = fwd + B * L * E * 4 + E * n * 4
    out = {"batch": B, "neurons": n, "bound": "hbm", "peak": peaks["hbm_gbs"], "unit": "GB/s",
           "fwd_ms": res["readout_fwd"], "bwd_ms": res["readout_bwd"], "fwd_bwd_autograd_ms": ms_total,
           "fwd_bytes": fwd, "bwd_bytes": bwd,
           "fwd_achieved": fwd / (res["readout_fwd"] / 1e3) / 1e9 if res["readout_fwd"] > 0 else 0.0,
           "bwd_achieved": bwd / (res["readout_bwd"] / 1e3) / 1e9 if res["readout_bwd"] > 0 else 0.0}
    out["achieved"] = (fwd + bwd) / ((res["readout_fwd"] + res["readout_bwd"]) / 1e3) / 1e9
    out["frac"] = out["achieved"] / peaks["hbm_gbs"]
    fmap.grad = None
    return out


def kernel_traffic(kernel):
    """dram read + write bytes per launch of a kernel from the committed ncu captures (profiles/), newest round first."""
    for fname in ("r2_kernel_traffic.json", "r1_kernel_traffic.json", "r1_extras_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fname)) as fh:
                kernels = json.load(fh)["kernels"]
        except (OSError, ValueError, KeyError):
            continue
        for name, k in kernels.items():
            if kernel in name:
                return k["dram_read_bytes_per_launch"] + k["dram_write_bytes_per_launch"]
    return None


# ------------------------------------------------------------------------------------------------------
# configs[4]: 5-model ensemble inference + attention-rollout extraction
# ------------------------------------------------------------------------------------------------------
def run_ensemble(args, cfg):
    import v1t_b200
    from v1t_b200 import _lib, parallel
    from v1t_b200.rollout import Recorder, attention_rollouts

    rank, local, world = parallel.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    neurons = neuron_counts(args.mice, args.neurons)
    margs = make_args(neurons, dev, impl=args.b200_impl, in_shape=cfg["in_shape"], ds_name=cfg["ds_name"],
                      ensemble_mode=0, **cfg["core"])
    members = {}
    for k in range(cfg["members"]):
        torch.manual_seed(SEED + k)
        members[f"m{k}"] = v1t_b200.Model(margs, ds=make_ds(neurons)).to(dev)
    ens = v1t_b200.EnsembleModel(margs, members).to(dev)
    ens.train(False)
    rec = Recorder(members["m0"].core)
    host = host_batches(neurons, args.batch, rank, pin=True, in_shape=cfg["in_shape"])
    mouse = next(iter(neurons))
    hb = host[mouse]
    out_host = torch.empty((args.batch, neurons[mouse]), dtype=torch.float32).pin_memory()
    heat_host = torch.empty((args.batch,) + tuple(cfg["in_shape"][1:]), dtype=torch.float32).pin_memory()

    @torch.no_grad()
    def step(b):
        y, _, _ = ens(b["image"], mouse_id=mouse, behaviors=b["behavior"], pupil_centers=b["pupil_center"])
        _, attn = rec(images=b["image"], behaviors=b["behavior"], pupil_centers=b["pupil_center"], mouse_id=mouse)
        heat = attention_rollouts(attn, image_shape=b["image"].shape[2:])
        return y, heat

    resident = {k: v.to(dev) for k, v in hb.items()}

    def step_resident():
        return step(resident)

    def step_e2e():
        b = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
        y, heat = step(b)
        out_host.copy_(y, non_blocking=True)
        heat_host.copy_(heat, non_blocking=True)
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / steps

    for _ in range(max(args.warmup, 3)):
        step_resident()
    l0 = lib.v1t_launch_count()
    with ClockSampler(local) as clocks:
        ms = timed(step_resident, args.steps)
    launches = int(lib.v1t_launch_count() - l0)
    ms_e2e = timed(step_e2e, args.steps)
    peaks = load_peaks()
    D = core_dims(cfg, margs)
    # dominant kernel of this config: emit-P (materialised softmax(QK^T) of 4 blocks, 2.8 GB written) + rollout read
    stack_bytes = args.batch * D["blocks"] * D["H"] * D["T"] ** 2 * 4
    flops = (cfg["members"] + 1) * core_flops_fwd(**D) * args.batch
    line = {"metric": "ensemble inference + attention rollout samples/sec", "value": args.batch / (ms / 1e3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 (fp32 accumulate)", "data": "synthetic", "config": workload_config(args, cfg, world),
            "e2e": {"value": args.batch / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(sum(v.numel() * 4 for v in hb.values())),
                    "d2h_bytes_per_step": int(out_host.numel() * 4 + heat_host.numel() * 4)},
            "gpu_launches": launches, "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "kernel": "emit-P stack write (attn_fwd2_kernel<EMIT>) + rollout read (rollout_step_kernel)",
                         "achieved": 2 * stack_bytes / (ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": 2 * stack_bytes / (ms / 1e3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                         "note": "whole step time against the bytes of the [B,L,H,T,T] stack written once and read once; "
                                 "the step also runs 6 core forwards"},
            "model_tflops": flops / (ms / 1e3) / 1e12}
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="sensorium", choices=sorted(CONFIGS))
    ap.add_argument("--b200-impl", dest="b200_impl", default=None, help="fp32 | bf16x3 (exact) | bf16 (fast)")
    ap.add_argument("--mice", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--neurons", type=int, default=None)
    ap.add_argument("--micro", type=int, default=None, help="micro-batch size (0 = whole mouse batch at once)")
    ap.add_argument("--dp-mode", dest="dp_mode", default="mouse2d", choices=["batch", "mouse", "mouse2d"])
    ap.add_argument("--no-fuse-core", action="store_true",
                    help="one core pass per mouse (as the reference loops) instead of one over all mice of the step")
    ap.add_argument("--ref-batch", dest="ref_batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the optimizer / rollout / readout roofline runs")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    for k in ("mice", "batch", "neurons", "micro"):
        if getattr(args, k) is None:
            setattr(args, k, cfg[k])
    if args.b200_impl is None:
        args.b200_impl = os.environ.get("V1T_IMPL", cfg["impl"])
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.config == "ensemble":
        run_ensemble(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
