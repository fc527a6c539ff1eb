"""Per-contraction precision budget of the fused attention, run on the GPU box:

    python scripts/attn_prec_experiment.py > gpurun_out/attn_prec.json

For every V1T_ATTN_PREC setting (kernels.cuh: attn_prec_env) it reports (a) the error of responses / loss / every
gradient against the numpy fp64 oracle at the full sequence length (T = 1654, default widths, 2 blocks, B = 2,
N = 1000: the setting of tests/test_gpu_parity.py::test_full_default_shape_matches_oracle) and (b) the device time
of the attention kernels in the bench workload (7 mice x 16, CUDA events around the launches).
"""
import json
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SETTINGS = {0: "all three bf16x3 terms everywhere (default)", 1: "P / Pd / dS operand hi-only in the accumulating MMAs",
            2: "dP without the resident-lo (SS-form) term", 3: "1 + 2", 4: "dP from the hi planes only", 5: "1 + 4"}


def accuracy():
    from oracle import v1t_oracle as O
    from golden_util import rel_err
    from test_gpu_parity import _default_model, cu

    rng = np.random.default_rng(7)
    n, B = 1000, 2
    model, crit, cfg, sd = _default_model(n, 2, rng)
    model.train(True)
    images = rng.standard_normal((B, 1, 36, 64)).astype(np.float32)
    beh, pup = rng.uniform(size=(B, 3)).astype(np.float32), rng.uniform(size=(B, 2)).astype(np.float32)
    y_true = rng.uniform(0, 2, size=(B, n)).astype(np.float32)
    noise = rng.standard_normal((B, n, 2)).astype(np.float32)
    ref = O.path_forward_backward(sd, cfg, "A", images, beh, pup, y_true, ds_size=4500, noise=noise)
    out = {}
    for prec in SETTINGS:
        os.environ["V1T_ATTN_PREC"] = str(prec)
        model.zero_grad(set_to_none=True)
        im = cu(images).requires_grad_(True)
        y, _, _ = model(im, mouse_id="A", behaviors=cu(beh), pupil_centers=cu(pup), noise=cu(noise))
        loss = crit(y_true=cu(y_true), y_pred=y, mouse_id="A", batch_size=B)
        loss.backward()
        errs = {k: rel_err(p.grad.cpu().numpy(), ref["grads"][k]) for k, p in model.named_parameters()}
        worst = max(errs, key=errs.get)
        out[prec] = {"responses": rel_err(y.detach().cpu().numpy(), ref["y"]),
                     "loss": abs(loss.item() - ref["loss"]) / abs(ref["loss"]),
                     "dimages": rel_err(im.grad.cpu().numpy(), ref["dimages"]),
                     "worst_grad": errs[worst], "worst_grad_name": worst,
                     "qkv_grad_block0": errs["core.transformer.blocks.0.mha.to_qkv.weight"]}
    os.environ.pop("V1T_ATTN_PREC", None)
    return out


def timing(prec):
    env = dict(os.environ, V1T_ATTN_PREC=str(prec))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu-baseline",
                        "--no-eager-baseline", "--no-extras"], env=env, capture_output=True, text=True, timeout=600)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    ph = line["phases"]
    us = lambda k: 1e3 * ph[k]["ms_per_step"] / max(ph[k]["scopes_per_step"], 1)  # noqa: E731
    return {"samples_per_s": line["value"], "ms_per_step": line["ms_per_step"], "attn_bwd_kernel_us": us("attn_bwd_kernel"),
            "attn_fwd_kernel_us": us("attn_fwd_kernel")}


if __name__ == "__main__":
    res = {"settings": SETTINGS, "accuracy": accuracy()}
    res["timing"] = {p: timing(p) for p in SETTINGS}
    print(json.dumps(res, indent=1))
