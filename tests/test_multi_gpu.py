"""Two-rank NCCL test of the data-parallel step with the REAL kernels: the gradients a k-GPU
step leaves on a rank equal those of one GPU processing the k-times larger batch.  Needs >= 2 visible GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

MICE = {"A": 96, "B": 80, "C": 64}
BATCH = 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(dev):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import v1t_b200
    from bench import make_args, make_ds

    margs = make_args(MICE, dev, impl="bf16x3", patch_stride=4, num_blocks=2, p_dropout=0.0, t_dropout=0.0)
    torch.manual_seed(99)
    model = v1t_b200.Model(margs, ds=make_ds(MICE)).to(dev)
    crit = v1t_b200.get_criterion(margs, ds=make_ds(MICE))
    with torch.no_grad():
        g = torch.Generator(device=dev).manual_seed(5)
        for r in model.readouts.values():
            r.features.add_(torch.randn(r.features.shape, device=dev, generator=g) * 0.05)
    model.train(False)  # deterministic: no dropout, positions at mu
    return model, crit


def _global_batches(plan, dev):
    g = torch.Generator().manual_seed(4242)
    out = {}
    for m, n in MICE.items():
        gb = plan.global_batch[m]
        out[m] = {"image": torch.randn((gb, 1, 36, 64), generator=g).to(dev), "behavior": torch.rand((gb, 3), generator=g).to(dev),
                  "pupil_center": torch.rand((gb, 2), generator=g).to(dev),
                  "response": (torch.rand((gb, n), generator=g) * 2).to(dev)}
    return out


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from v1t_b200 import parallel

    parallel.init_from_env(backend="nccl")
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    model, crit = _model(dev)
    plan = parallel.make_plan(list(MICE), rank, world, mode, BATCH)
    full = _global_batches(plan, dev)
    mine = {m: {k: v[lo:hi] for k, v in full[m].items()} for m, (lo, hi) in plan.my_slices.items()}
    sync = parallel.GradSync(model, plan)
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        parallel.sweep(model, crit, mine, plan.global_batch, sync, fused_accumulate=True)
    got = {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
    # the same step on ONE GPU: every mouse's whole global batch, plain autograd accumulation
    model.zero_grad(set_to_none=True)
    parallel.sweep(model, crit, full, plan.global_batch, None, fused_accumulate=False)
    want = {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None}
    torch.cuda.synchronize(dev)
    q.put((rank, got, want, sorted(plan.my_slices)))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", ["mouse2d", "batch"])
def test_two_gpu_step_equals_one_gpu_step_with_twice_the_batch(mode):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get() for _ in range(world)]
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for rank, got, want, my_mice in results:
        worst = 0.0
        for k, w in want.items():
            private = k.split(".")[0] in ("readouts", "core_shifter")
            if private and k.split(".")[1] not in my_mice:
                assert k not in got or float(np.abs(got[k]).max()) == 0.0, (rank, k)  # never left its group
                continue
            assert k in got, (rank, k)
            denom = max(float(np.abs(w).max()), 1e-30)
            e = float(np.abs(got[k] - w).max()) / denom
            worst = max(worst, e)
            # split-batch sums differ from the one-pass sums only by fp32 / bf16x3 summation order
            assert e < 2e-4, (mode, rank, k, e)
        print(f"[{mode}] rank {rank}: worst relative gradient difference {worst:.2e}")


def test_fused_core_pass_equals_per_mouse_passes():
    """sweep(fuse_core=True): ONE core pass over the concatenated batches of all mice gives the gradients of the
    reference's mouse-by-mouse accumulation (train.py:84-111), and is refused for per-mouse behaviour MLPs."""
    from v1t_b200 import parallel

    dev = torch.device("cuda", 0)
    model, crit = _model(dev)
    plan = parallel.make_plan(list(MICE), 0, 1, "mouse2d", BATCH)
    g = torch.Generator().manual_seed(11)
    batches = {}
    for i, (m, n) in enumerate(MICE.items()):
        rows = BATCH + i  # unequal batch sizes per mouse
        batches[m] = {"image": torch.randn((rows, 1, 36, 64), generator=g).to(dev), "behavior": torch.rand((rows, 3), generator=g).to(dev),
                      "pupil_center": torch.rand((rows, 2), generator=g).to(dev),
                      "response": (torch.rand((rows, n), generator=g) * 2).to(dev)}
    gb = {m: BATCH for m in MICE}
    out = []
    for fuse in (False, True):
        model.zero_grad(set_to_none=True)
        loss = parallel.sweep(model, crit, batches, gb, None, fused_accumulate=True, fuse_core=fuse)
        out.append((float(loss), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}))
    assert abs(out[0][0] - out[1][0]) / abs(out[0][0]) < 1e-5
    assert set(out[0][1]) == set(out[1][1])
    for k, gref in out[0][1].items():
        denom = max(float(gref.abs().max()), 1e-30)
        assert float((out[1][1][k] - gref).abs().max()) / denom < 1e-4, k
    model.core.behavior_mode = 4
    assert not parallel._can_fuse_core(model, batches, 0)
    model.core.behavior_mode = 3
    assert parallel._can_fuse_core(model, batches, 0) and not parallel._can_fuse_core(model, batches, 8)
