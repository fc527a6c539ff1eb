"""world_size-2 gloo tests (CPU) of the data-parallel plumbing: mouse assignment and the flat-bucket gradient
all-reduce give the same result as a single process summing both shards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from v1t_b200 import parallel



def _to_np(x):
    """Queue payloads travel by value: a tensor put on a multiprocessing queue is shared through a file descriptor that
    dies with the sending process (the parent then fails with ConnectionResetError if the worker has already exited)."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy().copy()
    if isinstance(x, dict):
        return {k: _to_np(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_np(v) for v in x)
    return x


def _to_t(x):
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x)
    if isinstance(x, dict):
        return {k: _to_t(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_to_t(v) for v in x)
    return x


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s)) for s in ((7, 5), (1000,), (3,), (64, 33))]
    frozen = torch.nn.Parameter(torch.randn(4), requires_grad=False)
    g = torch.Generator().manual_seed(100 + rank)
    for i, p in enumerate(params):
        if not (rank == 1 and i == 2):  # rank 1 has no gradient for param 2 (mouse-sharded readout case)
            p.grad = torch.randn(p.shape, generator=g)
    sync = parallel.GradSync(params + [frozen], bucket_mb=0.002)  # force several buckets
    assert len(list(sync.buckets())) > 1
    sync.all_reduce()
    # second exchange: the first two parameters live in a flat gradient sink (shared-core case): reduced in place
    from v1t_b200.functional import GradSink

    shared = [torch.nn.Parameter(torch.randn(s)) for s in ((5, 3), (130,))]
    own = [torch.nn.Parameter(torch.randn(s)) for s in ((9,), (4, 4))]
    sink = GradSink(shared)
    sink.arm()
    g2 = torch.Generator().manual_seed(200 + rank)
    for p in shared:
        p.grad.copy_(torch.randn(p.shape, generator=g2))  # written through the view, like the core backward's add
    for p in own:
        p.grad = torch.randn(p.shape, generator=g2)
    sync2 = parallel.GradSync(shared + own, bucket_mb=0.0002)
    sync2.all_reduce(sinks=[sink])
    views_ok = all(p.grad.data_ptr() == sink.view_of(p, sink.flat).data_ptr() for p in shared)
    if rank == 0:
        out.put(_to_np(([p.grad.clone() for p in params], [p.grad.clone() for p in shared + own], views_ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, got2, views_ok = _to_t(q.get())
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    shapes = ((7, 5), (1000,), (3,), (64, 33))
    want = []
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    for i, s in enumerate(shapes):
        tot = torch.zeros(s)
        for r in range(world):
            if not (r == 1 and i == 2):
                tot += torch.randn(s, generator=gens[r])
        want.append(tot)
    for a, b in zip(got, want):
        assert torch.allclose(a, b, atol=1e-6)
    # sink exchange: same sums, and the shared parameters' .grad still view the sink's flat buffer
    assert views_ok
    shapes2 = ((5, 3), (130,), (9,), (4, 4))
    gens2 = [torch.Generator().manual_seed(200 + r) for r in range(world)]
    for i, s in enumerate(shapes2):
        tot = torch.zeros(s)
        for r in range(world):
            tot += torch.randn(s, generator=gens2[r])
        assert torch.allclose(got2[i], tot, atol=1e-6), i


def test_mouse_assignment():
    mice = list("ABCDEFG")
    assert parallel.mice_of_rank(mice, 3, 8, "batch") == mice
    owned = [parallel.mice_of_rank(mice, r, 8, "mouse") for r in range(8)]
    assert sorted(sum(owned, [])) == mice and owned[7] == []
    assert parallel.mice_of_rank(mice, 1, 2, "mouse") == ["B", "D", "F"]


# ---- sweep() end to end on two ranks with a CPU stand-in for the model ------------------------------------------
class _SinkLinear(torch.autograd.Function):
    """y = x @ W^T with the core's sink protocol (functional._CoreFunction.backward): when the sink is armed the
    weight gradient is added into the sink's flat buffer and autograd is told there is none."""

    @staticmethod
    def forward(ctx, x, w, sink_arg):
        ctx.save_for_backward(x, w)
        ctx.sink = sink_arg
        return x @ w.t()

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        gw = dy.t() @ x
        if ctx.sink is not None and ctx.sink[0].armed:
            sink, slots = ctx.sink
            tmp = torch.zeros_like(sink.flat)
            tmp[slots[0][0]:slots[0][0] + slots[0][1]].view(w.shape).copy_(gw)
            sink.flat.add_(tmp)
            gw = None
        return None, gw, None


class _FakeCore(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.randn(6, 4))
        self.unused = torch.nn.Parameter(torch.randn(3))  # e.g. another mouse's behaviour MLP: zero gradient
        self.frozen = False
        self.grad_sink = None

    def fused_grad_accumulation(self, on=True):
        from v1t_b200.functional import GradSink

        if not on:
            if self.grad_sink is not None:
                self.grad_sink.disarm()
            return None
        if self.grad_sink is None:
            self.grad_sink = GradSink(list(self.parameters()))
        self.grad_sink.arm()
        return self.grad_sink

    def forward(self, x):
        arg = (self.grad_sink, self.grad_sink.slots([self.w])) if (self.grad_sink and self.grad_sink.armed) else None
        return _SinkLinear.apply(x, self.w, arg)


class _FakeModel(torch.nn.Module):
    def __init__(self, mice):
        super().__init__()
        self.core = _FakeCore()
        self.readouts = torch.nn.ModuleDict({m: torch.nn.Linear(6, 5) for m in mice})

    def forward(self, inputs, mouse_id, behaviors, pupil_centers):
        return self.readouts[mouse_id](self.core(inputs)), None, None


def _sweep_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_from_env(backend="gloo")
    mice = ["A", "B", "C"]
    torch.manual_seed(3)
    model = _FakeModel(mice)
    g = torch.Generator().manual_seed(50 + rank)
    batches = {m: {"image": torch.randn(4, 4, generator=g), "behavior": None, "pupil_center": None,
                   "response": torch.randn(4, 5, generator=g)} for m in mice}
    crit = lambda y_true, y_pred, mouse_id, batch_size: ((y_pred - y_true) ** 2).sum() / batch_size  # noqa: E731
    gb = {m: 4 * world for m in mice}
    sync = parallel.GradSync(model.parameters(), bucket_mb=0.0001)
    results = []
    for fused in (False, True, True):  # the second fused sweep starts from set_to_none gradients again
        model.zero_grad(set_to_none=True)
        total = parallel.sweep(model, crit, batches, gb, sync, fused_accumulate=fused)
        results.append(({k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()},
                        float(total)))
    if rank == 0:
        out.put(_to_np((results, {k: v.detach().clone() for k, v in model.state_dict().items()})))
    dist.barrier()
    dist.destroy_process_group()


def test_sweep_two_ranks_fused_accumulation_matches_plain_and_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_sweep_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results, sd = _to_t(q.get())
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    (plain, loss0), (fused, loss1), (fused2, loss2) = results
    assert loss0 == loss1 == loss2
    for k in plain:
        if plain[k] is None:  # never touched without the sink; the sink reports an explicit zero gradient
            assert fused[k] is None or float(fused[k].abs().max()) == 0.0
            continue
        assert torch.allclose(fused[k], plain[k], atol=1e-6), k
        assert torch.allclose(fused2[k], plain[k], atol=1e-6), k
    # single-process reference: both ranks' batches through plain autograd on one copy of the model
    mice = ["A", "B", "C"]
    model = _FakeModel(mice)
    model.load_state_dict(sd)
    for rank in range(world):
        g = torch.Generator().manual_seed(50 + rank)
        for m in mice:
            x, y = torch.randn(4, 4, generator=g), torch.randn(4, 5, generator=g)
            pred, _, _ = model(x, m, None, None)
            (((pred - y) ** 2).sum() / (4 * world)).backward()
    for k, p in model.named_parameters():
        if p.grad is not None:
            assert torch.allclose(plain[k], p.grad, atol=1e-5), k


# ---- rank plans (2-D mouse layout) and the group-wise exchange ----------------------------------------------------
def test_plan_mouse2d_partitions_every_mouse_exactly_once_and_balances_ranks():
    mice = list("ABCDEFG")
    for world in (1, 2, 3, 4, 8):
        batch = 16
        plans = [parallel.make_plan(mice, r, world, "mouse2d", batch) for r in range(world)]
        gb = batch * world
        assert all(p.global_batch == {m: gb if world > 1 else batch for m in mice} for p in plans)
        for p in plans:  # equal work, no idle rank, at most two mice at 7 mice / 8 ranks
            rows = sum(hi - lo for lo, hi in p.my_slices.values())
            assert rows == len(mice) * batch
        if world == 8:
            assert all(1 <= len(p.my_slices) <= 2 for p in plans)
        for m in mice:  # the slices of one mouse tile [0, global batch) without gaps or overlap
            cover = sorted(p.my_slices[m] for p in plans if m in p.my_slices)
            assert cover[0][0] == 0 and cover[-1][1] == plans[0].global_batch[m]
            assert all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
            assert plans[0].group_of(m) == tuple(r for r, p in enumerate(plans) if m in p.my_slices)
    assert parallel.make_plan(mice, 0, 8, "mouse2d", 16).scaling == "weak"
    assert parallel.make_plan(mice, 0, 8, "mouse", 16).scaling == "strong"
    assert parallel.make_plan(mice, 7, 8, "mouse", 16).my_slices == {}


def _plan_worker(rank, world, port, mode, out, n_mice=4):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_from_env(backend="gloo")
    mice = [chr(ord("A") + i) for i in range(n_mice)]
    batch = 3
    torch.manual_seed(3)
    model = _FakeModel(mice)
    plan = parallel.make_plan(mice, rank, world, mode, batch)
    sync = parallel.GradSync(model, plan)
    g = torch.Generator().manual_seed(77)  # the GLOBAL batch of every mouse, identical on all ranks
    full = {m: {"image": torch.randn(plan.global_batch[m], 4, generator=g), "behavior": None, "pupil_center": None,
                "response": torch.randn(plan.global_batch[m], 5, generator=g)} for m in mice}
    mine = {m: {k: (v[lo:hi] if v is not None else None) for k, v in full[m].items()}
            for m, (lo, hi) in plan.my_slices.items()}
    crit = lambda y_true, y_pred, mouse_id, batch_size: ((y_pred - y_true) ** 2).sum() / batch_size  # noqa: E731
    for _ in range(2):  # the second sweep re-arms the flat buffers after set_to_none
        model.zero_grad(set_to_none=True)
        parallel.sweep(model, crit, mine, plan.global_batch, sync, fused_accumulate=True, micro_batch=2)
    grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    out.put(_to_np((rank, grads, dict(plan.my_slices), {k: v.detach().clone() for k, v in model.state_dict().items()})))
    dist.barrier()
    dist.destroy_process_group()


def _run_plan_case(world, mode, n_mice=4):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_plan_worker, args=(r, world, port, mode, q, n_mice)) for r in range(world)]
    for p in procs:
        p.start()
    got = [_to_t(q.get()) for _ in range(world)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got.sort(key=lambda t: t[0])
    mice = [chr(ord("A") + i) for i in range(n_mice)]
    gbatch = parallel.make_plan(mice, 0, world, mode, 3).global_batch
    model = _FakeModel(mice)
    model.load_state_dict(got[0][3])
    g = torch.Generator().manual_seed(77)
    for m in mice:  # single process: the whole global batch of every mouse through plain autograd
        x, y = torch.randn(gbatch[m], 4, generator=g), torch.randn(gbatch[m], 5, generator=g)
        pred, _, _ = model(x, m, None, None)
        (((pred - y) ** 2).sum() / gbatch[m]).backward()
    want = {k: p.grad for k, p in model.named_parameters()}
    for rank, grads, slices, _ in got:
        assert torch.allclose(grads["core.w"], want["core.w"], atol=1e-5), (rank, "core")  # shared core: every rank
        for m in mice:
            for suffix in ("weight", "bias"):
                k = f"readouts.{m}.{suffix}"
                if m in slices:  # a readout is complete on every rank of its mouse's group ...
                    assert torch.allclose(grads[k], want[k], atol=1e-5), (rank, k)
                else:            # ... and never reaches the others
                    assert grads[k] is None, (rank, k)


def test_sweep_mouse2d_three_ranks_gloo_matches_single_process():
    _run_plan_case(3, "mouse2d")


def test_sweep_batch_mode_two_ranks_gloo_matches_single_process():
    _run_plan_case(2, "batch")


def test_sweep_mouse2d_eight_ranks_seven_mice_gloo_matches_single_process():
    """The driver's largest case: 7 mice on 8 ranks (every rank holds rows of one or two mice; most readout groups
    are pairs of neighbouring ranks)."""
    _run_plan_case(8, "mouse2d", n_mice=7)

