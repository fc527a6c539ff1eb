"""world_size-2 gloo tests (CPU) of the data-parallel plumbing: mouse assignment and the flat-bucket gradient
all-reduce give the same result as a single process summing both shards."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from v1t_b200 import parallel


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
        out.put(([p.grad.clone() for p in params], [p.grad.clone() for p in shared + own], views_ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, got2, views_ok = q.get()
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
