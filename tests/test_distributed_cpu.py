"""world_size-2 gloo tests of the data-parallel plumbing (bucketed gradient mean, env:// init) on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from autoregressive_diffusion_b200.train import GradientBuckets, init_distributed
    r, w, _ = init_distributed()
    assert (r, w) == (rank, world) and dist.get_backend() == "gloo"
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (7, 1000, 3, 64 * 64)]
    params.append(torch.nn.Parameter(torch.zeros(5)))           # never receives a gradient (like emb_time / out_res)
    for i, p in enumerate(params[:-1]):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    buckets = GradientBuckets(params, bucket_bytes=2048)        # several small buckets
    buckets.all_reduce_mean()
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(params[:-1]))
    ok = ok and params[-1].grad is None and len(buckets.buckets) >= 2
    # gradients now live in one flat buffer: a second reduction works in place on the same views
    for i, p in enumerate(params[:-1]):
        p.grad.fill_(float(rank) * (i + 1))
    buckets.all_reduce_mean()
    ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 0.5 * (i + 1))) for i, p in enumerate(params[:-1]))
    ok = ok and all(p.grad.data_ptr() >= buckets.flat.data_ptr() for p in params[:-1])
    out.put((rank, ok))
    dist.destroy_process_group()


def test_bucketed_gradient_mean_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(0, True), (1, True)]


def _worker_groups(rank, world, port, out):
    """Completion-ordered layout: the early gradient groups form the tail of the flat buffer, the group that completes first
    last, each in buckets of its own; the bucketed mean still reduces every section."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from autoregressive_diffusion_b200.train import GradientBuckets, init_distributed
    init_distributed()
    params = [torch.nn.Parameter(torch.zeros(n)) for n in (70, 1000, 30, 640, 130, 900)]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    g0, g1 = [params[4], params[5]], [params[1]]            # group 0 completes first (decoder), group 1 second (deep encoder)
    buckets = GradientBuckets(params, bucket_bytes=1024, early=[g0, g1])
    buckets.all_reduce_mean()
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(params))
    off = {id(p): (p.grad.data_ptr() - buckets.flat.data_ptr()) // 4 for p in params}
    late = [params[0], params[2], params[3]]
    ok = ok and max(off[id(p)] for p in late) < off[id(params[1])] < min(off[id(p)] for p in g0)      # late | group 1 | group 0
    ok = ok and buckets.bucket_group == sorted(buckets.bucket_group, key=lambda k: (k >= 0, -k))         # -1.., 1.., 0..
    ok = ok and set(buckets.bucket_group) == {-1, 0, 1} and buckets.n_late_buckets == buckets.bucket_group.count(-1)
    lo = 0
    for b, k in zip(buckets.buckets, buckets.bucket_group):   # buckets tile the buffer and never straddle a section
        ok = ok and b.data_ptr() == buckets.flat.data_ptr() + 4 * lo
        lo += b.numel()
        inside = [p for p in params if off[id(p)] < lo and off[id(p)] + p.numel() > lo - b.numel()]
        want = g0 if k == 0 else g1 if k == 1 else late
        ok = ok and all(any(p is q for q in want) for p in inside)
    ok = ok and lo == buckets.flat.numel()
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_completion_ordered_bucket_layout_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_groups, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [(0, True), (1, True)]
