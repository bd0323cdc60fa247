"""CPU, world_size 2, gloo: the data-parallel gradient path (bucketed, hook-driven
all-reduce over the flat gradient buffer) must reproduce the world-size-1 gradients of
the same global batch (per-rank loss normalisation, then averaging: DDP semantics)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.GELU(), torch.nn.Linear(32, 32), torch.nn.GELU(), torch.nn.Linear(32, 5))


def _worker(rank, world, port, overlap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from linnaeus_b200.flat import FlatGroup
    from linnaeus_b200.parallel import DataParallel

    m = _model()
    if rank == 1:  # construction must broadcast rank 0's parameters
        with torch.no_grad():
            for p in m.parameters():
                p.add_(1.0)
    g = FlatGroup(list(m.parameters()), with_state=False)
    dp = DataParallel(m, [g], bucket_mb=0.0005, overlap=overlap)
    assert len(dp.buckets) > 1
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(8, 12, generator=gen)
    y = torch.randn(8, 5, generator=gen)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    out = []
    for step in range(2):  # two steps: counters must re-arm
        g.g.zero_()
        # micro-batch accumulation without sync, then the syncing micro-batch
        with dp.no_sync():
            ((dp(xs[:2]) - ys[:2]) ** 2).mean().mul(0.5).backward()
        ((dp(xs[2:]) - ys[2:]) ** 2).mean().mul(0.5).backward()
        dp.finish_gradients()
        out.append(g.g.clone())
    # plain numpy payloads: torch tensors travel through a Queue as shared file descriptors, which the parent may try to fetch
    # after this process has already exited
    q.put((rank, [o.numpy().copy() for o in out], g.p.detach().numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _launch(overlap):
    """Two gloo ranks on a free localhost port -> {rank: (outputs, params)}, or None when the rendezvous itself failed
    (the port can be taken between _free_port() and the workers' bind)."""
    import queue

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    try:
        for _ in range(2):
            r, out, params = q.get(timeout=180)
            res[r] = ([torch.from_numpy(o) for o in out], torch.from_numpy(params))
    except queue.Empty:
        res = None
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()  # the exact process we started
            res = None
        elif p.exitcode != 0:
            res = None
    return res


def _run(overlap):
    res = _launch(overlap)
    if res is None:  # one retry on a fresh port
        res = _launch(overlap)
    assert res is not None and set(res) == {0, 1}
    # single-process truth: mean over ranks of per-rank mean losses
    from linnaeus_b200.flat import FlatGroup

    m = _model()
    g = FlatGroup(list(m.parameters()), with_state=False)
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(8, 12, generator=gen)
    y = torch.randn(8, 5, generator=gen)
    loss = 0.5 * (((m(x[:4]) - y[:4]) ** 2).mean() + ((m(x[4:]) - y[4:]) ** 2).mean())
    loss.backward()
    for r in (0, 1):
        out, params = res[r]
        assert torch.equal(params, g.p.detach())  # broadcast happened
        for o in out:
            torch.testing.assert_close(o, g.g, rtol=1e-5, atol=1e-7)


def test_data_parallel_overlapped_buckets():
    _run(True)


def test_data_parallel_single_reduce_at_end():
    _run(False)
