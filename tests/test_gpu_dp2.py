"""2 x GPU (NCCL): data parallelism on the REAL model.  World-2 flat gradients (per-rank loss normalisation, then averaging: DDP
semantics, R/main.py:936-982) equal the world-1 gradients of the same global batch; the bucketed all-reduce overlapped with backward
(metadata branch on its side stream included) gives the same result as one blocking all-reduce; and a train step captured as ONE CUDA
graph with the NCCL collectives inside it leaves identical parameters on both ranks, equal to the single-process step on the global
batch.  Skipped with fewer than two GPUs (run with: gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp2.py -m gpu)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev, dtype):
    import linnaeus_b200 as L
    from linnaeus_b200.optim import FlatAdamW

    torch.manual_seed(0)
    cfg, nc = L.make_synthetic_config("sm", 64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(2, 1), conv_depths=(1, 1, 1, 1), n_tasks=3)
    model = L.build_model(cfg, nc).to(dev).set_compute_dtype(dtype).train()
    g = torch.Generator().manual_seed(1)
    B = 8
    x = torch.randn(B, 3, 64, 64, generator=g)
    meta = torch.randn(B, 15, generator=g)
    tg = {k: torch.randint(1, nc[k], (B,), generator=g) for k in nc}
    return cfg, nc, model, x, meta, tg, FlatAdamW


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from linnaeus_b200.engine import TrainStep
    from linnaeus_b200.parallel import DataParallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        out = {}
        q.put((rank, _run_rank(rank, world, dev)))
    except Exception as e:  # report instead of leaving the parent waiting on the queue
        import traceback

        q.put((rank, {"error": f"{e}\n{traceback.format_exc()}"}))
    finally:
        dist.destroy_process_group()


def _run_rank(rank, world, dev):
    from linnaeus_b200.engine import TrainStep
    from linnaeus_b200.parallel import DataParallel

    if True:
        out = {}
        for overlap in (True, False):
            cfg, nc, model, x, meta, tg, FlatAdamW = _build(dev, torch.float32)
            opt = FlatAdamW(model.named_parameters(), lr=1e-2, clip_grad=5.0, grad_scale=1.0 / world)
            dp = DataParallel(model, opt.flat, average=False, overlap=overlap, bucket_mb=0.25)
            h = x.shape[0] // world
            sl = slice(rank * h, (rank + 1) * h)
            ts = TrainStep(model, opt, list(nc), nc, kind="ce", config=cfg, dp=dp)
            ts._fwd_bwd(x[sl].to(dev), meta[sl].to(dev), {k: v[sl].to(dev) for k, v in tg.items()})
            dp.finish_gradients()
            torch.cuda.synchronize()
            out[f"grads_overlap{int(overlap)}"] = [(f.g / world).cpu().numpy() for f in opt.flat if f is not None]
            opt.zero_grad()
        # one captured step (all-reduce inside the graph), then one more replay
        cfg, nc, model, x, meta, tg, FlatAdamW = _build(dev, torch.float32)
        opt = FlatAdamW(model.named_parameters(), lr=1e-2, clip_grad=5.0, grad_scale=1.0 / world)
        dp = DataParallel(model, opt.flat, average=False, bucket_mb=0.25)
        ts = TrainStep(model, opt, list(nc), nc, kind="ce", config=cfg, dp=dp)
        h = x.shape[0] // world
        sl = slice(rank * h, (rank + 1) * h)
        xs, ms, ts_ = x[sl].to(dev), meta[sl].to(dev), {k: v[sl].to(dev) for k, v in tg.items()}
        ts.capture(xs, ms, ts_, warmup=2)
        ts.replay(xs, ms, ts_)
        ts.replay(xs, ms, ts_)
        torch.cuda.synchronize()
        out["params_graph"] = [f.p.cpu().numpy() for f in opt.flat if f is not None]
        out["steps"] = opt.step_count()
        return out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_world2_equals_world1_on_the_real_model():
    import numpy as np
    import torch.multiprocessing as mp

    from linnaeus_b200.engine import TrainStep

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=150) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for r in range(2):
        assert "error" not in res[r], res[r]["error"]
    # single-process reference on the global batch
    dev = torch.device("cuda", 0)
    cfg, nc, model, x, meta, tg, FlatAdamW = _build(dev, torch.float32)
    opt = FlatAdamW(model.named_parameters(), lr=1e-2, clip_grad=5.0)
    ts = TrainStep(model, opt, list(nc), nc, kind="ce", config=cfg)
    ts._fwd_bwd(x.to(dev), meta.to(dev), {k: v.to(dev) for k, v in tg.items()})
    torch.cuda.synchronize()
    ref_g = [f.g.cpu().numpy() for f in opt.flat if f is not None]
    gmax = max(float(np.abs(g).max()) for g in ref_g)
    for key in ("grads_overlap1", "grads_overlap0"):
        for r in range(2):
            for a, b in zip(res[r][key], ref_g):
                assert float(np.abs(a - b).max()) <= 1e-4 * gmax, (key, r)
        for a, b in zip(res[0][key], res[1][key]):
            assert np.array_equal(a, b), key  # the all-reduce leaves bit-identical buffers on both ranks
    opt.zero_grad()
    for _ in range(2):
        ts.step(x.to(dev), meta.to(dev), {k: v.to(dev) for k, v in tg.items()})
    torch.cuda.synchronize()
    ref_p = [f.p.cpu().numpy() for f in opt.flat if f is not None]
    assert res[0]["steps"] == 2 and res[1]["steps"] == 2
    for a, b in zip(res[0]["params_graph"], res[1]["params_graph"]):
        assert np.array_equal(a, b)  # replicas stay in lock step
    for a, b in zip(res[0]["params_graph"], ref_p):
        # Adam turns noise on ~0 gradients into O(lr) steps (see tests/test_gpu_model.py): compare in relative L2 over the buffer
        assert float(np.linalg.norm(a - b)) <= 2e-3 * float(np.linalg.norm(b)), float(np.linalg.norm(a - b))
