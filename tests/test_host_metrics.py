"""CPU: the host-side finalisation of HierMetricsAccumulator (no kernel involved) - the tracker's arithmetic on counter rows, and the
cross-rank reduction under torch.distributed (gloo, world size 2) against the single-process result on the union of the batches."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from linnaeus_b200.metrics import HierMetricsAccumulator
from oracle import metrics_oracle as MO
from tests.support.golden_metrics import make_case


def _rows(batches, keys):
    return [MO.phase_counters([o[k] for k in keys], [t[k] for k in keys]).tolist() for o, t in batches]


@pytest.mark.parametrize("name", ["metrics_six", "metrics_small_heads", "metrics_all_null"])
def test_finalize_reproduces_the_tracker_arithmetic(name):
    keys, batches = make_case(name)
    got = HierMetricsAccumulator._finalize(_rows(batches, keys), keys, torch.device("cpu"), all_reduce=False)
    ref = MO.phase_metrics(batches, keys)
    assert got["samples"] == ref["samples"]
    for f in ("chain_accuracy", "partial_chain_accuracy"):
        assert got[f] == pytest.approx(ref[f], abs=1e-12)
    assert got["acc1"] == pytest.approx(ref["acc1"]) and got["acc3"] == pytest.approx(ref["acc3"])
    assert got["null_acc1"] == pytest.approx(ref["null_acc1"]) and got["non_null_acc1"] == pytest.approx(ref["non_null_acc1"])
    assert HierMetricsAccumulator._finalize([], keys, torch.device("cpu"), all_reduce=False)["samples"] == 0


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys, batches = make_case("metrics_six")
    mine = batches[rank::world]  # rank 0: batches 0 and 2, rank 1: batch 1
    out = HierMetricsAccumulator._finalize(_rows(mine, keys), keys, torch.device("cpu"))  # all_reduce=None -> initialised -> reduce
    q.put((rank, out))
    dist.destroy_process_group()


def test_finalize_all_reduces_over_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    # spawned children copy sys.path: earlier tests put /root/reference (which has its own `tests` package) in front of it,
    # and the children must resolve `tests.test_host_metrics` to THIS repository
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    saved = list(sys.path)
    sys.path.insert(0, repo)
    try:
        for p in procs:
            p.start()
    finally:
        sys.path[:] = saved
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    keys, batches = make_case("metrics_six")
    ref = MO.phase_metrics(batches, keys)
    for r in (0, 1):
        assert res[r]["samples"] == ref["samples"]
        assert res[r]["chain_accuracy"] == pytest.approx(ref["chain_accuracy"], abs=1e-12)
        assert res[r]["partial_chain_accuracy"] == pytest.approx(ref["partial_chain_accuracy"], abs=1e-12)
        assert res[r]["acc1"] == pytest.approx(ref["acc1"]) and res[r]["acc3"] == pytest.approx(ref["acc3"])
        assert res[r]["null_acc1"] == pytest.approx(ref["null_acc1"]) and res[r]["non_null_acc1"] == pytest.approx(ref["non_null_acc1"])
