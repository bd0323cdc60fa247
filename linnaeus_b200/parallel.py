"""Single-node data parallelism for the hot path (one process per GPU).

Semantics of the reference's ``DDP(model)`` (R/main.py:936-982): every rank runs the
same model on its shard of the global batch, the loss is normalised per rank, and the
parameter gradients are **averaged** across ranks before the optimizer step.

B200-first mechanics: gradients already live in contiguous float32 buffers
(:mod:`linnaeus_b200.flat`), laid out in parameter-registration order, which is the
reverse of the order in which backward produces them.  The buffer is cut into buckets;
a post-accumulate hook on every parameter counts its bucket down, and when the last
gradient of a bucket has landed the bucket's slice is all-reduced asynchronously
(NCCL over NVLink/NVSwitch; the collective runs on NCCL's stream, overlapping the rest
of backward).  ``TrainStep.capture`` records these collectives INSIDE the step's CUDA graph
(NCCL is capturable), so the benchmarked graph path overlaps them exactly like the eager one.
With ``average=False`` the 1/world factor is folded into the optimizer kernel
(``FlatAdamW.grad_scale``) instead of a separate pass over the gradients.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from .flat import FlatGroup, weights_changed


class _Bucket:
    __slots__ = ("group", "lo", "hi", "n_params", "pending", "work", "streams")

    def __init__(self, group, lo, hi, n_params):
        self.group, self.lo, self.hi, self.n_params = group, lo, hi, n_params
        self.pending = n_params
        self.work = None
        self.streams = {}  # stream id -> last "gradient written" event of every CUDA stream that produced gradients of this bucket


def plan_buckets(groups: list[FlatGroup], bucket_bytes: int, first_bucket_bytes: int | None = None) -> tuple[list[_Bucket], dict[int, int]]:
    """Cut each flat group into contiguous buckets of about ``bucket_bytes``; returns the
    buckets and a map id(param) -> bucket index.

    The FIRST bucket of a group (the first-registered parameters: the stem and the convolutional stages) is the last one whose
    gradients become final, so its all-reduce can overlap nothing: it is kept small (``first_bucket_bytes``, default a quarter of
    ``bucket_bytes``) so that the exposed collective at the end of backward is latency- rather than bandwidth-sized -- the mirror image
    of torch DDP's small first bucket (there: the first gradients READY, to start communicating early)."""
    buckets, owner = [], {}
    full_cap = max(1, bucket_bytes // 4)
    first_cap = max(1, (bucket_bytes // 4 if first_bucket_bytes is None else first_bucket_bytes) // 4)
    for gi, g in enumerate(groups):
        lo, count = 0, 0
        cap = min(first_cap, full_cap)
        for i, p in enumerate(g.params):
            a, b = g.span(i)
            end = g.offsets[i + 1] if i + 1 < len(g.params) else g.numel  # next tensor's (aligned) start
            count += 1
            owner[id(p)] = len(buckets)
            last = i == len(g.params) - 1
            if end - lo >= cap or last:
                buckets.append(_Bucket(gi, lo, g.numel if last else end, count))
                lo, count = end, 0
                cap = full_cap
    return buckets, owner


class DataParallel(nn.Module):
    """``DataParallel(model, flat_groups)`` -- call ``finish_gradients()`` after backward.

    ``flat_groups`` are the optimizer's :class:`FlatGroup` s (``FlatAdamW.flat``).  With
    ``average=True`` the reduced buffers are divided by the world size here; pass
    ``average=False`` when the optimizer folds 1/world (``grad_scale``)."""

    def __init__(self, module: nn.Module, flat_groups: list[FlatGroup], bucket_mb: float = 25.0, process_group=None,
                 average: bool = True, overlap: bool = True):
        super().__init__()
        self.module = module
        self.groups = [g for g in flat_groups if g is not None]
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        self.overlap = overlap
        self.require_sync = True
        self.buckets, self._owner = plan_buckets(self.groups, int(bucket_mb * (1 << 20)))
        self._seen: set[int] = set()
        self._hooks = []
        if self.world > 1 and overlap:
            for g in self.groups:
                for p in g.params:
                    if p.requires_grad:
                        self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
                        p._lnx_grad_ready = self._on_grad  # kernels that write p.grad directly bypass the autograd hook
        self.broadcast_parameters()

    def broadcast_parameters(self) -> None:
        """Rank 0's parameters to everyone (DDP does the same at construction)."""
        if self.world > 1:
            for g in self.groups:
                dist.broadcast(g.p, src=0, group=self.pg)
        weights_changed()

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # attribute passthrough so callers touching model.head / .use_checkpoint keep working
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.module, name)

    def _launch(self, b: _Bucket) -> None:
        buf = self.groups[b.group].g[b.lo:b.hi]
        if buf.is_cuda and b.streams:
            # NCCL orders the collective after the CURRENT stream only: wait for the last gradient write of every other stream
            # that produced gradients of this bucket (under CUDA-graph capture the event pair becomes a graph edge)
            cur = torch.cuda.current_stream().cuda_stream
            for sid, ev in b.streams.items():
                if sid != cur:
                    torch.cuda.current_stream().wait_event(ev)
            b.streams = {}
        b.work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def _on_grad(self, p: torch.Tensor) -> None:
        if not self.require_sync:
            return
        # one notification per parameter and iteration: a kernel that accumulates straight into p.grad reports through
        # p._lnx_grad_ready, and autograd's post-accumulate hook can fire for the same parameter later in the same backward
        if id(p) in self._seen:
            return
        self._seen.add(id(p))
        b = self.buckets[self._owner[id(p)]]
        if p.is_cuda and self.world > 1:
            s = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(s)  # marks "this parameter's gradient is written" on the producing stream; a later event on the same stream subsumes it
            b.streams[s.cuda_stream] = ev
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def no_sync(self):
        """Context manager for gradient accumulation micro-batches (reduce on the last one only)."""
        dp = self

        class _Ctx:
            def __enter__(self):
                dp.require_sync = False

            def __exit__(self, *exc):
                dp.require_sync = True

        return _Ctx()

    def finish_gradients(self) -> None:
        """Launch whatever has not been reduced yet, wait for all buckets, average (unless the optimizer folds 1/world).
        Capturable: inside a CUDA-graph capture the collectives and the joins become graph nodes / edges."""
        if self.world == 1:
            return
        for b in self.buckets:
            if b.work is None:
                self._launch(b)
        for b in self.buckets:
            b.work.wait()
            b.work = None
            b.pending = b.n_params
            b.streams = {}
        self._seen.clear()
        if self.average:
            for g in self.groups:
                g.g.mul_(1.0 / self.world)

    def reset(self) -> None:
        """Forget bucket progress (after a micro-batch that ran under ``no_sync`` nothing is pending anyway; this is for error paths)."""
        for b in self.buckets:
            b.work = None
            b.pending = b.n_params
            b.streams = {}
        self._seen.clear()
