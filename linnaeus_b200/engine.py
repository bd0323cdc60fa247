"""The training step of the hot path (R/train.py:115-377 at accumulation boundaries):
forward -> hierarchical loss -> backward -> (all-reduce) -> clip -> AdamW -> zero grads.

``TrainStep.step(images, meta, targets)`` is the eager path.  ``capture()`` records the
whole step - including the data-parallel all-reduces - into one CUDA graph over static
input buffers, which removes the ~1.5 k Python/launch overheads per step (the reference's
eager loop is launch bound at small batch); learning rate and step count then live in
device memory so replays stay valid.  Gradient accumulation (``accum_steps``) follows
R/train.py:173-192: loss / accum, optimizer step and all-reduce on the boundary only.
"""
from __future__ import annotations

import torch

from . import loss as LL
from .flat import weights_changed
from .optim import FlatAdamW
from .parallel import DataParallel


class TrainStep:
    """``TrainStep(model, opt, keys, num_classes, ..., dp=None, accum_steps=1)``.

    Every ``step()`` / ``replay()`` call is ONE iteration of the reference's loop (R/train.py:173-192): forward, loss / accum_steps,
    backward; on the accumulation boundary (every ``accum_steps``-th call) the gradients are all-reduced (data parallel: the earlier
    micro-batches run under ``no_sync``), clipped, applied with AdamW and zeroed.  With a :class:`DataParallel` wrapper built
    with ``average=False`` set ``optimizer.grad_scale = 1 / world`` so the averaging rides on the AdamW kernel."""

    def __init__(self, model, optimizer: FlatAdamW, task_keys: list[str], num_classes: dict[str, int], kind: str = "ce",
                 smoothing: float = 0.1, soft_matrices: dict | None = None, task_weights: dict | None = None, config=None,
                 dp: DataParallel | None = None, accum_steps: int = 1):
        self.model = model
        self.opt = optimizer
        self.keys = list(task_keys)
        self.config = config
        self.dp = dp
        self.accum_steps = max(1, int(accum_steps))
        dev = next(model.parameters()).device
        if kind == "ce":
            self.criteria = {k: LL.CrossEntropyLoss() for k in self.keys}
        elif kind == "ls":
            self.criteria = {k: LL.LabelSmoothingCrossEntropy(smoothing=smoothing) for k in self.keys}
        elif kind == "taxonomy":
            self.criteria = {k: LL.TaxonomyAwareLabelSmoothingCE(soft_matrices[k]).to(dev) for k in self.keys}
        else:
            raise ValueError(kind)
        self.weighting = LL.StaticTaskWeighting(self.keys, task_weights)
        self._inv_accum = torch.full((), 1.0 / self.accum_steps, dtype=torch.float32, device=dev) if self.accum_steps > 1 else None
        self._micro = 0  # micro-batches since the last optimizer step
        self._graph = None
        self._graph_micro = None
        self._static = None
        self._stage = None
        self._pending = False
        self.loss = None

    def _fwd_bwd(self, images, meta, targets: dict, sync: bool = True):
        net = self.dp if self.dp is not None else self.model
        if self.dp is not None and not sync:
            with self.dp.no_sync():
                return self._fwd_bwd_inner(net, images, meta, targets)
        return self._fwd_bwd_inner(net, images, meta, targets)

    def _fwd_bwd_inner(self, net, images, meta, targets):
        out = net(images, meta)
        total, comps, _ = LL.weighted_hierarchical_loss(out, targets, self.criteria, self.weighting, None, 0, config=self.config)
        if self._inv_accum is not None:
            total.backward(self._inv_accum)  # loss / accum_steps (train.py:173-175) without an extra pass
        else:
            total.backward()
        return total.detach()

    def _iteration(self, images, meta, targets, boundary: bool):
        loss = self._fwd_bwd(images, meta, targets, sync=boundary)
        if boundary:
            if self.dp is not None:
                self.dp.finish_gradients()
            self.opt.step()
            self.opt.zero_grad()
        return loss

    def _is_boundary(self) -> bool:
        return (self._micro + 1) % self.accum_steps == 0

    def step(self, images, meta, targets: dict) -> torch.Tensor:
        """One eager iteration (an optimizer step on the accumulation boundary); returns the (device) loss scalar."""
        boundary = self._is_boundary()
        loss = self._iteration(images, meta, targets, boundary)
        self._micro = 0 if boundary else self._micro + 1
        self.loss = loss
        return loss

    # ---- CUDA graph ---------------------------------------------------------
    def capture(self, images, meta, targets: dict, warmup: int = 3) -> None:
        """Capture the whole iteration - forward, loss, backward, the bucketed gradient all-reduces (NCCL collectives are
        captured on their own stream, so they overlap the rest of backward inside the graph), clip + AdamW, gradient zeroing -
        into one CUDA graph; with ``accum_steps > 1`` a second graph holds the non-boundary micro-batch (forward + backward,
        gradients accumulate).  The warm-up iterations run real optimizer steps on the capture batch, so parameters, Adam state
        and the step counter are snapshotted before and restored after: capture leaves the training state untouched."""
        dev = images.device
        self._static = (images.clone(), None if meta is None else meta.clone(), {k: v.clone() for k, v in targets.items()})
        si, sm, st = self._static
        self.opt.set_device_lr(self.opt.param_groups[0]["lr"])
        snap = [(f.p.clone(), f.m.clone(), f.v.clone()) if f is not None else None for f in self.opt.flat]
        step_snap = (self.opt._step, self.opt._step_dev.clone())
        self.opt.zero_grad()
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                if self.accum_steps > 1:
                    self._iteration(si, sm, st, False)
                self._iteration(si, sm, st, True)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from . import _lib

        self._graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count
        # thread-local capture mode: NCCL's watchdog thread polls CUDA events of earlier collectives; in the default global mode a
        # query from ANY thread invalidates the capture
        mode = "thread_local" if self.dp is not None else "global"
        with torch.cuda.graph(self._graph, capture_error_mode=mode):
            self._static_loss = self._iteration(si, sm, st, True)
        self.launches_per_step = _lib.launch_count - n0  # C-ABI kernel launches recorded in the boundary graph
        if self.accum_steps > 1:
            self._graph_micro = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_micro, pool=self._graph.pool(), capture_error_mode=mode):
                self._static_loss_micro = self._iteration(si, sm, st, False)
        # capture executes nothing, but the warm-up iterations did: put the training state back
        with torch.no_grad():
            for f, sn in zip(self.opt.flat, snap):
                if f is not None:
                    f.p.copy_(sn[0])
                    f.m.copy_(sn[1])
                    f.v.copy_(sn[2])
            self.opt._step = step_snap[0]
            self.opt._step_dev.copy_(step_snap[1])
        self.opt.zero_grad()
        weights_changed()
        self._micro = 0

    # ---- input pipelining ---------------------------------------------------
    def prefetch(self, images, meta, targets: dict | None) -> None:
        """Start the host->device copy of the NEXT step's inputs on a side stream (pinned host tensors), into staging
        buffers, so it overlaps the graph of the step in flight; the next ``replay()`` waits for it, moves the staged
        batch into the graph's static input buffers (device-to-device) and runs.  This is what a prefetching data
        loader does for the reference's ``train.py`` loop (``non_blocking=True`` copies one batch ahead)."""
        si, sm, st = self._static
        if self._stage is None:
            self._stage = (torch.empty_like(si), None if sm is None else torch.empty_like(sm), {k: torch.empty_like(v) for k, v in st.items()})
            self._copy_stream = torch.cuda.Stream(device=si.device)
            self._copy_done = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record()
        gi, gm, gt = self._stage
        cs = self._copy_stream
        cs.wait_event(self._stage_free)  # the previous staged batch has been moved into the static buffers
        with torch.cuda.stream(cs):
            gi.copy_(images, non_blocking=True)
            if gm is not None and meta is not None:
                gm.copy_(meta, non_blocking=True)
            if targets is not None:
                for k, v in targets.items():
                    gt[k].copy_(v, non_blocking=True)
            self._copy_done.record(cs)
        self._pending = True

    def _consume_prefetch(self) -> None:
        si, sm, st = self._static
        gi, gm, gt = self._stage
        cur = torch.cuda.current_stream()
        cur.wait_event(self._copy_done)
        si.copy_(gi, non_blocking=True)
        if sm is not None:
            sm.copy_(gm, non_blocking=True)
        for k in st:
            st[k].copy_(gt[k], non_blocking=True)
        self._stage_free.record(cur)
        self._pending = False

    def replay(self, images=None, meta=None, targets: dict | None = None) -> torch.Tensor:
        """One iteration through the captured graph (the boundary graph on every ``accum_steps``-th call)."""
        si, sm, st = self._static
        if self._pending:
            self._consume_prefetch()
        if images is not None:
            si.copy_(images, non_blocking=True)
        if meta is not None and sm is not None:
            sm.copy_(meta, non_blocking=True)
        if targets is not None:
            for k, v in targets.items():
                st[k].copy_(v, non_blocking=True)
        boundary = self._is_boundary()
        if boundary:
            self._graph.replay()
            self.opt._step += 1  # host mirror of the device step counter the captured AdamW advanced
            weights_changed()    # the captured optimizer kernel rewrote the parameters; no host-side step() ran to say so
            self.loss = self._static_loss
        else:
            self._graph_micro.replay()
            self.loss = self._static_loss_micro
        self._micro = 0 if boundary else self._micro + 1
        return self.loss


@torch.no_grad()
def validate_one_pass(config, model, data_loader, criteria: dict, task_weighting, mask_meta: bool = False, all_reduce: bool | None = None) -> dict:
    """One validation pass right after the hot path (R/validation.py:49-340): ``model.eval()``, forward, the hierarchical loss with
    null masking disabled (``is_validation=True``), and the tracker's per-batch accumulation - without a single host sync per
    batch: the loss sum stays on the device and the metrics go through ``HierMetricsAccumulator`` (one kernel per batch); the only
    device->host reads happen once, after the loop.  ``mask_meta`` zeroes the metadata (the reference's "val_mask_meta" phase).
    ``data_loader`` yields ``(images, targets_dict, aux_info, ...)``.  Returns ``{"loss": avg loss per batch, "batches": n,
    **HierMetricsAccumulator.compute()}``; under ``torch.distributed`` the sums are reduced over the ranks like the tracker's."""
    from .metrics import HierMetricsAccumulator

    was_training = model.training
    model.eval()
    acc = HierMetricsAccumulator()
    loss_sum, n_batches, dev = None, 0, None
    for batch in data_loader:
        images, targets, aux = batch[0], batch[1], batch[2]
        dev = images.device
        if mask_meta and aux is not None:
            aux = torch.zeros_like(aux)
        outputs = model(images, aux)
        total, _, _ = LL.weighted_hierarchical_loss(outputs, targets, criteria, task_weighting, None, 0, is_validation=True, config=config)
        loss_sum = total.detach().float() if loss_sum is None else loss_sum + total.detach().float()
        acc.update(outputs, targets)
        n_batches += 1
    model.train(was_training)
    out = acc.compute(all_reduce=all_reduce)
    local = torch.stack([loss_sum if loss_sum is not None else torch.zeros((), device=dev or "cpu"),
                         torch.tensor(float(n_batches), device=dev or "cpu")])
    import torch.distributed as dist

    if all_reduce or (all_reduce is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
    s, n = local.tolist()
    out["loss"] = s / n if n > 0 else 0.0
    out["batches"] = int(n)
    return out
