"""The training step of the hot path (R/train.py:115-377 at accumulation boundaries):
forward -> hierarchical loss -> backward -> (all-reduce) -> clip -> AdamW -> zero grads.

``TrainStep.step(images, meta, targets)`` is the eager path.  ``capture()`` records the
whole step into one CUDA graph over static input buffers, which removes the ~1.5 k
Python/launch overheads per step (the reference's eager loop is launch bound at small
batch); learning rate and step count then live in device memory so replays stay valid.
"""
from __future__ import annotations

import torch

from . import loss as LL
from .flat import weights_changed
from .optim import FlatAdamW
from .parallel import DataParallel


class TrainStep:
    def __init__(self, model, optimizer: FlatAdamW, task_keys: list[str], num_classes: dict[str, int], kind: str = "ce",
                 smoothing: float = 0.1, soft_matrices: dict | None = None, task_weights: dict | None = None, config=None,
                 dp: DataParallel | None = None):
        self.model = model
        self.opt = optimizer
        self.keys = list(task_keys)
        self.config = config
        self.dp = dp
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
        self._graph = None
        self._static = None
        self._stage = None
        self._pending = False
        self.loss = None

    def _fwd_bwd(self, images, meta, targets: dict):
        net = self.dp if self.dp is not None else self.model
        out = net(images, meta)
        total, comps, _ = LL.weighted_hierarchical_loss(out, targets, self.criteria, self.weighting, None, 0, config=self.config)
        total.backward()
        return total.detach()

    def step(self, images, meta, targets: dict) -> torch.Tensor:
        """One eager optimizer step; returns the (device) loss scalar."""
        loss = self._fwd_bwd(images, meta, targets)
        if self.dp is not None:
            self.dp.finish_gradients()
        self.opt.step()
        self.opt.zero_grad()
        self.loss = loss
        return loss

    # ---- CUDA graph ---------------------------------------------------------
    def capture(self, images, meta, targets: dict, warmup: int = 3) -> None:
        """Capture forward+backward+optimizer into one graph (single-GPU, or the compute part
        of a data-parallel step: the all-reduce stays outside the graph)."""
        dev = images.device
        self._static = (images.clone(), None if meta is None else meta.clone(), {k: v.clone() for k, v in targets.items()})
        si, sm, st = self._static
        self.opt.set_device_lr(self.opt.param_groups[0]["lr"])
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._graph_body(si, sm, st)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from . import _lib

        self._graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count
        with torch.cuda.graph(self._graph):
            self._static_loss = self._graph_body(si, sm, st)
        self.launches_per_step = _lib.launch_count - n0  # C-ABI kernel launches recorded in the graph

    def _graph_body(self, si, sm, st):
        if self.dp is not None:
            with self.dp.no_sync():  # collectives stay outside the graph
                return self._fwd_bwd(si, sm, st)
        loss = self._fwd_bwd(si, sm, st)
        if self.dp is None:
            self.opt.step()
            self.opt.zero_grad()
        return loss

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
        self._graph.replay()
        weights_changed()  # the captured optimizer kernel rewrote the parameters; no host-side step() ran to say so
        if self.dp is not None:
            self.dp.finish_gradients()
            self.opt.step()
            self.opt.zero_grad()
        self.loss = self._static_loss
        return self._static_loss


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
