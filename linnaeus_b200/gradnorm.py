"""GradNorm task weighting right around the hot path.

Reference: ``linnaeus.loss.gradnorm.GradNormModule`` (R/loss/gradnorm.py:35-302) and the re-forward driver
``GradientWeighting.update_gradnorm_weights_reforward`` (R/loss/gradient_weighting.py:367-923), which - every
``UPDATE_INTERVAL`` steps - runs, PER TASK, a full forward (other tasks' logits detached), the task's mean loss over non-null
samples, ``torch.autograd.grad`` w.r.t. the shared backbone parameters, and the L2 norm of the flattened gradient; then
``measure_and_update`` rescales the task weights.

Here:

* ``GradNormModule`` - same constructor, buffers (``task_weights``, ``initial_losses``), init strategies and
  ``measure_and_update(unweighted_losses, grad_tensors)`` arithmetic (it accepts flattened gradient vectors like the reference or
  the already reduced norms as 0-dim tensors: ``x.norm(2)`` of a non-negative scalar is the scalar).  ``return_metrics=False``
  skips the reference's ~5K ``.item()`` reads.
* ``task_gradient_norms`` - ONE forward of the B200 model, then K backward passes through the retained graph, one per task
  cotangent (the shared trunk is not recomputed K times; the reference's K re-forwards are not needed because detaching the other
  tasks' logits only removes their cotangents).  Gradients land in ``param.grad`` (the flat buffer when FlatAdamW owns the
  parameters), the norm over the backbone tensors is one ``torch._foreach_norm``.  Call it between optimizer steps: it expects the
  gradients to be unused at entry and leaves them zeroed.

No kernels of its own: it drives the model's existing forward / backward kernels; the weight arithmetic is a handful of torch
ops on K-element tensors.
"""
from __future__ import annotations

import math
from typing import Any

import torch
import torch.distributed as dist
from torch import nn

__all__ = ["GradNormModule", "GradientWeighting", "backbone_parameters", "task_gradient_norms", "update_gradnorm_weights"]


def _allreduce_mean(value: torch.Tensor) -> torch.Tensor:
    """R/loss/gradnorm.py:20-32."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return value
    dist.all_reduce(value, op=dist.ReduceOp.SUM)
    return value / dist.get_world_size()


class GradNormModule(nn.Module):
    """Per-task weights driven towards equal (alpha = 0) or loss-ratio-shaped (alpha > 0) gradient norms on the shared backbone."""

    def __init__(self, task_keys: list[str], alpha: float = 1.5, init_weights: torch.Tensor | None = None,
                 label_densities: dict[str, float] | None = None, num_classes: dict[str, int] | None = None,
                 init_strategy: str = "inverse_density", config: Any | None = None):
        super().__init__()
        self.num_tasks = len(task_keys)
        self.task_keys = task_keys
        self.alpha = alpha
        self.config = config
        if init_weights is None:
            init_weights = self._compute_init_weights(task_keys, label_densities, num_classes, init_strategy)
        self.register_buffer("task_weights", init_weights.clone())
        self.register_buffer("initial_losses", torch.zeros(self.num_tasks))
        self.has_initted = False

    def _compute_init_weights(self, task_keys, label_densities=None, num_classes=None, strategy="inverse_density") -> torch.Tensor:
        """R/loss/gradnorm.py:96-155: equal weights without densities; 1/density (floored at 0.001), optionally times
        log(C_k)/log(C_max); normalised to sum to the number of tasks."""
        if not label_densities:
            return torch.ones(len(task_keys), dtype=torch.float32)
        dens = [label_densities.get(k, 1.0) for k in task_keys]
        if strategy == "inverse_density" or (strategy == "class_complexity" and num_classes is None):
            w = [1.0 / max(d, 0.001) for d in dens]
        elif strategy == "class_complexity":
            counts = [num_classes.get(k, 1) for k in task_keys]
            top = max(counts)
            w = [1.0 / max(d, 0.001) * (math.log(c) / math.log(top)) for d, c in zip(dens, counts)]
        else:
            w = [1.0] * self.num_tasks
        total = sum(w)
        return torch.tensor([x * self.num_tasks / total for x in w], dtype=torch.float32)

    def forward(self, losses: dict[str, torch.Tensor]) -> torch.Tensor:
        """sum_k w_k * loss_k over the SORTED task keys (R/loss/gradnorm.py:157-166)."""
        return (torch.stack([losses[k] for k in sorted(losses.keys())]) * self.task_weights).sum()

    @torch.no_grad()
    def measure_and_update(self, unweighted_losses: dict[str, torch.Tensor], grad_tensors: dict[str, torch.Tensor],
                           return_metrics: bool = True) -> dict[str, Any]:
        """R/loss/gradnorm.py:168-302.  Task order = sorted(task_keys).  First call with alpha > 0 records the (rank-mean) losses
        as ``initial_losses``.  target_k = mean(norms) * (normalised loss ratio_k)^alpha; w_k *= norm_k / target_k (tasks with
        target < 1e-8 keep their weight); weights renormalised to sum to K."""
        tasks = sorted(self.task_keys)
        K = len(tasks)
        dev = next(iter(grad_tensors.values())).device if grad_tensors else torch.device("cpu")
        loss_values = torch.zeros(K, device=dev)
        for i, tk in enumerate(tasks):
            if tk in unweighted_losses:
                loss_values[i] = unweighted_losses[tk].to(device=dev)
        if not self.has_initted and self.alpha > 0:
            self.initial_losses.copy_(_allreduce_mean(loss_values.clone()))
            self.has_initted = True
        grad_norms = torch.zeros(K, device=dev)
        for i, tk in enumerate(tasks):
            if tk in grad_tensors:
                gn = grad_tensors[tk].norm(p=2)
                grad_norms[i] = _allreduce_mean(gn.unsqueeze(0)).squeeze()
        g_avg = grad_norms.mean()
        if self.alpha > 0:
            ratio = loss_values / self.initial_losses.to(dev).clamp(min=1e-8)
            ratio_normalized = ratio * (self.num_tasks / ratio.sum().clamp(min=1e-8))
            target = g_avg * (ratio_normalized ** self.alpha)
        else:
            target = g_avg * torch.ones_like(grad_norms)
        w = self.task_weights.to(dev)
        new_w = torch.where(target < 1e-8, w, w * (grad_norms / target))
        new_w = new_w * (K / new_w.sum().clamp(min=1e-8))
        self.task_weights.copy_(new_w)
        if not return_metrics:
            return {}
        metrics = {"gradnorm/avg_norm": g_avg.item()}
        lv, gn, tg, nw = loss_values.tolist(), grad_norms.tolist(), target.tolist(), new_w.tolist()
        for i, tk in enumerate(tasks):
            metrics[f"gradnorm/loss/{tk}"] = lv[i]
            metrics[f"gradnorm/norm/{tk}"] = gn[i]
            metrics[f"gradnorm/target/{tk}"] = tg[i]
            metrics[f"gradnorm/weight/{tk}"] = nw[i]
        if self.alpha > 0:
            rn = (ratio * (K / ratio.sum().clamp(min=1e-8))).tolist()
            for i, tk in enumerate(tasks):
                metrics[f"gradnorm/ratio/{tk}"] = rn[i]
        return metrics


def backbone_parameters(model: nn.Module, exclude_patterns=("head", "meta_")) -> list[nn.Parameter]:
    """The shared trunk GradNorm measures on: every trainable parameter whose name contains none of the patterns - the default
    ``LOSS.GRAD_WEIGHTING.TASK.EXCLUDE_CONFIG`` (R/config.py:498-507: name filters "head" and "meta_", OR-combined)."""
    return [p for n, p in model.named_parameters() if p.requires_grad and not any(pat in n for pat in exclude_patterns)]


def task_gradient_norms(model, images, aux_info, targets: dict, criteria: dict, task_keys: list[str], backbone_params=None,
                        zero_aux_info: bool = True, optimizer=None, dp=None, accum_steps: int = 1) -> tuple[dict, dict]:
    """-> (unweighted_losses {task: 0-dim}, grad_norms {task: 0-dim}), no host sync.

    Per task k (R/loss/gradient_weighting.py:478-760, one sub-batch): loss_k = sum of the criterion's per-sample losses over the
    non-null samples (label != 0, or one-hot[:, 0] <= 0.5) / max(#non-null, 1); norm_k = || d loss_k / d backbone ||_2.  The
    metadata input is zeroed when ``zero_aux_info`` (the reference's default).  ``optimizer`` (FlatAdamW) / ``dp`` (DataParallel)
    are only used to zero the flat gradient buffer and to keep the measurement passes out of the gradient all-reduce.

    ``accum_steps`` > 1 (``GRADNORM_ACCUM_STEPS``, a memory device of the reference): the reference sums, over sub-batches of
    B // accum_steps samples (the last one takes the remainder), the gradients of each sub-batch's own mean loss, and reports
    total loss / total non-null count (:478-800).  By linearity that gradient is the gradient of sum_i loss_i * valid_i /
    max(n_valid(sub-batch of i), 1), so it is measured here with the same single forward - no sub-batch passes."""
    import contextlib

    params = list(backbone_params) if backbone_params is not None else backbone_parameters(model)
    all_params = [p for p in model.parameters() if p.requires_grad]

    def zero_grads():
        if optimizer is not None:
            optimizer.zero_grad()
        else:
            for p in all_params:
                if p.grad is not None:
                    p.grad.zero_()

    was_training = model.training
    model.train()
    aux = torch.zeros_like(aux_info) if (zero_aux_info and aux_info is not None) else aux_info
    losses, norms = {}, {}
    ctx = dp.no_sync() if dp is not None else contextlib.nullcontext()
    with ctx, torch.enable_grad():
        outputs = model(images, aux)
        for i, k in enumerate(task_keys):
            tgt = targets[k]
            valid = (tgt != 0) if tgt.dim() == 1 else (tgt[:, 0] <= 0.5)
            loss_vec = criteria[k](outputs[k], tgt)
            vf = valid.to(loss_vec.dtype)
            partial = (loss_vec * vf).sum() / vf.sum().clamp(min=1.0)
            if accum_steps > 1:
                B = vf.shape[0]
                sb = B // accum_steps
                w = torch.empty_like(vf)
                for s_idx in range(accum_steps):
                    lo, hi = s_idx * sb, ((s_idx + 1) * sb if s_idx < accum_steps - 1 else B)
                    w[lo:hi] = vf[lo:hi] / vf[lo:hi].sum().clamp(min=1.0)
                grad_loss = (loss_vec * w).sum()
            else:
                grad_loss = partial
            zero_grads()
            grad_loss.backward(retain_graph=i + 1 < len(task_keys))
            grads = [p.grad for p in params if p.grad is not None]
            norms[k] = torch.stack(torch._foreach_norm(grads)).norm(2) if grads else torch.zeros((), device=images.device)
            losses[k] = partial.detach()
    zero_grads()
    model.train(was_training)
    return losses, norms


def update_gradnorm_weights(gradnorm: GradNormModule, model, data_batch, criteria: dict, zero_aux_info: bool = True, optimizer=None, dp=None,
                            backbone_params=None, return_metrics: bool = True, accum_steps: int = 1) -> dict:
    """The B200 counterpart of ``GradientWeighting.update_gradnorm_weights_reforward``: measure, then update.
    ``data_batch`` = (images, targets_dict, aux_info, ...)."""
    images, targets, aux = data_batch[0], data_batch[1], data_batch[2]
    losses, norms = task_gradient_norms(model, images, aux, targets, criteria, list(gradnorm.task_keys), backbone_params, zero_aux_info,
                                        optimizer, dp, accum_steps)
    return gradnorm.measure_and_update(losses, norms, return_metrics=return_metrics)


class GradientWeighting(nn.Module):
    """Task weighting for the hierarchical loss, static or GradNorm (R/loss/gradient_weighting.py:178-365 plus the driver :367-923).

    Same constructor arguments and attributes as the reference class (``task_keys``, ``task_weights``, ``class_weights``,
    ``gradnorm``, ``update_interval``, ``zero_aux_info``, ``backbone_params``, ``model``), so
    ``linnaeus_b200.loss.weighted_hierarchical_loss(..., grad_weighting, ...)`` - which reads ``gradnorm.task_weights`` /
    ``task_weights`` / ``class_weights`` on the device, without the reference's per-sample ``.item()`` class-weight loop - and
    ``train.py``'s ``grad_weighting.update_gradnorm_weights_reforward(batch, criteria, ...)`` call work unchanged.
    ``forward`` is the reference's per-task reduction (mean over ``num_valid`` with optional class weights, times the task weight)
    with the class weights gathered by one indexing op instead of a Python loop."""

    def __init__(self, task_keys: list[str], config, task_weighting_type: str = "static", init_weights=None, class_weights=None,
                 use_subset_weights: bool = False, alpha: float = 1.5, label_densities=None, num_classes=None,
                 init_strategy: str = "inverse_density", update_interval: int = 100, exclude_patterns=None, zero_aux_info: bool = True):
        super().__init__()
        self.task_keys = task_keys
        self.config = config
        self.task_weighting_type = task_weighting_type
        try:
            self.zero_aux_info = getattr(config.LOSS.GRAD_WEIGHTING.TASK, "ZERO_AUX_INFO", zero_aux_info)
        except AttributeError:
            self.zero_aux_info = zero_aux_info
        if isinstance(init_weights, dict):
            init_weights = [init_weights.get(k, 1.0) for k in task_keys]
        init_weights = init_weights or [1.0] * len(task_keys)
        self.task_weights = torch.tensor(init_weights, dtype=torch.float32)
        if task_weighting_type == "gradnorm":
            self.gradnorm = GradNormModule(task_keys=task_keys, alpha=alpha, init_weights=torch.tensor(init_weights, dtype=torch.float32),
                                           label_densities=label_densities, num_classes=num_classes, init_strategy=init_strategy, config=config)
            self.update_interval = update_interval
            self.exclude_patterns = exclude_patterns or ["head", "meta_"]
        else:
            self.gradnorm = None
            self.update_interval = 0
            self.exclude_patterns = []
        self.backbone_params = None
        self.model = None
        self.class_weights = class_weights
        self.use_subset_weights = use_subset_weights

    def set_model(self, model: nn.Module) -> None:
        """Identify the shared backbone (R/loss/gradient_weighting.py:265-299; name patterns, the default EXCLUDE_CONFIG)."""
        if self.task_weighting_type == "gradnorm":
            self.model = model
            self.backbone_params = backbone_parameters(model, tuple(self.exclude_patterns))

    def _normalize_weights(self, weights: torch.Tensor) -> torch.Tensor:
        return weights  # identity in the reference as well (:360-365)

    def forward(self, per_task_losses: dict, targets: dict, subset_ids=None, mixed_subset_ids=None, num_valid_samples_per_task=None):
        """-> ({task: weighted mean loss}, {task: weight}) (R/loss/gradient_weighting.py:301-358)."""
        first = next(iter(per_task_losses.values()))
        device, dtype = first.device, first.dtype
        src = self.gradnorm.task_weights if self.gradnorm is not None else self._normalize_weights(self.task_weights)
        norm_w = src.to(device=device, dtype=dtype)
        weighted = {}
        for i, k in enumerate(self.task_keys):
            loss_vec = per_task_losses[k]
            num_valid = loss_vec.size(0) if num_valid_samples_per_task is None else num_valid_samples_per_task.get(k, loss_vec.size(0))
            if self.class_weights and k in self.class_weights:
                cw = self.class_weights[k]
                tgt = targets[k]
                C = tgt.size(1) if tgt.dim() > 1 else (int(max(cw.keys(), default=0)) + 1)
                with torch.no_grad():
                    vec = torch.ones(max(C, 1), dtype=dtype)
                    for idx, w in cw.items():
                        if 0 <= int(idx) < vec.numel():
                            vec[int(idx)] = float(w)
                    vec = vec.to(device)
                    if tgt.dim() == 1:
                        # labels beyond the largest weighted class keep weight 1.0 like dict.get(label, 1.0)
                        inside = tgt < vec.numel()
                        sample_wt = torch.where(inside, vec[tgt.clamp(max=vec.numel() - 1)], torch.ones((), dtype=dtype, device=device))
                    else:
                        sample_wt = (tgt.to(dtype) * vec.unsqueeze(0)).sum(dim=1)
                loss_vec = loss_vec * sample_wt
            weighted[k] = (loss_vec.sum() / max(float(num_valid), 1e-6)) * norm_w[i]
        return weighted, dict(zip(self.task_keys, norm_w.tolist()))

    def update_gradnorm_weights_reforward(self, data_batch, criteria: dict, amp_enabled: bool = True, ops_schedule=None, current_step=None,
                                          optimizer=None, dp=None, return_metrics: bool = True) -> dict:
        """The reference's GradNorm update entry point (:367-923), here one forward + K backward passes (``update_gradnorm_weights``).
        ``amp_enabled`` / ``ops_schedule`` / ``current_step`` are accepted for call compatibility (the model's compute dtype decides
        the precision).  ``GRADNORM_ACCUM_STEPS`` is read from the config when present."""
        if self.task_weighting_type != "gradnorm" or self.gradnorm is None or self.model is None or not self.backbone_params:
            return {}
        try:
            accum = int(self.config.LOSS.GRAD_WEIGHTING.TASK.GRADNORM_ACCUM_STEPS)
        except (AttributeError, KeyError, TypeError):
            accum = 1
        return update_gradnorm_weights(self.gradnorm, self.model, data_batch, criteria, zero_aux_info=bool(self.zero_aux_info), optimizer=optimizer,
                                       dp=dp, backbone_params=self.backbone_params, return_metrics=return_metrics, accum_steps=max(accum, 1))
