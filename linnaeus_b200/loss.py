"""Hierarchical masked loss behind the reference's loss surface.

Drop-in for ``linnaeus.loss.hierarchical_loss.weighted_hierarchical_loss``
(R/loss/hierarchical_loss.py:24-406): same call signature, same returned triple
``(total, loss_components, task_weights)`` with the same ``loss_components`` keys.
All K tasks are evaluated by one fused CUDA kernel on the concatenated logits; values
in ``loss_components`` are 0-dim device tensors (``float(x)`` works and is the only
host sync, taken by the caller if and when it logs).

Criteria are declared with the small marker modules below (same constructor arguments
as R/loss/basic_loss.py:15-228 and R/loss/taxonomy_label_smoothing.py:131-217); the
reference's own criterion instances are recognised by class name.
"""
from __future__ import annotations

from typing import Any

import torch
import torch.nn as nn

from . import functional as F
from ._lib import LOSS_CE, LOSS_LS, LOSS_TAXONOMY

__all__ = [
    "CrossEntropyLoss",
    "LabelSmoothingCrossEntropy",
    "TaxonomyAwareLabelSmoothingCE",
    "SoftTargetCrossEntropy",
    "StaticTaskWeighting",
    "weighted_hierarchical_loss",
    "install_into_linnaeus_loss",
]


LOSS_SOFT = 3  # host-side tag: runs the kernel's soft-label-row mode (LNX_LOSS_TAXONOMY) with the targets themselves as the rows


class _Criterion(nn.Module):
    kind = LOSS_CE
    smoothing = 0.0

    def forward(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """Per-sample losses [B] of one task through the fused kernel (K = 1), differentiable."""
        if self.kind == LOSS_SOFT:
            rows = target.float().contiguous()
            loss = F.per_sample_loss(logits, torch.arange(rows.shape[0], device=logits.device), LOSS_TAXONOMY, 0.0, [rows], False)
        else:
            mats = [self.soft_labels] if self.kind == LOSS_TAXONOMY else None
            loss = F.per_sample_loss(logits, _hard(target), self.kind, self.smoothing, mats, getattr(self, "ignore_index", None) == 0)
        sw = _criterion_sample_weight(self, target, logits.device)
        return loss if sw is None else loss * sw


class CrossEntropyLoss(_Criterion):
    kind = LOSS_CE

    def __init__(self, weight=None, apply_class_weights: bool = False, ignore_index: int | None = None):
        super().__init__()
        self.weight, self.apply_class_weights, self.ignore_index = weight, apply_class_weights, ignore_index


class LabelSmoothingCrossEntropy(_Criterion):
    kind = LOSS_LS

    def __init__(self, weight=None, smoothing: float = 0.1, apply_class_weights: bool = False, ignore_index: int | None = None, config=None):
        super().__init__()
        assert 0.0 <= smoothing < 1.0
        self.smoothing = smoothing
        self.weight, self.apply_class_weights, self.ignore_index = weight, apply_class_weights, ignore_index


class TaxonomyAwareLabelSmoothingCE(_Criterion):
    kind = LOSS_TAXONOMY

    def __init__(self, soft_label_matrix: torch.Tensor, weight=None, apply_class_weights: bool = False, ignore_index: int | None = None,
                 config=None):
        super().__init__()
        if soft_label_matrix.dim() != 2 or soft_label_matrix.shape[0] != soft_label_matrix.shape[1]:
            raise ValueError("soft_label_matrix must be square [C, C].")
        self.num_classes = soft_label_matrix.shape[0]
        self.register_buffer("soft_labels", soft_label_matrix.clone().float().contiguous())
        self.weight, self.apply_class_weights, self.ignore_index = weight, apply_class_weights, ignore_index


class SoftTargetCrossEntropy(_Criterion):
    """-sum_c t_c log softmax(z)_c for [B, C] soft targets (mixup / CutMix), R/loss/basic_loss.py:188-228."""

    kind = LOSS_SOFT

    def __init__(self, weight=None, apply_class_weights: bool = False):
        super().__init__()
        self.weight, self.apply_class_weights, self.ignore_index = weight, apply_class_weights, None


def _criterion_sample_weight(crit, target: torch.Tensor, dev) -> torch.Tensor | None:
    """Criterion-level class weights (``weight=`` with ``apply_class_weights=True``; basic_loss.py:76-90,160-175,217-221):
    weight[y] for hard / one-hot targets, sum_c t_c weight_c for SoftTargetCrossEntropy."""
    w = getattr(crit, "weight", None)
    if w is None or not getattr(crit, "apply_class_weights", False):
        return None
    w = w.to(device=dev, dtype=torch.float32)
    kind = _KIND_BY_NAME.get(type(crit).__name__)
    if kind == LOSS_SOFT:
        return (target.to(dev).float() * w[None]).sum(1)
    return w[_hard(target).to(dev)]


class StaticTaskWeighting:
    """Minimal stand-in for GradientWeighting(type='static') (gradient_weighting.py:171-365)."""

    def __init__(self, task_keys: list[str], init_weights=None, class_weights=None):
        self.task_keys = list(task_keys)
        if isinstance(init_weights, dict):
            init_weights = [init_weights.get(k, 1.0) for k in task_keys]
        self.task_weights = torch.tensor(init_weights or [1.0] * len(task_keys), dtype=torch.float32)
        self.class_weights = class_weights
        self.gradnorm = None


_KIND_BY_NAME = {
    "CrossEntropyLoss": LOSS_CE,
    "LabelSmoothingCrossEntropy": LOSS_LS,
    "TaxonomyAwareLabelSmoothingCE": LOSS_TAXONOMY,
    "SoftTargetCrossEntropy": LOSS_SOFT,
}


class _LazyRowMeans(dict):
    """``{task: mean of row k of a [K, B] device tensor}`` evaluated on first access: the reference computes these for its log line
    (hierarchical_loss.py:361-372); computing them eagerly costs 2 K tiny reductions in every training step."""

    def __init__(self, mat: torch.Tensor, keys: list[str]):
        super().__init__()
        self._mat, self._keys = mat, list(keys)

    def _fill(self):
        if self._mat is not None:
            m = self._mat.mean(dim=1)
            for i, k in enumerate(self._keys):
                dict.__setitem__(self, k, m[i])
            self._mat = None

    def __getitem__(self, k):
        self._fill()
        return dict.__getitem__(self, k)

    def __iter__(self):
        return iter(self._keys)

    def __len__(self):
        return len(self._keys)

    def __contains__(self, k):
        return k in self._keys

    def keys(self):
        return list(self._keys)

    def items(self):
        self._fill()
        return dict.items(self)

    def values(self):
        self._fill()
        return dict.values(self)

    def get(self, k, default=None):
        return self[k] if k in self._keys else default


def _hard(t: torch.Tensor) -> torch.Tensor:
    return t.argmax(dim=1) if t.dim() == 2 else t.long()


def _sorted_keys(outputs) -> list[str]:
    return sorted(outputs.keys(), key=lambda k: int(k.split("_L")[-1]))  # core_loss.py:46


def _class_weight_vec(cw: dict, C: int, device) -> torch.Tensor:
    v = torch.ones(C, dtype=torch.float32)
    for idx, w in cw.items():
        if 0 <= int(idx) < C:
            v[int(idx)] = float(w)
    return v.to(device)


def _task_weight_vector(task_weighting, keys: list[str], dev) -> torch.Tensor | None:
    """float32 [K] task weights on ``dev`` in the order of ``keys`` (gradient_weighting.py:317-323: the GradNorm buffer when GradNorm
    is active, else the static vector).  Static weights are uploaded once and cached on the weighting object (keeps the step
    CUDA-graph capturable); GradNorm weights are read from the live buffer every call, on the device, with no host round trip."""
    gn = getattr(task_weighting, "gradnorm", None)
    w_src = gn.task_weights if gn is not None else getattr(task_weighting, "task_weights", None)
    if w_src is None:
        return None
    order = list(getattr(task_weighting, "task_keys", keys))
    idx = [order.index(k) for k in keys]
    ck = (str(dev), tuple(keys))
    if gn is not None:
        w = w_src.detach().to(device=dev, dtype=torch.float32)
        if idx == list(range(len(order))):
            return w
        cache = getattr(task_weighting, "_lnx_idx_cache", None)
        if cache is None:
            cache = {}
            try:
                task_weighting._lnx_idx_cache = cache
            except Exception:
                pass
        if ck not in cache:
            cache[ck] = torch.tensor(idx, dtype=torch.int64, device=dev)
        return w[cache[ck]]
    cache = getattr(task_weighting, "_lnx_tw_cache", None)
    if cache is not None and ck in cache:
        return cache[ck]
    w_cpu = w_src.detach().float().cpu()
    tw = torch.tensor([float(w_cpu[i]) for i in idx], dtype=torch.float32, device=dev)
    try:
        if cache is None:
            task_weighting._lnx_tw_cache = {}
        task_weighting._lnx_tw_cache[ck] = tw
    except Exception:
        pass
    return tw


def weighted_hierarchical_loss(
    outputs: dict[str, torch.Tensor],
    targets: dict[str, torch.Tensor],
    criteria: dict[str, nn.Module],
    task_weighting: Any,
    ops_schedule: Any,
    current_step: int,
    subset_ids: torch.Tensor | None = None,
    mixed_subset_ids: torch.Tensor | None = None,
    is_validation: bool = False,
    logger=None,
    config=None,
):
    keys = _sorted_keys(outputs)
    if not isinstance(targets, dict):
        targets = dict(zip(keys, targets))
    K = len(keys)
    kinds = {_KIND_BY_NAME.get(type(criteria[k]).__name__) for k in keys}
    if None in kinds or len(kinds) != 1:
        raise NotImplementedError(
            "linnaeus_b200 fuses CrossEntropyLoss / LabelSmoothingCrossEntropy / TaxonomyAwareLabelSmoothingCE / SoftTargetCrossEntropy, "
            f"one kind for all tasks; got {[type(criteria[k]).__name__ for k in keys]}"
        )
    kind = kinds.pop()
    smoothing = float(getattr(criteria[keys[0]], "smoothing", 0.0))
    # PHASE1 training branch (hierarchical_loss.py:241-276): null losses zeroed, divisor = batch size.  Independently of it the
    # criteria may have been built with ignore_index = 0 (loss/utils.py:104-145 does so for train AND validation criteria under
    # TRAIN.PHASE1_MASK_NULL_LOSS): null losses are zero at the source and drop out of n_valid = #(loss != 0).
    phase1 = bool(config is not None and getattr(config.TRAIN, "PHASE1_MASK_NULL_LOSS", False) and not is_validation)
    ign0 = [getattr(criteria[k], "ignore_index", None) == 0 for k in keys]
    zero_null = phase1 or all(ign0)

    # concatenated logits: reuse the model's single head-GEMM output when the dict carries it
    cat = getattr(outputs, "cat", None)
    if cat is not None and list(outputs.keys()) == keys:
        class_off = tuple(outputs.class_off)
    else:
        cat = torch.cat([outputs[k].float() for k in keys], dim=1)
        class_off = [0]
        for k in keys:
            class_off.append(class_off[-1] + outputs[k].shape[1])
        class_off = tuple(class_off)
    dev = cat.device
    B = cat.shape[0]

    if kind == LOSS_SOFT:
        if any(targets[k].dim() != 2 for k in keys):
            raise ValueError("SoftTargetCrossEntropy needs [B, C] targets")
        tg = torch.arange(B, device=dev).repeat(K, 1).contiguous()  # row index into the per-task target matrix
    else:
        tg = torch.stack([_hard(targets[k]) for k in keys]).to(dev).contiguous()
    null_flag = None
    if any(targets[k].dim() == 2 for k in keys):
        null_flag = torch.stack([(targets[k][:, 0] > 0.5) if targets[k].dim() == 2 else (targets[k] == 0) for k in keys])
        null_flag = null_flag.to(device=dev, dtype=torch.uint8).contiguous()

    # null masking probability (masking.py:548-564)
    if is_validation:
        p = 1.0
    elif phase1:
        p = 0.0
    else:
        p = float(ops_schedule.get_null_mask_prob(current_step)) if ops_schedule is not None else 1.0
    keep = None
    is_null = None
    if ((not phase1) and p < 1.0) or (any(ign0) and not zero_null):
        is_null = null_flag.bool() if null_flag is not None else (tg == 0)
    if (not phase1) and p < 1.0:
        coin = torch.rand((K, B), device=dev) < p
        keep = torch.where(is_null & ~coin, 0.0, 1.0).float()
    if any(ign0) and not zero_null:  # only some criteria ignore the null class: zero those tasks' null rows through the multiplier
        rows = torch.tensor(ign0, device=dev)[:, None]
        m = torch.where(is_null & rows, 0.0, 1.0).float()
        keep = m if keep is None else keep * m
    # per-sample multipliers from class weights.  The reference applies the GradientWeighting class-weight lookup repeatedly
    # (SURVEY section 0): inside apply_loss_masking (masking.py:696-698; not on the PHASE1 training branch), in
    # hierarchical_loss.py:313-334 when LOSS.GRAD_WEIGHTING.CLASS.TRAIN / .VAL says so, and in GradientWeighting.forward
    # (gradient_weighting.py:334-352).  Soft [B, C] targets weigh in as sum_c t_c w_c each time (masking.py:505-515).
    cw = getattr(task_weighting, "class_weights", None)
    if cw:
        try:
            apply_cw = bool(config.LOSS.GRAD_WEIGHTING.CLASS.VAL if is_validation else config.LOSS.GRAD_WEIGHTING.CLASS.TRAIN)
        except Exception:
            apply_cw = True  # hierarchical_loss.py:322-325
        times = (0 if phase1 else 1) + (1 if apply_cw else 0) + 1
        mult = torch.ones((K, B), device=dev)
        for i, k in enumerate(keys):
            if k in cw:
                C = class_off[i + 1] - class_off[i]
                vec = _class_weight_vec(cw[k], C, dev)
                t_k = targets[k].to(dev)
                sw = (t_k.float() * vec[None]).sum(1) if t_k.dim() == 2 else vec[t_k.long()]
                mult[i] = sw.pow(times)
        keep = mult if keep is None else keep * mult
    # criterion-level class weights (weight= with apply_class_weights=True): once, at the source
    for i, k in enumerate(keys):
        sw = _criterion_sample_weight(criteria[k], targets[k], dev)
        if sw is not None:
            if keep is None:
                keep = torch.ones((K, B), device=dev)
            keep[i] = keep[i] * sw

    tw = _task_weight_vector(task_weighting, keys, dev)

    soft = None
    kernel_kind = kind
    if kind == LOSS_TAXONOMY:
        soft = [criteria[k].soft_labels.to(dev).float().contiguous() for k in keys]
    elif kind == LOSS_SOFT:
        soft = [targets[k].to(dev).float().contiguous() for k in keys]
        kernel_kind = LOSS_TAXONOMY

    stats: dict = {}
    total = F.hier_loss(cat, tg, class_off, kernel_kind, smoothing, soft, tw, keep, null_flag, zero_null, stats, count_all=phase1)

    raw = stats["raw"]
    per = stats["per_sample"]
    loss_components = {
        "total": total.detach(),
        "tasks": _LazyRowMeans(raw, keys),          # per-task means, reduced only when somebody reads them (logging)
        "masked_tasks": _LazyRowMeans(per, keys),
        "weighted_tasks": {k: stats["task_sum"][i] for i, k in enumerate(keys)},
        "raw_per_sample_losses": {k: raw[i] for i, k in enumerate(keys)},
        "null_masking": {
            "null_mask_prob": 0.0 if phase1 else p,
            "num_valid_samples_per_task": {k: stats["nvalid"][i] for i, k in enumerate(keys)},
            "phase1_active": phase1,
            "null_samples_total": 0,
            "null_samples_included": 0,
            "inclusion_percentage": 0.0,
        },
    }
    task_weights = {k: (tw[i] if tw is not None else 1.0) for i, k in enumerate(keys)}
    return total, loss_components, task_weights


def install_into_linnaeus_loss() -> None:
    """Rebind the reference's call sites to the fused loss (loss boundary)."""
    import linnaeus.train as lt
    import linnaeus.validation as lv

    lt.weighted_hierarchical_loss = weighted_hierarchical_loss
    lv.weighted_hierarchical_loss = weighted_hierarchical_loss
    try:
        import linnaeus.utils.autobatch as la

        la.weighted_hierarchical_loss = weighted_hierarchical_loss
    except Exception:  # pragma: no cover
        pass
