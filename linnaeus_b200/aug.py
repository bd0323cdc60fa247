"""Selective mixup on the device, feeding the model.

Mirror of ``linnaeus.aug.gpu.selective_mixup.GPUSelectiveMixup`` (R/aug/gpu/selective_mixup.py:13-604) and
``linnaeus.aug.utils.exclude_null_samples_from_mixup`` (R/aug/utils.py:46-230): same constructor, same call signature,
same outputs ``(mixed_images, mixed_targets, mixed_aux_info, mixed_meta_masks)``, same in-place "all-or-nothing"
side effect on the caller's ``aux_info`` / ``meta_masks``.

Two layers:

* ``mixup_apply(images, targets, aux, masks, perm, lam, pick, chunk_bounds)`` - the deterministic apply step on the CUDA
  kernels ``lnx_mix_pairs`` / ``lnx_mix_meta_chunks``; bit-equal to the reference for the same ``perm`` / ``lam`` / ``pick``
  (that is the parity surface: tests/test_gpu_aug.py, oracle/aug_oracle.py).
* ``GPUSelectiveMixup.__call__`` - draws ``perm`` / ``lam`` / ``pick``.  ``rng="reference"`` makes exactly the reference's
  torch RNG calls in the reference's order (``rand(1)``, one ``randperm`` per group in ``unique()`` order, the Beta sample,
  ``rand(B)``), so a seeded run reproduces the reference on the same device - at the reference's cost of ~2 + #groups host
  syncs.  ``rng="device"`` (default) never synchronises: the probability gate and the "no group at all" early-out become
  ``lam = 1`` (an exact identity for the blend: 1 * x + 0 * x[perm]; metadata is then taken from the original whenever it
  is non-zero), and the in-group permutation comes from two sorts instead of a Python loop over groups.

The reference's metadata mix is a Python loop with two ``.item()``-style syncs per (sample, chunk); here it is two launches.
"""
from __future__ import annotations

import ctypes
from typing import Any

import torch

from ._lib import call


def _require_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("linnaeus_b200.aug runs on CUDA (sm_100a) only; there is no CPU fallback")


def exclude_null_samples_from_mixup(batch, null_task_keys=None, config=None):
    """group_id := -1 for every sample that is null (label 0, or one-hot[:, 0] > 0.5) in any of ``null_task_keys``
    (default: all tasks).  R/aug/utils.py:46-230, without its per-task ``.item()`` logging syncs."""
    images, targets, aux_info, meta_masks, group_ids = batch
    if null_task_keys is None:
        null_task_keys = list(targets.keys())
    elif isinstance(null_task_keys, str):
        null_task_keys = [null_task_keys]
    null_mask = torch.zeros_like(group_ids, dtype=torch.bool)
    for k in null_task_keys:
        if k not in targets:
            continue
        t = targets[k].to(null_mask.device)
        null_mask |= (t == 0) if t.dim() == 1 else (t[:, 0] > 0.5)
    new_group_ids = torch.where(null_mask, torch.full_like(group_ids, -1), group_ids)
    return images, targets, aux_info, meta_masks, new_group_ids


def ingroup_permutation(group_ids: torch.Tensor, generator: torch.Generator | None = None) -> torch.Tensor:
    """A uniformly random permutation that only moves samples within their group (group -1 and singletons stay put), with no
    host sync: members of a group listed in index order are mapped onto the same members listed in random order
    (R/aug/gpu/selective_mixup.py:326-369 does this with a Python loop over ``unique()`` groups)."""
    B = group_ids.shape[0]
    dev = group_ids.device
    idx = torch.arange(B, device=dev)
    # ungrouped samples become singleton groups of their own (distinct negative ids)
    g = torch.where(group_ids == -1, -2 - idx.to(group_ids.dtype), group_ids)
    by_index = torch.sort(g, stable=True)[1]  # group-major, index order inside a group
    r = torch.rand(B, device=dev, generator=generator)
    by_rand = by_index[torch.sort(g[by_index].to(torch.float64) * 2.0 + r[by_index].to(torch.float64), stable=True)[1]]
    perm = torch.empty_like(idx)
    perm[by_index] = by_rand
    return perm


def mixup_apply(images, targets: dict, aux_info, meta_masks, perm, lam, pick, chunk_bounds):
    """The apply step of selective mixup for given draws.  ``lam``: 0-dim / 1-element float32 DEVICE tensor; ``pick``: float32
    [B]; ``perm``: int64 [B]; ``chunk_bounds``: list of (start, end) or None (= one chunk over all of ``aux_info``).
    ``aux_info`` / ``meta_masks`` (bool) are all-or-nothing enforced IN PLACE like the reference.  Returns the four mixed outputs."""
    _require_cuda(images)
    B = images.shape[0]
    perm = perm.to(device=images.device, dtype=torch.int64).contiguous()
    lam = lam.to(device=images.device, dtype=torch.float32).reshape(1)
    pick = pick.to(device=images.device, dtype=torch.float32).contiguous()

    def blend(x):
        if x.dtype != torch.float32:
            raise TypeError(f"mixup_apply blends float32 tensors, got {x.dtype}")
        xc = x.contiguous()
        out = torch.empty_like(xc)
        call("lnx_mix_pairs", xc.data_ptr(), perm.data_ptr(), lam.data_ptr(), out.data_ptr(), B, xc.numel() // B)
        return out

    mixed_images = blend(images)
    mixed_targets = {k: blend(v) for k, v in targets.items()}
    if aux_info.ndim < 2 or aux_info.shape[1] == 0:
        return mixed_images, mixed_targets, torch.empty_like(aux_info), torch.empty_like(meta_masks)
    D = aux_info.shape[1]
    bounds = list(chunk_bounds) if chunk_bounds is not None else [(0, D)]
    if not (aux_info.is_contiguous() and meta_masks.is_contiguous() and aux_info.dtype == torch.float32 and meta_masks.dtype == torch.bool):
        raise TypeError("aux_info must be contiguous float32 and meta_masks contiguous bool (they are enforced in place)")
    out_aux = torch.zeros_like(aux_info)     # entries outside every chunk: the reference leaves torch.empty_like garbage there
    out_mask = torch.zeros_like(meta_masks)
    if bounds:
        flat = (ctypes.c_int * (2 * len(bounds)))(*[int(v) for b in bounds for v in b])
        call("lnx_mix_meta_chunks", aux_info.data_ptr(), meta_masks.data_ptr(), perm.data_ptr(), pick.data_ptr(), flat, len(bounds),
             out_aux.data_ptr(), out_mask.data_ptr(), B, D)
    return mixed_images, mixed_targets, out_aux, out_mask


class GPUSelectiveMixup:
    """Group-aware pairwise mixup (R/aug/gpu/selective_mixup.py:13-330).  ``mix_config``: ``PROB``, ``ALPHA``,
    ``meta_chunk_bounds_list`` (list of (start, end); absent -> one chunk over the whole aux vector)."""

    def __init__(self, mix_config: dict[str, Any], config=None, rng: str = "device"):
        if rng not in ("device", "reference"):
            raise ValueError("rng must be 'device' or 'reference'")
        self.mix_config = mix_config
        self.config = config
        self.rng = rng
        cb = mix_config.get("meta_chunk_bounds_list") if isinstance(mix_config, dict) else None
        self.chunk_bounds = cb if isinstance(cb, list) else None
        self.last_permutation = None

    # -- the reference's draw sequence, call for call (syncs like the reference) --------------------------------------
    def _reference_permutation(self, group_ids: torch.Tensor) -> torch.Tensor:
        dev = group_ids.device
        perm = torch.arange(group_ids.size(0), device=dev)
        for g in group_ids.unique():
            if g.item() == -1:
                continue
            idx = (group_ids == g).nonzero(as_tuple=True)[0]
            if idx.numel() > 1:
                perm[idx] = idx[torch.randperm(idx.numel(), device=dev)]
        return perm

    def __call__(self, batch, exclude_null_samples: bool = True, null_task_keys=None):
        if exclude_null_samples:
            batch = exclude_null_samples_from_mixup(batch, null_task_keys, config=self.config)
        images, targets, aux_info, meta_masks, group_ids = batch
        _require_cuda(images)
        dev = images.device
        B = images.shape[0]
        alpha = float(self.mix_config["ALPHA"])
        if self.rng == "reference":
            if torch.rand(1, device=dev).item() > self.mix_config["PROB"]:
                return images, targets, aux_info, meta_masks
            if (group_ids == -1).all():
                return images, targets, aux_info, meta_masks
            perm = self._reference_permutation(group_ids)
            lam = torch.distributions.beta.Beta(alpha, alpha).sample().to(dev)
            self.last_permutation = perm
            # the reference enforces all-or-nothing before drawing pick_rand; the order of RNG calls is what matters here
            pick = torch.rand(B, device=dev)
            return mixup_apply(images, targets, aux_info, meta_masks, perm, lam, pick, self.chunk_bounds)
        # sync-free: gate and early-out folded into lam = 1 (exact identity for the blend)
        gate = torch.rand(1, device=dev) > float(self.mix_config["PROB"])
        skip = gate | (group_ids == -1).all().reshape(1)
        perm = ingroup_permutation(group_ids)
        a = torch.full((1,), alpha, device=dev)
        lam = torch.where(skip, torch.ones(1, device=dev), torch.distributions.beta.Beta(a, a).sample().reshape(1).float())
        pick = torch.where(skip.expand(B), torch.zeros(B, device=dev), torch.rand(B, device=dev))
        perm = torch.where(skip.expand(B), torch.arange(B, device=dev), perm)
        self.last_permutation = perm
        # a skipped batch must come back untouched, including the in-place all-or-nothing enforcement of the metadata
        aux0, mask0 = aux_info.clone(), meta_masks.clone()
        mi, mt, ma, mm = mixup_apply(images, targets, aux_info, meta_masks, perm, lam, pick, self.chunk_bounds)
        aux_info.copy_(torch.where(skip, aux0, aux_info))
        meta_masks.copy_(torch.where(skip, mask0, meta_masks))
        return mi, mt, torch.where(skip, aux0, ma), torch.where(skip, mask0, mm)


# ----------------------------------------------------------------------------- selective CutMix
def rand_bbox(size, lam: float):
    """R/aug/utils.py:16-43 (Python ``random`` for the centre; note its W = size[2], H = size[3] naming)."""
    import math
    import random

    W, H = size[2], size[3]
    cut_rat = math.sqrt(1.0 - lam)
    cut_w, cut_h = int(W * cut_rat), int(H * cut_rat)
    cx, cy = random.randint(0, W), random.randint(0, H)
    return max(0, cx - cut_w // 2), max(0, cy - cut_h // 2), min(W, cx + cut_w // 2), min(H, cy + cut_h // 2)


def cutmix_apply(images, targets: dict, aux_info, meta_masks, group_ids, perm, box, pick, chunk_bounds):
    """The apply step of selective CutMix for given draws (R/aug/gpu/selective_cutmix.py:204-437).  ``box`` = (bbx1, bby1, bbx2,
    bby2) HOST ints indexing dims 2 / 3 of ``images`` as the reference does; grouped samples (group id != -1) take their
    partner's pixels inside it and blend soft targets with lam_adjusted = 1 - box_area / (H * W); metadata as in mixup."""
    _require_cuda(images)
    B, C, H, W = images.shape
    dev = images.device
    perm = perm.to(device=dev, dtype=torch.int64).contiguous()
    gids = group_ids.to(device=dev, dtype=torch.int64).contiguous()
    pick = pick.to(device=dev, dtype=torch.float32).contiguous()
    bbx1, bby1, bbx2, bby2 = (int(v) for v in box)
    lam_adjusted = 1.0 - ((bbx2 - bbx1) * (bby2 - bby1)) / (H * W)
    if images.dtype != torch.float32:
        raise TypeError(f"cutmix_apply works on float32 images, got {images.dtype}")
    xc = images.contiguous()
    mixed_images = torch.empty_like(xc)
    call("lnx_cutmix_paste", xc.data_ptr(), perm.data_ptr(), gids.data_ptr(), mixed_images.data_ptr(), B, C, H, W, bbx1, bby1, bbx2, bby2)
    mixed_targets = {}
    for k, v in targets.items():
        if v.dtype != torch.float32:
            raise TypeError(f"cutmix_apply blends float32 targets, got {v.dtype} for {k}")
        vc = v.contiguous()
        out = torch.empty_like(vc)
        call("lnx_mix_pairs_valid", vc.data_ptr(), perm.data_ptr(), gids.data_ptr(), float(lam_adjusted), float(1 - lam_adjusted), out.data_ptr(), B,
             vc.numel() // B)
        mixed_targets[k] = out
    if aux_info.ndim < 2 or aux_info.shape[1] == 0:
        return mixed_images, mixed_targets, torch.empty_like(aux_info), torch.empty_like(meta_masks)
    D = aux_info.shape[1]
    bounds = list(chunk_bounds) if chunk_bounds is not None else [(0, D)]
    if not (aux_info.is_contiguous() and meta_masks.is_contiguous() and aux_info.dtype == torch.float32 and meta_masks.dtype == torch.bool):
        raise TypeError("aux_info must be contiguous float32 and meta_masks contiguous bool (they are enforced in place)")
    out_aux, out_mask = torch.zeros_like(aux_info), torch.zeros_like(meta_masks)
    if bounds:
        flat = (ctypes.c_int * (2 * len(bounds)))(*[int(v) for b in bounds for v in b])
        call("lnx_mix_meta_chunks", aux_info.data_ptr(), meta_masks.data_ptr(), perm.data_ptr(), pick.data_ptr(), flat, len(bounds),
             out_aux.data_ptr(), out_mask.data_ptr(), B, D)
    return mixed_images, mixed_targets, out_aux, out_mask


class GPUSelectiveCutMix(GPUSelectiveMixup):
    """Group-aware CutMix (R/aug/gpu/selective_cutmix.py:14-543): ``PROB`` (default 1.0), ``ALPHA`` (default 1.0), optional
    ``MINMAX`` bounds on lambda, ``meta_chunk_bounds_list``.  lambda is sampled on the host (as the reference's Beta sample is)
    and the box is drawn with Python ``random`` like ``rand_bbox``; ``rng="device"`` additionally takes the probability gate
    from the CPU generator and builds the permutation with two sorts, so nothing waits for the device."""

    def __init__(self, mix_config: dict[str, Any], config=None, rng: str = "device"):
        super().__init__(mix_config, config, rng)
        self.minmax = mix_config.get("MINMAX", None)
        self.last_box = None

    def __call__(self, batch, exclude_null_samples: bool = True, null_task_keys=None):
        if exclude_null_samples:
            batch = exclude_null_samples_from_mixup(batch, null_task_keys, config=self.config)
        images, targets, aux_info, meta_masks, group_ids = batch
        _require_cuda(images)
        dev = images.device
        B, C, H, W = images.shape
        prob = self.mix_config.get("PROB", 1.0)
        alpha = float(self.mix_config.get("ALPHA", 1.0))
        if self.rng == "reference":
            if torch.rand(1, device=dev).item() > prob:
                return images, targets, aux_info, meta_masks
            if (group_ids == -1).all():
                return images, targets, aux_info, meta_masks
            perm = self._reference_permutation(group_ids)
        else:
            if torch.rand(1).item() > prob:  # CPU generator: a host-side decision, no device sync
                return images, targets, aux_info, meta_masks
            perm = ingroup_permutation(group_ids)
        lam = torch.distributions.beta.Beta(alpha, alpha).sample()  # host tensor, as in the reference
        if self.minmax is not None:
            lam = self.minmax[0] + (self.minmax[1] - self.minmax[0]) * lam
        box = rand_bbox((1, C, H, W), lam.item())
        self.last_permutation, self.last_box = perm, box
        pick = torch.rand(B, device=dev)
        if self.rng == "reference":
            return cutmix_apply(images, targets, aux_info, meta_masks, group_ids, perm, box, pick, self.chunk_bounds)
        # "no group at all" cannot be tested without a sync: the image / target kernels are exact no-ops for ungrouped samples,
        # and the metadata (including its in-place enforcement) is restored on the device
        skip = (group_ids == -1).all().reshape(1)
        aux0, mask0 = aux_info.clone(), meta_masks.clone()
        mi, mt, ma, mm = cutmix_apply(images, targets, aux_info, meta_masks, group_ids, perm, box, pick, self.chunk_bounds)
        aux_info.copy_(torch.where(skip, aux0, aux_info))
        meta_masks.copy_(torch.where(skip, mask0, meta_masks))
        return mi, mt, torch.where(skip, aux0, ma), torch.where(skip, mask0, mm)
