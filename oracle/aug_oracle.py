"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the apply step of the reference's selective mixup: for GIVEN draws (in-group
permutation, lambda, one uniform number per sample) it reproduces what ``GPUSelectiveMixup.__call__`` returns and the in-place
side effect on the metadata (R/aug/gpu/selective_mixup.py:140-330, 371-392, 394-560), plus ``exclude_null_samples_from_mixup``
(R/aug/utils.py:46-230).  Pure-Python loops; small cases only.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this file.  Parity pinning:
``tests/test_oracle_aug_vs_reference.py`` runs the unmodified reference class on CPU tensors with a seeded generator and
replays its RNG calls to obtain the same draws; ``tests/golden/make_golden_aug.py`` freezes reference outputs (and the
replayed draws) into ``tests/golden/aug_*.npz`` for the GPU box.  ``R/`` = ``/root/reference/linnaeus``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["exclude_null_group_ids", "enforce_all_or_nothing", "mixup_apply", "cutmix_apply", "rand_bbox_from", "is_ingroup_permutation"]


def exclude_null_group_ids(targets: dict, group_ids: np.ndarray, null_task_keys=None) -> np.ndarray:
    """R/aug/utils.py:87-228: group id -1 for samples whose label is 0 (hard) or whose one-hot[:, 0] > 0.5 in any listed task."""
    keys = list(targets.keys()) if null_task_keys is None else ([null_task_keys] if isinstance(null_task_keys, str) else list(null_task_keys))
    null = np.zeros(group_ids.shape[0], dtype=bool)
    for k in keys:
        if k not in targets:
            continue
        t = np.asarray(targets[k])
        null |= (t == 0) if t.ndim == 1 else (t[:, 0] > 0.5)
    out = group_ids.copy()
    out[null] = -1
    return out


def enforce_all_or_nothing(aux: np.ndarray, mask: np.ndarray, chunk_bounds) -> None:
    """In place (R/aug/gpu/selective_mixup.py:371-392): a chunk with any zero entry is zeroed and marked invalid."""
    for lo, hi in chunk_bounds:
        partial = (aux[:, lo:hi] == 0.0).any(axis=1)
        aux[partial, lo:hi] = 0.0
        mask[partial, lo:hi] = False


def mixup_apply(images, targets: dict, aux, mask, perm, lam, pick, chunk_bounds=None):
    """-> (mixed_images, mixed_targets, mixed_aux, mixed_mask); ``aux`` / ``mask`` are enforced in place like the reference.
    ``lam`` float32 scalar, ``pick`` float32 [B]; fp32 arithmetic in the reference's order: lam * v + (1 - lam) * v[perm]."""
    lam = np.float32(lam)
    oml = np.float32(1.0) - lam
    perm = np.asarray(perm)

    def blend(v):
        v = np.asarray(v, dtype=np.float32)
        return (lam * v).astype(np.float32) + (oml * v[perm]).astype(np.float32)

    mi = blend(images)
    mt = {k: blend(v) for k, v in targets.items()}
    B, D = aux.shape
    bounds = list(chunk_bounds) if chunk_bounds is not None else [(0, D)]
    enforce_all_or_nothing(aux, mask, bounds)
    a2, m2 = aux[perm], mask[perm]
    oa, om = np.zeros_like(aux), np.zeros_like(mask)
    for i in range(B):
        for lo, hi in bounds:
            z1, z2 = bool(np.all(aux[i, lo:hi] == 0.0)), bool(np.all(a2[i, lo:hi] == 0.0))
            if not z1 and not z2:
                src = (aux, mask) if pick[i] < 0.5 else (a2, m2)
            elif not z1:
                src = (aux, mask)
            elif not z2:
                src = (a2, m2)
            else:
                continue
            oa[i, lo:hi], om[i, lo:hi] = src[0][i, lo:hi], src[1][i, lo:hi]
    return mi, mt, oa, om


def _mix_meta(aux, mask, perm, pick, bounds):
    enforce_all_or_nothing(aux, mask, bounds)
    a2, m2 = aux[perm], mask[perm]
    oa, om = np.zeros_like(aux), np.zeros_like(mask)
    for i in range(aux.shape[0]):
        for lo, hi in bounds:
            z1, z2 = bool(np.all(aux[i, lo:hi] == 0.0)), bool(np.all(a2[i, lo:hi] == 0.0))
            if not z1 and not z2:
                src = (aux, mask) if pick[i] < 0.5 else (a2, m2)
            elif not z1:
                src = (aux, mask)
            elif not z2:
                src = (a2, m2)
            else:
                continue
            oa[i, lo:hi], om[i, lo:hi] = src[0][i, lo:hi], src[1][i, lo:hi]
    return oa, om


def rand_bbox_from(size, lam: float, cx: int, cy: int):
    """R/aug/utils.py:16-43 with the two ``random.randint`` draws (centre) given."""
    import math

    W, H = size[2], size[3]
    cut_rat = math.sqrt(1.0 - lam)
    cut_w, cut_h = int(W * cut_rat), int(H * cut_rat)
    return max(0, cx - cut_w // 2), max(0, cy - cut_h // 2), min(W, cx + cut_w // 2), min(H, cy + cut_h // 2)


def cutmix_apply(images, targets: dict, aux, mask, group_ids, perm, box, pick, chunk_bounds=None):
    """R/aug/gpu/selective_cutmix.py:204-437 for given draws: box = (bbx1, bby1, bbx2, bby2) over dims 2 / 3; samples with group
    id != -1 take their partner's pixels inside the box and blend targets with the Python-float lam_adjusted (cast to fp32 the
    way torch casts a Python scalar operand); metadata as in mixup.  ``aux`` / ``mask`` are enforced in place."""
    images = np.asarray(images, dtype=np.float32)
    B, C, H, W = images.shape
    perm = np.asarray(perm)
    x1, y1, x2, y2 = box
    lam_adj = 1.0 - ((x2 - x1) * (y2 - y1)) / (H * W)
    ca, cb = np.float32(lam_adj), np.float32(1 - lam_adj)
    valid = np.nonzero(np.asarray(group_ids) != -1)[0]
    mi = images.copy()
    mt = {k: np.asarray(v, dtype=np.float32).copy() for k, v in targets.items()}
    if len(valid):
        mi[valid, :, x1:x2, y1:y2] = images[perm[valid], :, x1:x2, y1:y2]
        for k, v in targets.items():
            v = np.asarray(v, dtype=np.float32)
            mt[k][valid] = (ca * v[valid]).astype(np.float32) + (cb * v[perm[valid]]).astype(np.float32)
    D = aux.shape[1]
    oa, om = _mix_meta(aux, mask, perm, pick, list(chunk_bounds) if chunk_bounds is not None else [(0, D)])
    return mi, mt, oa, om


def is_ingroup_permutation(perm: np.ndarray, group_ids: np.ndarray) -> bool:
    """perm is a bijection that keeps every sample inside its group and fixes group -1 (selective_mixup.py:326-369)."""
    perm = np.asarray(perm)
    if sorted(perm.tolist()) != list(range(len(perm))):
        return False
    g = np.asarray(group_ids)
    return bool(np.all(g[perm] == g) and np.all(perm[g == -1] == np.nonzero(g == -1)[0]))
