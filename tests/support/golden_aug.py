"""Test support: seeded selective-mixup cases shared by tests/golden/make_golden_aug.py (reference side, build container), the
oracle tests and the GPU tests.  Inputs come from numpy seeds; the reference's outputs AND the replayed draws (perm, lam, pick)
are stored in tests/golden/aug_*.npz."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden")
CHUNKS3 = [(0, 2), (2, 5), (5, 15)]  # TEMPORAL(2), SPATIAL(3), ELEVATION(10): SURVEY.md 8(d)
CASES = {
    # name: (B, image side, classes per task, chunk bounds or None, seed, number of groups, torch seed)
    "aug_three_chunks": (24, 8, (7, 4), CHUNKS3, 0, 3, 11),
    "aug_single_chunk": (9, 6, (5,), None, 1, 2, 12),
    "aug_pairs_only": (16, 4, (3, 3, 2), [(0, 2), (2, 5)], 2, 3, 13),
}
KEYS = ["taxa_L10", "taxa_L20", "taxa_L30"]


def make_case(name):
    """-> dict(images f32 [B,3,S,S], targets {key: one-hot f32 [B,C]}, aux f32 [B,15], masks bool [B,15], group_ids int64 [B], chunks)"""
    B, S, classes, chunks, seed, n_groups, _ = CASES[name]
    rng = np.random.default_rng(seed)
    images = rng.standard_normal((B, 3, S, S)).astype(np.float32)
    targets = {}
    for k, C in zip(KEYS, classes):
        y = rng.integers(0, C, size=B)
        y[rng.random(B) < 0.1] = 0  # nulls: excluded from mixing
        targets[k] = np.eye(C, dtype=np.float32)[y]
    aux = rng.standard_normal((B, 15)).astype(np.float32)
    # whole chunks missing (zeros), a few partially-zero chunks that all-or-nothing must wipe
    for lo, hi in CHUNKS3:
        aux[rng.random(B) < 0.3, lo:hi] = 0.0
    aux[rng.random((B, 15)) < 0.04] = 0.0
    masks = aux != 0.0
    group_ids = rng.integers(0, n_groups, size=B).astype(np.int64)
    group_ids[rng.random(B) < 0.15] = -1
    return dict(images=images, targets=targets, aux=aux, masks=masks, group_ids=group_ids, chunks=chunks)


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def covered_columns(chunks, D):
    """Boolean [D]: columns inside some chunk (the reference leaves torch.empty_like garbage everywhere else)."""
    cov = np.zeros(D, dtype=bool)
    for lo, hi in (chunks if chunks is not None else [(0, D)]):
        cov[lo:hi] = True
    return cov
