"""CPU: host-side logic of linnaeus_b200.aug that needs no kernel - the sync-free in-group permutation, the null exclusion and the
CutMix box - against the oracle / the reference's own helpers."""
import random

import numpy as np
import pytest
import torch

import linnaeus_b200.aug as A
from oracle import aug_oracle as AO
from tests.support import refload


@pytest.mark.parametrize("seed", range(8))
def test_ingroup_permutation_is_valid_and_moves_samples(seed):
    rng = np.random.default_rng(seed)
    B = int(rng.integers(1, 200))
    g = rng.integers(0, max(1, B // 6) + 1, size=B).astype(np.int64)
    g[rng.random(B) < 0.25] = -1
    gt = torch.from_numpy(g)
    torch.manual_seed(seed)
    moved = 0
    for _ in range(5):
        perm = A.ingroup_permutation(gt).numpy()
        assert AO.is_ingroup_permutation(perm, g)
        moved += int((perm != np.arange(B)).sum())
    if any((g == v).sum() > 3 for v in np.unique(g) if v != -1):
        assert moved > 0  # groups with several members do get shuffled


def test_ingroup_permutation_is_uniform_within_a_group():
    g = torch.tensor([5, 5, 5, -1, 7])
    torch.manual_seed(0)
    counts = {}
    for _ in range(3000):
        p = tuple(A.ingroup_permutation(g)[:3].tolist())
        counts[p] = counts.get(p, 0) + 1
    assert len(counts) == 6 and min(counts.values()) > 400  # 3! orderings, ~500 each


def test_exclude_null_samples_matches_oracle():
    rng = np.random.default_rng(3)
    B = 40
    hard = rng.integers(0, 4, size=B).astype(np.int64)
    onehot = np.eye(5, dtype=np.float32)[rng.integers(0, 5, size=B)]
    soft = 0.6 * onehot + 0.4 * np.eye(5, dtype=np.float32)[rng.integers(0, 5, size=B)]
    gids = rng.integers(0, 6, size=B).astype(np.int64)
    targets = {"taxa_L10": hard, "taxa_L20": onehot, "taxa_L30": soft}
    tt = {k: torch.from_numpy(v) for k, v in targets.items()}
    for keys in (None, "taxa_L10", ["taxa_L20", "taxa_L30"], ["missing"]):
        got = A.exclude_null_samples_from_mixup((None, tt, None, None, torch.from_numpy(gids)), keys)[4].numpy()
        assert np.array_equal(got, AO.exclude_null_group_ids(targets, gids, keys))


def test_rand_bbox_matches_oracle_and_reference():
    for seed in range(20):
        lam = (seed + 0.5) / 20
        size = (1, 3, 17 + seed, 40 - seed)
        random.seed(seed)
        box = A.rand_bbox(size, lam)
        random.seed(seed)
        cx, cy = random.randint(0, size[2]), random.randint(0, size[3])
        assert box == AO.rand_bbox_from(size, lam, cx, cy)
        if refload.reference_available():
            refload.import_reference()
            from linnaeus.aug.utils import rand_bbox as ref_bbox

            random.seed(seed)
            assert box == ref_bbox(size, lam)


def test_constructor_mirrors_the_reference_config_handling():
    m = A.GPUSelectiveMixup({"PROB": 0.5, "ALPHA": 0.2, "meta_chunk_bounds_list": [(0, 2), (2, 5)]})
    assert m.chunk_bounds == [(0, 2), (2, 5)]
    assert A.GPUSelectiveMixup({"PROB": 0.5, "ALPHA": 0.2, "meta_chunk_bounds_list": "bad"}).chunk_bounds is None
    c = A.GPUSelectiveCutMix({"MINMAX": [0.2, 0.8]})
    assert c.minmax == [0.2, 0.8] and c.chunk_bounds is None
    with pytest.raises(ValueError):
        A.GPUSelectiveMixup({}, rng="numpy")
    with pytest.raises(RuntimeError):  # no CPU fallback
        m((torch.zeros(2, 3, 4, 4), {}, torch.zeros(2, 5), torch.zeros(2, 5, dtype=torch.bool), torch.zeros(2, dtype=torch.int64)))
