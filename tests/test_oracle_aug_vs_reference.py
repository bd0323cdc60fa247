"""Build container only: the selective-mixup oracle against the UNMODIFIED reference class on CPU tensors, fresh seeds (skipped
where /root/reference is absent; the committed goldens cover that case)."""
import numpy as np
import pytest
import torch

from oracle import aug_oracle as AO
from tests.support import refload
from tests.support.mixup_replay import replay_draws

pytestmark = pytest.mark.skipif(not refload.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("chunks", [[(0, 2), (2, 5), (5, 15)], None])
def test_aug_oracle_vs_reference(seed, chunks):
    refload.import_reference()
    from linnaeus.aug.gpu.selective_mixup import GPUSelectiveMixup
    from linnaeus.aug.utils import exclude_null_samples_from_mixup

    rng = np.random.default_rng(500 + seed)
    B = [12, 31, 8, 20, 5][seed]
    images = rng.standard_normal((B, 3, 5, 5)).astype(np.float32)
    y = rng.integers(0, 6, size=B)
    hard = rng.integers(0, 4, size=B).astype(np.int64)  # a hard-label task next to a one-hot one
    targets = {"taxa_L10": np.eye(6, dtype=np.float32)[y], "taxa_L20": np.eye(4, dtype=np.float32)[hard]}
    aux = rng.standard_normal((B, 15)).astype(np.float32)
    aux[rng.random((B, 15)) < 0.1] = 0.0
    aux[rng.random(B) < 0.3, 2:5] = 0.0
    masks = aux != 0
    gids = rng.integers(0, 3, size=B).astype(np.int64)
    gids[rng.random(B) < 0.2] = -1
    alpha = 0.8
    cfg = {"PROB": 1.0, "ALPHA": alpha}
    if chunks is not None:
        cfg["meta_chunk_bounds_list"] = chunks
    t = lambda a: torch.from_numpy(a.copy())
    batch = (t(images), {k: t(v) for k, v in targets.items()}, t(aux), t(masks), t(gids))
    eff = exclude_null_samples_from_mixup(batch, "taxa_L10", config=None)[4]
    assert np.array_equal(AO.exclude_null_group_ids(targets, gids, "taxa_L10"), eff.numpy())
    mix = GPUSelectiveMixup(cfg, config=None)
    torch.manual_seed(77 + seed)
    mi, mt, ma, mm = mix(batch, exclude_null_samples=True, null_task_keys="taxa_L10")
    if (eff == -1).all():
        pytest.skip("nothing to mix")
    _, perm, lam, pick = replay_draws(eff, alpha, 77 + seed)
    assert torch.equal(perm, mix.last_permutation)
    a2, m2 = aux.copy(), masks.copy()
    oi, ot, oa, om = AO.mixup_apply(images, targets, a2, m2, perm.numpy(), lam.numpy(), pick.numpy(), chunks)
    assert np.array_equal(oi, mi.numpy())
    for k in targets:
        assert np.array_equal(ot[k], mt[k].numpy())
    assert np.array_equal(oa, ma.numpy()) and np.array_equal(om, mm.numpy())
    assert np.array_equal(a2, batch[2].numpy()) and np.array_equal(m2, batch[3].numpy())


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_cutmix_oracle_vs_reference(seed):
    import random

    refload.import_reference()
    from linnaeus.aug.gpu.selective_cutmix import GPUSelectiveCutMix
    from linnaeus.aug.utils import exclude_null_samples_from_mixup
    from tests.support.mixup_replay import replay_cutmix_draws

    rng = np.random.default_rng(900 + seed)
    B, H, W = [10, 24, 7, 16][seed], [12, 9, 16, 8][seed], [10, 9, 12, 20][seed]
    images = rng.standard_normal((B, 3, H, W)).astype(np.float32)
    targets = {"taxa_L10": np.eye(5, dtype=np.float32)[rng.integers(0, 5, size=B)], "taxa_L20": rng.random((B, 3)).astype(np.float32)}
    aux = rng.standard_normal((B, 15)).astype(np.float32)
    aux[rng.random((B, 15)) < 0.1] = 0.0
    masks = aux != 0
    gids = rng.integers(0, 3, size=B).astype(np.int64)
    gids[rng.random(B) < 0.2] = -1
    chunks = [(0, 2), (2, 5), (5, 15)]
    minmax = [0.2, 0.8] if seed % 2 else None
    cfg = {"PROB": 1.0, "ALPHA": 1.0, "meta_chunk_bounds_list": chunks}
    if minmax:
        cfg["MINMAX"] = minmax
    t = lambda a: torch.from_numpy(a.copy())
    batch = (t(images), {k: t(v) for k, v in targets.items()}, t(aux), t(masks), t(gids))
    eff = exclude_null_samples_from_mixup(batch, "taxa_L10", config=None)[4]
    mix = GPUSelectiveCutMix(cfg, config=None)
    torch.manual_seed(31 + seed)
    random.seed(31 + seed)
    mi, mt, ma, mm = mix(batch, exclude_null_samples=True, null_task_keys="taxa_L10")
    if (eff == -1).all():
        pytest.skip("nothing to mix")
    perm, lam, (cx, cy), pick = replay_cutmix_draws(eff, 1.0, 31 + seed, (1, 3, H, W), minmax)
    box = AO.rand_bbox_from((1, 3, H, W), lam, cx, cy)
    a2, m2 = aux.copy(), masks.copy()
    oi, ot, oa, om = AO.cutmix_apply(images, targets, a2, m2, eff.numpy(), perm.numpy(), box, pick.numpy(), chunks)
    assert np.array_equal(oi, mi.numpy())
    for k in targets:
        assert np.array_equal(ot[k], mt[k].numpy())
    assert np.array_equal(oa, ma.numpy()) and np.array_equal(om, mm.numpy())
    assert np.array_equal(a2, batch[2].numpy()) and np.array_equal(m2, batch[3].numpy())
