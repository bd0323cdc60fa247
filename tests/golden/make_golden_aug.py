"""Generates tests/golden/aug_*.npz with the UNMODIFIED reference linnaeus.aug.gpu.selective_mixup.GPUSelectiveMixup
(/root/reference, CPU tensors, seeded torch generator) and the draws it used (replayed: tests/support/mixup_replay.py).
Build container only:  python tests/golden/make_golden_aug.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.support import refload  # noqa: E402
from tests.support.golden_aug import CASES, make_case  # noqa: E402
from tests.support.mixup_replay import replay_draws  # noqa: E402

ALPHA = 0.4


def run_reference(name):
    from linnaeus.aug.gpu.selective_mixup import GPUSelectiveMixup
    from linnaeus.aug.utils import exclude_null_samples_from_mixup

    c = make_case(name)
    tseed = CASES[name][-1]
    cfg = {"PROB": 1.0, "ALPHA": ALPHA}
    if c["chunks"] is not None:
        cfg["meta_chunk_bounds_list"] = list(c["chunks"])
    mix = GPUSelectiveMixup(cfg, config=None)
    images = torch.from_numpy(c["images"])
    targets = {k: torch.from_numpy(v) for k, v in c["targets"].items()}
    aux, masks = torch.from_numpy(c["aux"].copy()), torch.from_numpy(c["masks"].copy())
    gids = torch.from_numpy(c["group_ids"])
    eff_gids = exclude_null_samples_from_mixup((images, targets, aux, masks, gids), None, config=None)[4]
    torch.manual_seed(tseed)
    mi, mt, ma, mm = mix((images, targets, aux, masks, gids), exclude_null_samples=True, null_task_keys=None)
    gate, perm, lam, pick = replay_draws(eff_gids, ALPHA, tseed)
    assert torch.equal(perm, mix.last_permutation), "RNG replay out of step with the reference"
    return c, dict(mixed_images=mi, mixed_targets=mt, mixed_aux=ma, mixed_masks=mm, aux_after=aux, masks_after=masks, eff_gids=eff_gids,
                   perm=perm, lam=lam, pick=pick)


def run_reference_cutmix(name, tseed):
    import random

    from linnaeus.aug.gpu.selective_cutmix import GPUSelectiveCutMix
    from linnaeus.aug.utils import exclude_null_samples_from_mixup
    from oracle import aug_oracle as AO
    from tests.support.mixup_replay import replay_cutmix_draws

    c = make_case(name)
    cfg = {"PROB": 1.0, "ALPHA": 1.0, "MINMAX": [0.3, 0.7]}
    if c["chunks"] is not None:
        cfg["meta_chunk_bounds_list"] = list(c["chunks"])
    mix = GPUSelectiveCutMix(cfg, config=None)
    images = torch.from_numpy(c["images"])
    targets = {k: torch.from_numpy(v) for k, v in c["targets"].items()}
    aux, masks = torch.from_numpy(c["aux"].copy()), torch.from_numpy(c["masks"].copy())
    gids = torch.from_numpy(c["group_ids"])
    eff = exclude_null_samples_from_mixup((images, targets, aux, masks, gids), None, config=None)[4]
    torch.manual_seed(tseed)
    random.seed(tseed)
    mi, mt, ma, mm = mix((images, targets, aux, masks, gids))
    perm, lam, (cx, cy), pick = replay_cutmix_draws(eff, 1.0, tseed, tuple(images.shape), [0.3, 0.7])
    box = AO.rand_bbox_from(tuple(images.shape), lam, cx, cy)
    rec = {"mixed_images": mi.numpy(), "mixed_aux": ma.numpy(), "mixed_masks": mm.numpy(), "aux_after": aux.numpy(), "masks_after": masks.numpy(),
           "eff_gids": eff.numpy(), "perm": perm.numpy(), "box": np.array(box), "pick": pick.numpy()}
    for k, v in mt.items():
        rec["mixed_targets." + k] = v.numpy()
    # the replayed draws must reproduce the reference through the oracle, or the golden is useless
    a2, m2 = c["aux"].copy(), c["masks"].copy()
    oi, _, oa, _ = AO.cutmix_apply(c["images"], c["targets"], a2, m2, eff.numpy(), perm.numpy(), box, pick.numpy(), c["chunks"])
    assert np.array_equal(oi, mi.numpy()), "RNG replay out of step with the reference (cutmix)"
    return rec


def main():
    refload.import_reference()
    for name, tseed in (("aug_three_chunks", 21), ("aug_single_chunk", 22)):
        rec = run_reference_cutmix(name, tseed)
        np.savez_compressed(os.path.join(HERE, name.replace("aug_", "cutmix_") + ".npz"), **rec)
        print("cutmix", name, "box", rec["box"], "moved", int((rec["perm"] != np.arange(len(rec["perm"]))).sum()))
    for name in CASES:
        c, r = run_reference(name)
        rec = {"mixed_images": r["mixed_images"].numpy(), "mixed_aux": r["mixed_aux"].numpy(), "mixed_masks": r["mixed_masks"].numpy(),
               "aux_after": r["aux_after"].numpy(), "masks_after": r["masks_after"].numpy(), "eff_gids": r["eff_gids"].numpy(),
               "perm": r["perm"].numpy(), "lam": r["lam"].numpy(), "pick": r["pick"].numpy()}
        for k, v in r["mixed_targets"].items():
            rec["mixed_targets." + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
        moved = int((r["perm"] != torch.arange(len(r["perm"]))).sum())
        print(name, "lam", float(r["lam"]), "moved", moved, "of", len(r["perm"]))


if __name__ == "__main__":
    main()
