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


def main():
    refload.import_reference()
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
