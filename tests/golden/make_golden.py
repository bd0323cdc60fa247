"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, CPU, fp32, standard attention path) on seeded synthetic inputs.

Run in the build container only:  python tests/golden/make_golden.py
The fixtures let the GPU box (which has no reference tree) check both the CPU
oracle and the CUDA path against outputs of the reference itself.

Per case the file holds: logits per task, the scalar loss, and for every
parameter gradient its L2 norm plus its first 32 flattened values.
Inputs/weights are NOT stored: they are regenerated from seeds by
oracle.mformer_oracle.synth_state_dict / synth_batch (CPU generators).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import mformer_oracle as O  # noqa: E402
from tests.support import refload  # noqa: E402

CASES = {
    # name: (reference_config kwargs, batch, weight seed, data seed, loss kind)
    "tiny_ce": (dict(variant="sm", img_size=64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(2, 1), conv_depths=(1, 1, 1, 1)), 4, 0, 0, "ce"),
    "tiny_taxonomy": (dict(variant="sm", img_size=64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(2, 1), conv_depths=(1, 1, 1, 1)), 4, 1, 1, "taxonomy"),
    "tiny_nometa": (dict(variant="sm", img_size=64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(2, 1), conv_depths=(1, 1, 1, 1), meta=False), 3, 2, 2, "ce"),
    "sm224_ce": (dict(variant="sm", img_size=224), 2, 0, 0, "ce"),
    "md224_ce": (dict(variant="md", img_size=224), 1, 0, 0, "ce"),
    # round 2: hierarchical head types on the CUDA path, a bench-like batch, and the long-sequence attention path (N = 580 / 148)
    "tiny_hsm": (dict(variant="sm", img_size=64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(2, 1), conv_depths=(1, 1, 1, 1),
                      head_type="HierarchicalSoftmax"), 4, 3, 3, "ce"),
    "tiny_cond": (dict(variant="sm", img_size=64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(2, 1), conv_depths=(1, 1, 1, 1),
                       head_type="ConditionalClassifier"), 4, 4, 4, "ce"),
    "sm224_b32": (dict(variant="sm", img_size=224), 32, 5, 5, "ce"),
    "xl384_shallow": (dict(variant="xl", img_size=384, rope_depths=(1, 1), conv_depths=(1, 1, 1, 1)), 1, 6, 6, "ce"),
}


def run_case(name):
    kw, batch, wseed, dseed, kind = CASES[name]
    refload.import_reference()
    from linnaeus.loss.basic_loss import CrossEntropyLoss
    from linnaeus.loss.gradient_weighting import GradientWeighting
    from linnaeus.loss.hierarchical_loss import weighted_hierarchical_loss
    from linnaeus.loss.taxonomy_label_smoothing import TaxonomyAwareLabelSmoothingCE
    from linnaeus.models import build_model

    cfg, nc = refload.reference_config(**kw)
    tree = refload.synthetic_taxonomy_tree(nc) if kw.get("head_type", "Linear") != "Linear" else None
    model = build_model(cfg, num_classes=nc, taxonomy_tree=tree)
    a = O.arch_from_config(cfg, nc)
    P = O.synth_state_dict(O.param_shapes(a), wseed)
    missing, unexpected = model.load_state_dict(P, strict=False)
    assert not unexpected and all("hmatrix_" in k for k in missing), (missing, unexpected)  # buffers come from the taxonomy tree
    x, meta, tg = O.synth_batch(a, batch, dseed)
    model.train()
    logits = model(x, meta)
    keys = list(logits.keys())
    if kind == "ce":
        crit = {k: CrossEntropyLoss() for k in keys}
    else:
        mats = O.synthetic_taxonomy_smoothing(a.tasks)
        crit = {k: TaxonomyAwareLabelSmoothingCE(mats[k]) for k in keys}
    gw = GradientWeighting(keys, cfg, "static")

    class Sched:
        def get_null_mask_prob(self, step):
            return 1.0

    total, comps, _ = weighted_hierarchical_loss(logits, tg, crit, gw, Sched(), 0, config=cfg)
    total.backward()
    out = {"loss": np.float32(total.item()), "batch": np.int64(batch), "wseed": np.int64(wseed), "dseed": np.int64(dseed)}
    for k in keys:
        out[f"logits/{k}"] = logits[k].detach().numpy()
    for n, p in model.named_parameters():
        g = p.grad.detach().flatten()
        out[f"gnorm/{n}"] = np.float32(g.norm().item())
        out[f"ghead/{n}"] = g[:32].numpy().copy()
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "loss", total.item(), "params", len(list(model.parameters())))


if __name__ == "__main__":
    for c in (sys.argv[1:] or CASES):
        run_case(c)
