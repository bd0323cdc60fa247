"""Generates tests/golden/v0_*.npz by running the UNMODIFIED reference mFormerV0 (/root/reference, CPU, fp32, eval mode)
on seeded synthetic weights / inputs.  Build container only:  python tests/golden/make_golden_v0.py
Inputs and weights are regenerated from seeds by oracle.mformer_v0_oracle (CPU generators); only logits are stored."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import mformer_v0_oracle as V  # noqa: E402
from tests.support import refload  # noqa: E402

TINY = dict(conv_embed=(16, 32), conv_out=(32, 64), conv_depths=(1, 2), conv_strides=((2,), (1, 1)), attn_dims=(64, 128), attn_depths=(2, 1),
            heads=(2, 4))
CASES = {
    # name: (img, arch kwargs, batch, weight seed, data seed)
    "v0_tiny64": (64, TINY, 3, 0, 0),
    "v0_tiny96": (96, TINY, 2, 1, 1),
    "v0_sm224": (224, {}, 2, 0, 0),
}


def main():
    refload.import_reference()
    from linnaeus.models import build_model

    for name, (img, kw, batch, wseed, dseed) in CASES.items():
        cfg, nc = refload.reference_config_v0(img_size=img, **kw)
        model = build_model(cfg, num_classes=nc, taxonomy_tree=None).eval()
        a = V.arch_from_config(cfg, nc)
        P = V.synth_state_dict(a, wseed)
        model.load_state_dict(P)
        x, m = V.synth_batch(a, batch, dseed)
        with torch.no_grad():
            out = model(x, m)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **{"logits." + k: v.numpy() for k, v in out.items()})
        print(name, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
