"""Test support: the committed mFormerV0 golden cases (tests/golden/v0_*.npz, made by tests/golden/make_golden_v0.py)."""
import os

import numpy as np

from linnaeus_b200.config import make_synthetic_config_v0

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden")
TINY = dict(conv_embed=(16, 32), conv_out=(32, 64), conv_depths=(1, 2), conv_strides=((2,), (1, 1)), attn_dims=(64, 128), attn_depths=(2, 1),
            heads=(2, 4))
CASES = {
    # name: (img, arch kwargs, batch, weight seed, data seed)  -- must match make_golden_v0.py
    "v0_tiny64": (64, TINY, 3, 0, 0),
    "v0_tiny96": (96, TINY, 2, 1, 1),
    "v0_sm224": (224, {}, 2, 0, 0),
}


def load_case(name):
    img, kw, batch, wseed, dseed = CASES[name]
    cfg, nc = make_synthetic_config_v0("sm", img, **kw)
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return cfg, nc, batch, wseed, dseed, z
