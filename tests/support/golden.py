"""Golden-fixture helpers shared by the CPU (oracle) and GPU (CUDA path) tests."""
import os

import numpy as np

from linnaeus_b200.config import make_synthetic_config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden")
_TINY = dict(dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(2, 1), conv_depths=(1, 1, 1, 1))

# must mirror tests/golden/make_golden.py::CASES
CASES = {
    "tiny_ce": (dict(variant="sm", img_size=64, **_TINY), "ce"),
    "tiny_taxonomy": (dict(variant="sm", img_size=64, **_TINY), "taxonomy"),
    "tiny_nometa": (dict(variant="sm", img_size=64, meta=False, **_TINY), "ce"),
    "sm224_ce": (dict(variant="sm", img_size=224), "ce"),
    "md224_ce": (dict(variant="md", img_size=224), "ce"),
    "tiny_hsm": (dict(variant="sm", img_size=64, head_type="HierarchicalSoftmax", **_TINY), "ce"),
    "tiny_cond": (dict(variant="sm", img_size=64, head_type="ConditionalClassifier", **_TINY), "ce"),
    "sm224_b32": (dict(variant="sm", img_size=224), "ce"),
    "xl384_shallow": (dict(variant="xl", img_size=384, rope_depths=(1, 1), conv_depths=(1, 1, 1, 1)), "ce"),
}


def load_case(name):
    """-> (cfg, num_classes, loss kind, npz dict)"""
    kw, kind = CASES[name]
    cfg, nc = make_synthetic_config(**kw)
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    return cfg, nc, kind, z
