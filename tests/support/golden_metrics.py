"""Test support: seeded metric cases shared by tests/golden/make_golden_metrics.py (reference side, build container) and the
oracle / GPU tests.  Inputs are regenerated from seeds (numpy Generator: the same stream on every platform); only the
reference's outputs are stored in tests/golden/metrics_*.npz."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden")
KEYS6 = ["taxa_L10", "taxa_L20", "taxa_L30", "taxa_L40", "taxa_L50", "taxa_L60"]
CASES = {
    # name: (batch sizes of the phase, classes per task, seed, fraction of null (0) targets, bias toward the target logit)
    "metrics_six": ((64, 64, 37), (1000, 400, 120, 40, 12, 4), 0, 0.3, 9.0),
    "metrics_small_heads": ((33,), (5, 2, 1, 3), 1, 0.5, 1.5),
    "metrics_all_null": ((16,), (7, 5), 2, 1.0, 2.0),
}


def make_case(name):
    """-> (keys, [(outputs {key: f32 [B, C]}, targets {key: int64 [B]}) per batch])"""
    batches, classes, seed, p_null, boost = CASES[name]
    rng = np.random.default_rng(seed)
    keys = KEYS6[: len(classes)]
    out = []
    for B in batches:
        outputs, targets = {}, {}
        for k, C in zip(keys, classes):
            z = rng.standard_normal((B, C)).astype(np.float32)
            y = rng.integers(0, C, size=B).astype(np.int64)
            y[rng.random(B) < p_null] = 0
            z[np.arange(B), y] += boost * rng.random(B).astype(np.float32)  # a realistic mix of right and wrong predictions
            outputs[k], targets[k] = z, y
        out.append((outputs, targets))
    return keys, out


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
