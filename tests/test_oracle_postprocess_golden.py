"""CPU: the consistency oracle against the frozen outputs of the unmodified reference function (tests/golden/consistency_*.npz)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, f"consistency_{name}.npz"))
    offs = z["class_off"]
    parent = [z["parent"][offs[k]:offs[k + 1]].tolist() for k in range(len(offs) - 1)]
    return z, parent


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_matches_reference_golden(name):
    from oracle.postprocess_oracle import enforce_consistency

    z, parent = load(name)
    K, B, kk = z["idx"].shape
    for b in range(B):
        preds = [[(int(c), float(p)) for c, p in zip(z["idx"][k, b], z["prob"][k, b])] for k in range(K)]
        got, changed = enforce_consistency(preds, parent, z["null_idx"].tolist())
        for k in range(K):
            assert bool(z["changed"][k, b]) == changed[k]
            if changed[k]:
                assert got[k] == [(int(z["out_idx"][k, b, 0]), 1.0)]
            else:
                assert [c for c, _ in got[k]] == z["out_idx"][k, b].tolist()
