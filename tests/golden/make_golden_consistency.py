"""Freeze outputs of the UNMODIFIED reference R/inference/postprocessing.py::enforce_hierarchical_consistency (build container only)
into tests/golden/consistency_*.npz: a random taxonomy as a flat parent table, per-rank top-k inputs, and the reference's outputs in
the kernel's encoding (nullified rows = {null, -1, ...} / {1, 0, ...}).   python tests/golden/make_golden_consistency.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.support.consistency_cases import make_case  # noqa: E402
from tests.test_oracle_postprocess_vs_reference import reference_objects, reference_run  # noqa: E402

CASES = {"a": dict(seed=11, K=6, B=256, kk=5, null_links=True, p_null=0.05, p_consistent=0.92), "b": dict(seed=12, K=6, B=100, kk=3, null_links=False),
         "c": dict(seed=13, K=3, B=33, kk=1, null_links=True, p_null=0.5)}

for name, kw in CASES.items():
    case = make_case(**kw)
    objs = reference_objects(case)
    K, B, kk = case["idx"].shape
    out_idx, out_prob = case["idx"].copy(), case["prob"].copy()
    changed = np.zeros((K, B), dtype=np.uint8)
    for b in range(B):
        ref = reference_run(case, b, objs)
        for k in range(K):
            if len(ref[k]) == 1 and kk >= 1 and (len(ref[k]) != kk or ref[k][0] != (int(case["idx"][k, b, 0]), float(case["prob"][k, b, 0]))):
                out_idx[k, b] = [ref[k][0][0]] + [-1] * (kk - 1)
                out_prob[k, b] = [ref[k][0][1]] + [0.0] * (kk - 1)
                changed[k, b] = 1
            else:
                assert [c for c, _ in ref[k]] == [int(c) for c in case["idx"][k, b]]
    offs = np.cumsum([0] + [len(r) for r in case["parent"]]).astype(np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"consistency_{name}.npz"), parent=np.concatenate([np.asarray(r, np.int32) for r in case["parent"]]),
                        class_off=offs, null_idx=np.asarray(case["null_idx"], np.int32), idx=case["idx"], prob=case["prob"], out_idx=out_idx, out_prob=out_prob,
                        changed=changed)
    print(name, "changed rows:", int(changed.sum()), "of", K * B)
