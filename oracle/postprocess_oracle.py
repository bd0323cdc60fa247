"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pure-Python restatement of the reference's hierarchical-consistency post-processing of inference results
(R/inference/postprocessing.py:14-171), in class-index space: one sample at a time, ranks walked from the highest
to the lowest.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this file.

Parity pinning: ``tests/test_oracle_postprocess_vs_reference.py`` runs the unmodified reference function (build container only; the
absent ``typus`` package is shimmed under tests/support/ref_shims) on the same seeded trees and predictions;
``tests/golden/make_golden_consistency.py`` freezes its outputs into ``tests/golden/consistency_*.npz`` for the GPU box.
``R/`` = ``/root/reference/linnaeus``.
"""
from __future__ import annotations

__all__ = ["enforce_consistency"]


def enforce_consistency(preds, parent, null_idx):
    """``preds[k]`` = [(class_idx, prob), ...] best first for task k (task 0 = lowest rank, K - 1 = highest), non-empty;
    ``parent[k][c]`` = class index in task k + 1 of the tree parent of class c of task k, or -1 (no link; also every class of the
    highest task); ``null_idx[k]`` = null class of task k or -1.  Returns (new preds, changed flags).

    R/inference/postprocessing.py:64-151: the highest rank stands (:149-150).  Below it: parent null (:108-117) -> this rank becomes
    [(null, 1.0)] if it has a null class (:121-124), else it stands (:126); otherwise if ``tree.get_parent`` of the top-1 node is not
    the parent rank's consistent node (:131-132) -> [(null, 1.0)] if possible (:138-141) else it stands (:143); else it stands (:146).
    """
    K = len(preds)
    out = [list(p) for p in preds]
    changed = [False] * K
    cons = None
    for k in range(K - 1, -1, -1):
        cur = preds[k][0][0]
        nullify = False
        if k < K - 1:
            parent_is_null = null_idx[k + 1] >= 0 and cons == null_idx[k + 1]
            if parent_is_null:
                nullify = True
            else:
                actual = parent[k][cur] if 0 <= cur < len(parent[k]) else -1
                nullify = actual != cons
        if nullify and null_idx[k] >= 0:
            out[k] = [(null_idx[k], 1.0)]
            changed[k] = True
            cons = null_idx[k]
        else:
            cons = cur
    return out, changed
