"""Seeded random cases for the hierarchical-consistency post-processing (tests and golden generation): a random taxonomy in class
index space, per-rank top-k prediction lists, and the mapping to the taxon-id space the reference function works in."""
from __future__ import annotations

import numpy as np

LEVELS = (10, 20, 30, 40, 50, 60, 70)


def make_case(seed: int, K: int = 6, B: int = 64, kk: int = 3, null_links: bool = True, p_null: float = 0.25, p_consistent: float = 0.5,
              missing_link: float = 0.05):
    """Returns dict(task_keys, num_classes, hierarchy_map {child_task: {child: parent}}, parent list-of-lists, null_idx,
    idx int32 [K, B, kk], prob float32 [K, B, kk]).  Task 0 = lowest rank.  Class 0 of every task is its null class."""
    rng = np.random.default_rng(seed)
    task_keys = [f"taxa_L{LEVELS[k]}" for k in range(K)]
    C = [int(max(3, 40 // (k + 1) + rng.integers(0, 5))) for k in range(K)]
    num_classes = {t: c for t, c in zip(task_keys, C)}
    hierarchy_map, parent = {}, []
    for k in range(K):
        row = [-1] * C[k]
        if k + 1 < K:
            links = {}
            for c in range(C[k]):
                if c == 0:
                    if null_links:
                        links[0] = 0
                        row[0] = 0
                    continue
                if rng.random() < missing_link:
                    continue
                p = int(rng.integers(1, C[k + 1]))
                links[c] = p
                row[c] = p
            hierarchy_map[task_keys[k]] = links
        parent.append(row)
    idx = np.zeros((K, B, kk), dtype=np.int32)
    prob = np.zeros((K, B, kk), dtype=np.float32)
    for b in range(B):
        above = None
        for k in range(K - 1, -1, -1):
            r = rng.random()
            if r < p_null:
                top = 0
            elif above is not None and above != 0 and rng.random() < p_consistent:
                kids = [c for c in range(1, C[k]) if parent[k][c] == above]
                top = int(rng.choice(kids)) if kids else int(rng.integers(1, C[k]))
            else:
                top = int(rng.integers(1, C[k]))
            rest = [c for c in rng.permutation(C[k]) if c != top][:kk - 1]
            idx[k, b] = [top] + [int(c) for c in rest]
            pr = np.sort(rng.random(kk).astype(np.float32))[::-1]
            prob[k, b] = pr / pr.sum() * np.float32(0.9)
            above = top
    return {"task_keys": task_keys, "num_classes": num_classes, "hierarchy_map": hierarchy_map, "parent": parent, "null_idx": [0] * K,
            "idx": idx, "prob": prob}


def taxon_id(level: int, c: int) -> int:
    return 100000 * level + 7 * c + 3
