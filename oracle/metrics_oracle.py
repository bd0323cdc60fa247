"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's validation metrics and inference top-k post-processing,
one function per reference function; pure-Python loops, meant for small cases.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this file.  Parity pinning:
``tests/test_oracle_metrics_vs_reference.py`` runs the unmodified reference functions (build container only) on the same
seeded inputs; ``tests/golden/make_golden_metrics.py`` freezes their outputs into ``tests/golden/metrics_*.npz`` for the GPU
box.  ``R/`` = ``/root/reference/linnaeus``.

Tie rule (stated, not inherited): equal logits rank by ascending class index.  ``torch.argmax`` documents exactly that
(first maximal index), so every top-1 based quantity is pinned; ``torch.topk`` leaves the order of exact ties unspecified, so
top-k (k > 1) parity with the reference is pinned on tie-free inputs only.
"""
from __future__ import annotations

import numpy as np

__all__ = ["target_rank", "accuracy", "chain_accuracy", "partial_chain_accuracy", "phase_counters", "phase_metrics", "softmax_topk"]


def _hard(t: np.ndarray) -> np.ndarray:
    """R/utils/metrics/chain_accuracy.py:145 / tracker.py:697-700: 2-D targets are arg-maxed."""
    t = np.asarray(t)
    return (t.argmax(axis=1) if t.ndim > 1 else t).astype(np.int64)


def target_rank(logits: np.ndarray, y: np.ndarray) -> np.ndarray:
    """rank[i] = #{c : z[i,c] > z[i,y_i] or (z[i,c] == z[i,y_i] and c < y_i)}; 0 <=> argmax (first maximal index) == y."""
    z = np.asarray(logits, dtype=np.float32)
    out = np.zeros(z.shape[0], dtype=np.int32)
    for i in range(z.shape[0]):
        yi = int(y[i])
        if not 0 <= yi < z.shape[1]:
            out[i] = z.shape[1]
            continue
        zy = z[i, yi]
        c = np.arange(z.shape[1])
        out[i] = int(np.sum((z[i] > zy) | ((z[i] == zy) & (c < yi))))
    return out


def accuracy(output: np.ndarray, target: np.ndarray, topk=(1,), ignore_index=None) -> list:
    """R/utils/metrics/basic.py:79-133: percent of valid samples whose target is among the k largest logits."""
    target = np.asarray(target)
    valid = np.ones(target.shape[0], dtype=bool) if ignore_index is None else target != ignore_index
    n = int(valid.sum())
    if n == 0:
        return [0.0] * len(topk)
    r = target_rank(output, np.clip(target, 0, output.shape[1] - 1))
    return [float(np.sum((r < k) & valid)) * 100.0 / n for k in topk]


def _eq_matrix(outputs_list, targets_list) -> tuple:
    gts = np.stack([_hard(t) for t in targets_list], axis=1)  # [B, K]
    eq = np.stack([target_rank(o, gts[:, k]) == 0 for k, o in enumerate(outputs_list)], axis=1)
    return eq, gts


def chain_accuracy(outputs_list, targets_list, ignore_index=None) -> float:
    """R/utils/metrics/chain_accuracy.py:51-175."""
    if ignore_index is not None:
        return 0.0
    eq, _ = _eq_matrix(outputs_list, targets_list)
    B = eq.shape[0]
    return float(eq.all(axis=1).sum()) / B if B > 0 else 1.0


def partial_chain_accuracy(outputs_list, targets_list) -> float:
    """R/utils/metrics/chain_accuracy.py:178-364: a sample counts when every task up to its highest non-null (target != 0)
    task is right; the denominator is the number of samples that have any non-null target (1.0 when there are none)."""
    eq, gts = _eq_matrix(outputs_list, targets_list)
    B, K = eq.shape
    ok = n = 0
    for i in range(B):
        nn_ranks = [k for k in range(K) if gts[i, k] != 0]
        if not nn_ranks:
            continue
        n += 1
        ok += int(all(eq[i, k] for k in range(nn_ranks[-1] + 1)))
    return ok / n if n > 0 else 1.0


def phase_counters(outputs_list, targets_list, null_index: int = 0) -> np.ndarray:
    """The int64 [4K+4] counter row one ``lnx_hier_metrics`` call adds (include/linnaeus_b200.h)."""
    eq, gts = _eq_matrix(outputs_list, targets_list)
    B, K = eq.shape
    c = np.zeros(4 * K + 4, dtype=np.int64)
    for k in range(K):  # R/utils/metrics/tracker.py:797-848: top-1 correct and count over the null-target samples of the task
        is_null = gts[:, k] == null_index
        c[2 * K + 4 + k] = int(np.sum(eq[:, k] & is_null))
        c[3 * K + 4 + k] = int(is_null.sum())
    for k, o in enumerate(outputs_list):
        r = target_rank(o, gts[:, k])
        c[k] = int(np.sum(r == 0))
        c[K + k] = c[k] if o.shape[1] < 3 else int(np.sum(r < 3))  # tracker.py:722-731
    c[2 * K] = int(eq.all(axis=1).sum())
    for i in range(B):
        nn_ranks = [k for k in range(K) if gts[i, k] != null_index]
        if nn_ranks:
            c[2 * K + 2] += 1
            c[2 * K + 1] += int(all(eq[i, k] for k in range(nn_ranks[-1] + 1)))
    c[2 * K + 3] = B
    return c


def phase_metrics(batches, keys) -> dict:
    """What MetricsTracker._update_phase_batch accumulates over a phase (R/utils/metrics/tracker.py:609-735) and
    _finalize_phase divides out: ``batches`` = iterable of (outputs dict, targets dict)."""
    keys = sorted(keys, key=lambda k: int(k.split("_L")[-1]))
    K = len(keys)
    s1 = np.zeros(K)
    s3 = np.zeros(K)
    n_ok = np.zeros(K)
    n_cnt = np.zeros(K)
    chain = partial = tot = 0.0
    for outputs, targets in batches:
        ol = [np.asarray(outputs[k], dtype=np.float32) for k in keys]
        tl = [np.asarray(targets[k]) for k in keys]
        B = ol[0].shape[0]
        c = phase_counters(ol, tl)
        s1 += c[:K]
        s3 += c[K:2 * K]
        n_ok += c[2 * K + 4:3 * K + 4]
        n_cnt += c[3 * K + 4:4 * K + 4]
        chain += chain_accuracy(ol, tl) * B
        partial += partial_chain_accuracy(ol, tl) * B
        tot += B
    return {"acc1": {k: 100.0 * s1[i] / tot for i, k in enumerate(keys)}, "acc3": {k: 100.0 * s3[i] / tot for i, k in enumerate(keys)},
            "chain_accuracy": chain / tot, "partial_chain_accuracy": partial / tot, "samples": int(tot),
            "null_acc1": {k: 100.0 * n_ok[i] / n_cnt[i] for i, k in enumerate(keys) if n_cnt[i] > 0},
            "non_null_acc1": {k: 100.0 * (s1[i] - n_ok[i]) / (tot - n_cnt[i]) for i, k in enumerate(keys) if tot - n_cnt[i] > 0}}


def softmax_topk(logits: np.ndarray, k: int) -> tuple:
    """R/inference/handler.py:196-203: probs = softmax(logits); top min(k, C) by (probability desc, index asc)."""
    z = np.asarray(logits, dtype=np.float32)
    kk = min(k, z.shape[1])
    e = np.exp(z - z.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    order = np.lexsort((np.broadcast_to(np.arange(z.shape[1]), z.shape), -z), axis=1)[:, :kk]
    return order.astype(np.int64), np.take_along_axis(p, order, axis=1).astype(np.float32)
