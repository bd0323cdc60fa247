"""Hierarchical consistency of inference predictions (second half of the handler's post-processing).

Reference: ``R/inference/postprocessing.py:14-171`` -- per sample, ranks from the highest to the lowest: a rank whose parent rank
ended up null, or whose top-1 taxon is not a child (``TaxonomyTree.get_parent``) of the parent rank's prediction, is replaced by the
single prediction (null taxon, 1.0).  Here that walk runs for the whole batch in one kernel (``lnx_hier_consistency``), in place on
the [K, B, k] index / probability tensors ``lnx_hier_topk`` produced: softmax + top-k + consistency = two launches and one
device-to-host copy per batch.

* :func:`parent_table` -- the tree as one int32 vector (built once per model).
* :func:`enforce_consistency_batch` -- the kernel on device tensors.
* :func:`topk_consistent_predictions` -- logits dict -> consistent top-k lists, one read-back.
* :func:`enforce_hierarchical_consistency` -- the reference's name and signature for ONE result object (duck-typed: any objects with
  the reference's attributes), so ``LinnaeusInferenceHandler`` can call it unchanged; it runs the same kernel on a batch of one.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import call, dt

__all__ = ["parent_table", "enforce_consistency_batch", "topk_consistent_predictions", "enforce_hierarchical_consistency"]


def parent_table(tree, task_keys, num_classes: dict, device="cuda"):
    """``(parent int32 [sum C_k] on device, class_off tuple)``: ``parent[off[k] + c]`` = class index, in ``task_keys[k + 1]``, of
    ``tree.get_parent((task_keys[k], c))``; -1 when there is none or it lives in another task (then the reference's node comparison
    fails as well).  ``task_keys`` ascend from the lowest rank.  ``tree``: a reference ``TaxonomyTree`` or anything with ``get_parent``."""
    offs, rows = [0], []
    for k, t in enumerate(task_keys):
        C = int(num_classes[t])
        nxt = task_keys[k + 1] if k + 1 < len(task_keys) else None
        row = [-1] * C
        if nxt is not None:
            for c in range(C):
                p = tree.get_parent((t, c))
                if p is not None and p[0] == nxt:
                    row[c] = int(p[1])
        rows += row
        offs.append(offs[-1] + C)
    return torch.tensor(rows, dtype=torch.int32, device=device), tuple(offs)


def enforce_consistency_batch(idx: torch.Tensor, prob: torch.Tensor, parent: torch.Tensor, class_off, null_idx, changed: torch.Tensor | None = None):
    """In place on ``idx`` int32 / ``prob`` float32 [K, B, k] (CUDA, contiguous; task 0 = lowest rank): rows that fail the
    parent / child check become {null, -1, ...} / {1, 0, ...}.  Returns ``changed`` uint8 [K, B].  No host sync."""
    if not (idx.is_cuda and prob.is_cuda and parent.is_cuda):
        raise RuntimeError("linnaeus_b200.postprocess runs on CUDA tensors only (there is no CPU path)")
    K, B, kk = idx.shape
    assert idx.dtype == torch.int32 and prob.dtype == torch.float32 and idx.is_contiguous() and prob.is_contiguous() and prob.shape == idx.shape
    assert len(class_off) == K + 1 and len(null_idx) == K and parent.numel() == class_off[-1]
    if changed is None:
        changed = torch.empty((K, B), dtype=torch.uint8, device=idx.device)
    call("lnx_hier_consistency", idx.data_ptr(), prob.data_ptr(), parent.data_ptr(), (ctypes.c_int * (K + 1))(*class_off),
         (ctypes.c_int * K)(*[int(n) for n in null_idx]), changed.data_ptr(), B, K, kk)
    return changed


def topk_consistent_predictions(outputs: dict, parent: torch.Tensor, class_off, null_idx, k: int = 5, keys: list[str] | None = None) -> dict:
    """softmax + top-k of every head (R/inference/handler.py:186-214) followed by the consistency walk
    (R/inference/postprocessing.py:14-171) for the whole batch: ``{task: (idx int64 [B, kk], prob float32 [B, kk], changed bool [B])}``
    as CPU tensors.  ``keys`` ascend from the lowest rank (the order of ``parent`` / ``class_off``)."""
    from .metrics import _cat, _require_cuda

    keys = list(outputs.keys()) if keys is None else list(keys)
    cat = getattr(outputs, "cat", None)
    if cat is not None and list(outputs.keys()) == keys:
        offs = tuple(outputs.class_off)
    else:
        cat, offs = _cat([outputs[t] for t in keys])
    _require_cuda(cat)
    assert tuple(offs) == tuple(class_off), "heads and parent table disagree on the class layout"
    K, B = len(keys), cat.shape[0]
    idx = torch.empty((K, B, k), dtype=torch.int32, device=cat.device)
    prob = torch.empty((K, B, k), dtype=torch.float32, device=cat.device)
    call("lnx_hier_topk", cat.data_ptr(), dt(cat), cat.stride(0), B, K, (ctypes.c_int * (K + 1))(*offs), int(k), idx.data_ptr(), prob.data_ptr())
    changed = enforce_consistency_batch(idx, prob, parent, offs, null_idx)
    idx_h, prob_h, ch_h = idx.cpu(), prob.cpu(), changed.cpu()
    out = {}
    for i, t in enumerate(keys):
        kk = min(k, offs[i + 1] - offs[i])
        out[t] = (idx_h[i, :, :kk].to(torch.int64), prob_h[i, :, :kk], ch_h[i].bool())
    return out


def enforce_hierarchical_consistency(result, taxonomy_data, class_maps):
    """Same name, arguments and return type as ``R/inference/postprocessing.py::enforce_hierarchical_consistency`` for one
    ``HierarchicalClassificationResult``: taxon ids are mapped to class indices, the batch-of-one kernel does the walk, and a new
    result of the same classes is returned.  Supported: every task of the result maps to a tree task key, has a non-empty
    prediction list whose top-1 taxon id is in ``class_maps``, and a null taxon id; anything else raises ``NotImplementedError``
    (the reference handles those data errors with warnings and per-case fallbacks)."""
    if not result.tasks:
        return result
    tree = taxonomy_data.taxonomy_tree
    tasks_desc = sorted(result.tasks, key=lambda t: t.rank_level.value, reverse=True)  # postprocessing.py:37
    tasks = tasks_desc[::-1]  # ascending: task 0 = lowest rank
    keys, nulls, cur = [], [], []
    for t in tasks:
        key = f"taxa_L{t.rank_level.value}"  # postprocessing.py:51-59
        if key not in tree.task_keys:
            key = f"L{t.rank_level.value}"
        null_tid = class_maps.null_taxon_ids.get(t.rank_level)
        if key not in tree.task_keys or null_tid is None or not t.predictions or t.predictions[0][0] not in class_maps.taxon_id_to_idx[t.rank_level]:
            raise NotImplementedError("enforce_hierarchical_consistency: unmapped rank, missing null taxon, empty or unknown prediction")
        keys.append(key)
        nulls.append(int(class_maps.taxon_id_to_idx[t.rank_level][null_tid]))
        cur.append(int(class_maps.taxon_id_to_idx[t.rank_level][t.predictions[0][0]]))
    parent, offs = parent_table(tree, keys, tree.num_classes)
    K = len(tasks)
    idx = torch.tensor(cur, dtype=torch.int32, device="cuda").view(K, 1, 1)
    prob = torch.tensor([float(t.predictions[0][1]) for t in tasks], dtype=torch.float32, device="cuda").view(K, 1, 1)
    changed = enforce_consistency_batch(idx, prob, parent, offs, nulls).cpu().view(K).tolist()
    new_preds = {}
    for t, ch in zip(tasks, changed):
        new_preds[t.rank_level] = [(class_maps.null_taxon_ids[t.rank_level], 1.0)] if ch else list(t.predictions)
    updated = [type(t)(rank_level=t.rank_level, temperature=t.temperature, predictions=new_preds[t.rank_level]) for t in tasks_desc]
    return type(result)(taxonomy_context=result.taxonomy_context, tasks=updated, subtree_roots=result.subtree_roots)
