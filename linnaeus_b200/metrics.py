"""Validation metrics and inference post-processing right after the hot path.

Host-side mirror of the reference's metric functions, same names and return conventions:

* ``accuracy(output, target, topk, ignore_index)``                R/utils/metrics/basic.py:79-133
* ``compute_chain_accuracy_vectorized(outputs_list, targets_list, ignore_index)``
* ``compute_partial_chain_accuracy_vectorized(outputs_list, targets_list)``
                                                                  R/utils/metrics/chain_accuracy.py:51-364
* ``HierMetricsAccumulator``: what ``MetricsTracker._update_phase_batch`` accumulates over a validation phase
  (R/utils/metrics/tracker.py:609-735): per-task top-1 / top-3, chain and partial-chain accuracy - but WITHOUT the
  reference's ~4K+6 ``.item()`` host syncs per batch: ``update()`` only enqueues one kernel, ``compute()`` does one
  device->host read for the whole phase.
* ``topk_predictions(outputs, k)``: softmax + top-k of every head for the whole batch in one launch and one read-back
  (R/inference/handler.py:186-214 loops over samples x tasks x k with ``.item()``).

Everything runs on the ``lnx_hier_metrics`` / ``lnx_hier_topk`` CUDA kernels; CPU tensors raise (no fallback).
Tie rule: among exactly equal logits the lower class index ranks first - ``torch.argmax``'s rule (first maximal index), so
top-1 matches the reference bit for bit; ``torch.topk`` leaves the order of exact ties unspecified.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import call, dt


def _task_sort_key(k: str) -> int:
    return int(k.split("_L")[-1])  # tracker.py:620 / chain_accuracy.py usage note


def _hard(t: torch.Tensor) -> torch.Tensor:
    """[B] class indices; one-hot / soft [B, C] targets are arg-maxed (chain_accuracy.py:145, tracker.py:697-700)."""
    return (t.argmax(dim=1) if t.dim() > 1 else t).to(torch.int64)


def _cat(outputs_list) -> tuple[torch.Tensor, tuple]:
    offs = [0]
    for o in outputs_list:
        offs.append(offs[-1] + o.shape[1])
    if len({o.dtype for o in outputs_list}) != 1 or outputs_list[0].dtype not in (torch.float32, torch.bfloat16):
        outputs_list = [o.float() for o in outputs_list]
    cat = outputs_list[0] if len(outputs_list) == 1 else torch.cat(list(outputs_list), dim=1)
    return cat.contiguous(), tuple(offs)


def _require_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("linnaeus_b200.metrics runs on CUDA (sm_100a) only; there is no CPU fallback")


def hier_metrics(cat: torch.Tensor, class_off, targets_kb: torch.Tensor, null_index: int = 0, counters: torch.Tensor | None = None,
                 want_ranks: bool = False):
    """Raw ``lnx_hier_metrics`` call.  cat [B, >= class_off[-1]] (f32 / bf16), targets_kb int64 [K, B].  ``counters`` int64
    [4K+4] is ADDED to (created zeroed when None and ``want_ranks`` is False).  Returns (ranks int32 [K, B] | None, counters)."""
    _require_cuda(cat)
    K = len(class_off) - 1
    B = cat.shape[0]
    if targets_kb.shape != (K, B):
        raise ValueError(f"targets must be [K={K}, B={B}], got {tuple(targets_kb.shape)}")
    targets_kb = targets_kb.to(device=cat.device, dtype=torch.int64).contiguous()
    ranks = torch.empty((K, B), dtype=torch.int32, device=cat.device) if want_ranks else None
    if counters is None and not want_ranks:
        counters = torch.zeros(4 * K + 4, dtype=torch.int64, device=cat.device)
    offs = (ctypes.c_int * (K + 1))(*class_off)
    call("lnx_hier_metrics", cat.data_ptr(), dt(cat), cat.stride(0), B, K, offs, targets_kb.data_ptr(), int(null_index),
         0 if ranks is None else ranks.data_ptr(), 0 if counters is None else counters.data_ptr())
    return ranks, counters


# ----------------------------------------------------------------------------- reference-named functions
def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,), ignore_index: int | None = None) -> list[float]:
    """Top-k accuracies in percent over the valid samples (R/utils/metrics/basic.py:79-133)."""
    _require_cuda(output)
    tgt = target.to(output.device)
    cat, offs = _cat([output])
    # an ignored target never counts; give the kernel a valid class index for those rows
    safe = tgt.to(torch.int64).clamp(0, output.shape[1] - 1)
    ranks, _ = hier_metrics(cat, offs, safe.view(1, -1), want_ranks=True)
    ranks = ranks[0]
    if ignore_index is not None:
        valid = tgt != ignore_index
        n = int(valid.sum().item())
        if n == 0:
            return [0.0] * len(topk)
    else:
        valid = torch.ones_like(tgt, dtype=torch.bool)
        n = tgt.shape[0]
        if n == 0:
            return [0.0] * len(topk)
    ks = torch.tensor([min(int(k), output.shape[1]) for k in topk], device=output.device, dtype=torch.int32)
    correct = ((ranks.view(1, -1) < ks.view(-1, 1)) & valid.view(1, -1)).sum(dim=1)
    return [float(c) * 100.0 / n for c in correct.tolist()]


def _chain_counts(outputs_list, targets_list) -> list[int]:
    cat, offs = _cat(list(outputs_list))
    tg = torch.stack([_hard(t).to(cat.device) for t in targets_list])
    _, counters = hier_metrics(cat, offs, tg)
    return counters.tolist()


def compute_chain_accuracy_vectorized(outputs_list, targets_list, ignore_index: int | None = None) -> float:
    """Fraction of samples with every task right; 0.0 when ``ignore_index`` is set (chain_accuracy.py:51-175)."""
    if ignore_index is not None:
        return 0.0
    c = _chain_counts(outputs_list, targets_list)
    K = len(outputs_list)
    n = c[2 * K + 3]
    return c[2 * K] / n if n > 0 else 1.0


def compute_partial_chain_accuracy_vectorized(outputs_list, targets_list) -> float:
    """Samples right on tasks 0..highest non-null task / samples with any non-null target; 1.0 when there are none
    (chain_accuracy.py:178-364)."""
    c = _chain_counts(outputs_list, targets_list)
    K = len(outputs_list)
    return c[2 * K + 1] / c[2 * K + 2] if c[2 * K + 2] > 0 else 1.0


class HierMetricsAccumulator:
    """Per-phase accumulation of MetricsTracker._update_phase_batch (tracker.py:609-735) with no host sync per batch.

    ``update(outputs, targets)`` takes the model's ``{task: logits}`` dict (tasks are ordered by their ``_L<n>`` suffix as the
    tracker does) and the targets dict; ``compute()`` returns ``{"acc1": {task: %}, "acc3": {task: %}, "chain_accuracy": f,
    "partial_chain_accuracy": f, "samples": n, "null_acc1": {task: %}, "non_null_acc1": {task: %}}`` (null = target index ==
    ``null_index``; for one-hot targets that is the reference's ``target[:, 0] > 0.5``) with the tracker's arithmetic: chain / partial-chain are the batch-size
    weighted means of the per-batch values (a batch without any non-null sample contributes 1.0, chain_accuracy.py:351)."""

    def __init__(self, null_index: int = 0):
        self.null_index = int(null_index)
        self._rows: list[torch.Tensor] = []
        self._keys: list[str] | None = None

    def reset(self) -> None:
        self._rows, self._keys = [], None

    def update(self, outputs: dict, targets: dict) -> None:
        keys = sorted(outputs.keys(), key=_task_sort_key)
        if self._keys is None:
            self._keys = keys
        elif keys != self._keys:
            raise ValueError(f"task keys changed within a phase: {keys} vs {self._keys}")
        cat = getattr(outputs, "cat", None)
        if cat is not None and list(outputs.keys()) == keys:
            offs = tuple(outputs.class_off)  # the model's single head-GEMM output, no concat pass
        else:
            cat, offs = _cat([outputs[k] for k in keys])
        tg = torch.stack([_hard(targets[k]).to(cat.device) for k in keys])
        _, counters = hier_metrics(cat, offs, tg, null_index=self.null_index)
        self._rows.append(counters)

    def compute(self, all_reduce: bool | None = None, group=None) -> dict:
        """One device->host read for the phase.  Under ``torch.distributed`` (``all_reduce`` None = "if initialised") the local
        sums and counts are summed over the ranks before dividing, as ``MetricsTracker._finalize_phase`` does with its accumulators
        (R/utils/metrics/tracker.py:1112-1136, 1209-1231)."""
        if not self._rows:
            rows, dev = [], torch.device("cpu")
        else:
            rows, dev = torch.stack(self._rows).tolist(), self._rows[0].device  # the phase's only device->host read
        return self._finalize(rows, self._keys or [], dev, all_reduce, group)

    @staticmethod
    def _finalize(rows: list, keys: list[str], device, all_reduce: bool | None = None, group=None) -> dict:
        """rows: per-batch counter rows (lnx_hier_metrics layout).  chain / partial-chain follow the tracker's arithmetic: the sum
        over batches of (per-batch value x batch size) over the sum of batch sizes (a batch without any non-null sample counts as
        1.0, chain_accuracy.py:351)."""
        import torch.distributed as dist

        K = len(keys)
        if all_reduce is None:
            all_reduce = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if not rows and not all_reduce:
            return {"acc1": {}, "acc3": {}, "chain_accuracy": 0.0, "partial_chain_accuracy": 0.0, "samples": 0}
        # top-1 sums, top-3 sums, chain sum, partial-chain sum, samples, null-target top-1 sums, null-target counts
        local = [0.0] * (4 * K + 3)
        for r in rows:
            for i in range(2 * K):
                local[i] += r[i]
            local[2 * K] += r[2 * K]
            local[2 * K + 1] += (r[2 * K + 1] / r[2 * K + 2] if r[2 * K + 2] > 0 else 1.0) * r[2 * K + 3]
            local[2 * K + 2] += r[2 * K + 3]
            for i in range(2 * K):
                local[2 * K + 3 + i] += r[2 * K + 4 + i]
        if all_reduce:
            t = torch.tensor(local, dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            local = t.tolist()
        tot = local[2 * K + 2]
        if tot <= 0:
            return {"acc1": {}, "acc3": {}, "chain_accuracy": 0.0, "partial_chain_accuracy": 0.0, "samples": 0}
        null_ok, null_n = local[2 * K + 3:3 * K + 3], local[3 * K + 3:4 * K + 3]
        return {"acc1": {k: 100.0 * local[i] / tot for i, k in enumerate(keys)}, "acc3": {k: 100.0 * local[K + i] / tot for i, k in enumerate(keys)},
                "chain_accuracy": local[2 * K] / tot, "partial_chain_accuracy": local[2 * K + 1] / tot, "samples": int(round(tot)),
                # tracker.py:1376-1392 / 1463-1479: only tasks that saw such samples get a value
                "null_acc1": {k: 100.0 * null_ok[i] / null_n[i] for i, k in enumerate(keys) if null_n[i] > 0},
                "non_null_acc1": {k: 100.0 * (local[i] - null_ok[i]) / (tot - null_n[i]) for i, k in enumerate(keys) if tot - null_n[i] > 0}}


def topk_predictions(outputs: dict, k: int = 5, keys: list[str] | None = None) -> dict:
    """softmax + top-k of every head for the whole batch: ``{task: (idx int64 [B, min(k, C_task)], prob float32 [same])}`` as
    CPU tensors (one launch, one read-back), the per-sample / per-task / per-k loop of R/inference/handler.py:186-214."""
    keys = list(outputs.keys()) if keys is None else list(keys)
    cat = getattr(outputs, "cat", None)
    if cat is not None and list(outputs.keys()) == keys:
        offs = tuple(outputs.class_off)
    else:
        cat, offs = _cat([outputs[t] for t in keys])
    _require_cuda(cat)
    K, B = len(keys), cat.shape[0]
    idx = torch.empty((K, B, k), dtype=torch.int32, device=cat.device)
    prob = torch.empty((K, B, k), dtype=torch.float32, device=cat.device)
    call("lnx_hier_topk", cat.data_ptr(), dt(cat), cat.stride(0), B, K, (ctypes.c_int * (K + 1))(*offs), int(k), idx.data_ptr(), prob.data_ptr())
    idx_h, prob_h = idx.cpu(), prob.cpu()
    out = {}
    for i, t in enumerate(keys):
        kk = min(k, offs[i + 1] - offs[i])
        out[t] = (idx_h[i, :, :kk].to(torch.int64), prob_h[i, :, :kk])
    return out
