"""Measurement of the 8(f) N2 / N4 rows: the fused validation-metrics kernel and the batched softmax + top-k against the
reference's call sequence (torch ops + .item() per task / per sample, restated from tracker.py:609-735 and handler.py:186-214)
run on the same GPU tensors, B = 256, 6 tasks (1000, 400, 120, 40, 12, 4)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200.metrics as M

dev = "cuda"
B, classes = 256, (1000, 400, 120, 40, 12, 4)
keys = [f"taxa_L{10 * (i + 1)}" for i in range(len(classes))]
torch.manual_seed(0)
outs = {k: torch.randn(B, C, device=dev) for k, C in zip(keys, classes)}
tgts = {k: torch.randint(0, C, (B,), device=dev) for k, C in zip(keys, classes)}


def reference_sequence():
    ol, tl = [outs[k] for k in keys], [tgts[k] for k in keys]
    eq = torch.stack([o.argmax(1) == t for o, t in zip(ol, tl)], dim=1)
    chain = eq.all(dim=1).sum().item() / B
    g = torch.stack(tl, dim=1)
    nn_mask = g != 0
    idx = torch.arange(len(keys), device=dev).expand(B, -1)
    hi = idx.masked_fill(~nn_mask, -1).max(dim=1)[0]
    has = hi >= 0
    pc = (torch.logical_or(~(idx <= hi.unsqueeze(1)), eq).all(dim=1) & has)
    partial = pc.sum().item() / max(has.sum().item(), 1)
    res = []
    for o, t in zip(ol, tl):
        c1 = (o.argmax(1) == t).sum().item()
        c3 = (o.topk(3, dim=1)[1] == t.unsqueeze(1)).any(dim=1).sum().item()
        res.append((c1, c3))
    return chain, partial, res


def reference_topk(k=5):
    r = []
    for i in range(B):
        for key in keys:
            p = torch.softmax(outs[key][i], dim=-1)
            tp, ti = torch.topk(p, k=min(k, p.shape[0]))
            r.append([(ti[j].item(), tp[j].item()) for j in range(ti.shape[0])])
    return r


def timeit(fn, n):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


acc = M.HierMetricsAccumulator()
cat = torch.cat([outs[k] for k in keys], dim=1)
offs = [0]
for c in classes:
    offs.append(offs[-1] + c)
tg = torch.stack([tgts[k] for k in keys])
counters = torch.zeros(4 * len(keys) + 4, dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
M.hier_metrics(cat, offs, tg, counters=counters)
e0.record()
for _ in range(100):
    M.hier_metrics(cat, offs, tg, counters=counters)
e1.record()
torch.cuda.synchronize()
kern_us = e0.elapsed_time(e1) / 100 * 1e3
nbytes = cat.numel() * 4
print(f"lnx_hier_metrics kernel (device time incl. launch gap): {kern_us:.1f} us per batch; logits {nbytes / 1e6:.2f} MB -> {nbytes / kern_us / 1e3:.0f} GB/s (launch bound)")
print(f"HierMetricsAccumulator.update (host wall, no sync):     {timeit(lambda: acc.update(outs, tgts), 50):.3f} ms per batch")
print(f"reference call sequence on the same GPU tensors:        {timeit(reference_sequence, 20):.3f} ms per batch ({2 * len(keys) + 3} .item() syncs)")
print(f"topk_predictions(k=5), one launch + one read-back:      {timeit(lambda: M.topk_predictions(outs, 5), 20):.3f} ms per batch")
print(f"reference per-sample softmax/topk/.item() loop:         {timeit(reference_topk, 2):.1f} ms per batch")

# ---- hierarchical consistency (postprocessing.py:14-171): batch kernel behind the top-k against the per-sample Python walk
import numpy as np  # noqa: E402

import linnaeus_b200.postprocess as PP  # noqa: E402
from linnaeus_b200.config import SyntheticTaxonomy  # noqa: E402

tax = SyntheticTaxonomy({k: c for k, c in zip(keys, classes)})


class _Tree:  # the one method the walk needs from a TaxonomyTree
    task_keys, num_classes = keys, tax.num_classes

    @staticmethod
    def get_parent(node):
        t, c = node
        i = keys.index(t)
        return (keys[i + 1], tax.parent_of(t, c)) if i + 1 < len(keys) and c != 0 else None


parent, poffs = PP.parent_table(_Tree, keys, tax.num_classes)
parent_rows = [parent[poffs[i]:poffs[i + 1]].tolist() for i in range(len(keys))]
nulls = [0] * len(keys)


def reference_walk(top):
    """the reference's per-sample loop in class-index space (dict lookups + tuple compares per rank), after the top-k lists exist"""
    out = []
    for i in range(B):
        cons = None
        row = []
        for k in range(len(keys) - 1, -1, -1):
            cur = int(top[keys[k]][0][i, 0])
            nullify = False
            if k < len(keys) - 1:
                if cons == nulls[k + 1]:
                    nullify = True
                else:
                    nullify = parent_rows[k][cur] != cons
            cons = nulls[k] if nullify else cur
            row.append((cons, nullify))
        out.append(row)
    return out


top_cpu = M.topk_predictions(outs, 5)
t_fused = timeit(lambda: PP.topk_consistent_predictions(outs, parent, poffs, nulls, 5), 20)
t_walk = timeit(lambda: reference_walk(top_cpu), 5)
idx_d = torch.zeros((len(keys), B, 5), dtype=torch.int32, device=dev)
prob_d = torch.rand((len(keys), B, 5), device=dev)
ch_d = torch.empty((len(keys), B), dtype=torch.uint8, device=dev)
PP.enforce_consistency_batch(idx_d, prob_d, parent, poffs, nulls, ch_d)
e0.record()
for _ in range(100):
    PP.enforce_consistency_batch(idx_d, prob_d, parent, poffs, nulls, ch_d)
e1.record()
torch.cuda.synchronize()
print(f"topk + hierarchical consistency, 2 launches + 1 read-back: {t_fused:.3f} ms per batch (lnx_hier_consistency alone: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us)")
print(f"per-sample consistency walk in Python (after the top-k loop): {t_walk:.2f} ms per batch")
