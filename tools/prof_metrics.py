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
