"""Measurement of the 8(f) N1 row: one GradNorm weight update (one forward + K backward passes, K = 6) at the bench shape
(mFormerV1_sm, B = 256, bf16) against K re-forwards + backward passes on the same B200 model (the reference's schedule)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200 as L
import linnaeus_b200.loss as LL
from linnaeus_b200 import gradnorm as G
from linnaeus_b200.optim import FlatAdamW

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg, nc = L.make_synthetic_config("sm", 224)
torch.manual_seed(0)
model = L.build_model(cfg, nc).to(dev).set_compute_dtype(torch.bfloat16).train()
opt = FlatAdamW(model.named_parameters(), lr=1e-4)
keys = list(nc.keys())
x, meta = torch.randn(B, 3, 224, 224, device=dev), torch.randn(B, 15, device=dev)
tg = {k: torch.randint(0, nc[k], (B,), device=dev) for k in keys}
crit = {k: LL.CrossEntropyLoss() for k in keys}
gn = G.GradNormModule(keys, alpha=1.5).to(dev)


def ours():
    G.update_gradnorm_weights(gn, model, (x, tg, meta), crit, optimizer=opt, return_metrics=False)


def reforward_schedule():
    params = G.backbone_parameters(model)
    for k in keys:
        opt.zero_grad()
        out = model(x, torch.zeros_like(meta))
        valid = (tg[k] != 0).float()
        lv = crit[k](out[k], tg[k])
        ((lv * valid).sum() / valid.sum().clamp(min=1.0)).backward()
        torch.stack(torch._foreach_norm([p.grad for p in params])).norm(2)
    opt.zero_grad()


def timeit(fn, n=3):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"GradNorm update, one forward + {len(keys)} backward passes (eager): {timeit(ours):.1f} ms")
print(f"same measurement with {len(keys)} re-forwards + backward passes (the reference's schedule, eager): {timeit(reforward_schedule):.1f} ms")
print("task weights after the updates:", [round(v, 4) for v in gn.task_weights.tolist()])
