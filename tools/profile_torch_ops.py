"""Which torch (aten) ops still launch kernels inside one eager train step, and from where (launch hygiene).
python tools/profile_torch_ops.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import linnaeus_b200 as L
from linnaeus_b200.engine import TrainStep
from linnaeus_b200.optim import FlatAdamW

dev = torch.device("cuda", 0)
cfg, nc = L.make_synthetic_config("sm", 224)
torch.manual_seed(0)
model = L.build_model(cfg, nc).to(dev).set_compute_dtype(torch.bfloat16).train()
B = 64
x, meta = torch.randn(B, 3, 224, 224, device=dev), torch.randn(B, 15, device=dev)
tg = {k: torch.randint(0, c, (B,), device=dev) for k, c in nc.items()}
opt = FlatAdamW(model.named_parameters(), lr=1e-4, clip_grad=5.0)
ts = TrainStep(model, opt, list(nc), nc, kind="ce", config=cfg)
for _ in range(2):
    ts.step(x, meta, tg)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    ts.step(x, meta, tg)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_stack_n=6):
    if e.key.startswith("aten::") and e.device_time_total > 0 and e.count > 0:
        stack = [s for s in e.stack if "linnaeus_b200" in s][:2]
        rows.append((e.count, e.device_time_total, e.key, " <- ".join(s.split("/")[-1] for s in stack)))
rows.sort(key=lambda r: -r[0])
tot = sum(r[0] for r in rows)
print("aten ops with device time:", tot, "calls,", sum(r[1] for r in rows), "us")
for c, t, k, s in rows[:60]:
    print(f"{c:4d} {t:8.1f} us  {k:28s} {s}")
