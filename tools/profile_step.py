"""Per-entry-point timing of one eager training step (diagnostic; LNX_PROFILE event brackets).

    LNX_PROFILE=1 python tools/profile_step.py [batch] [variant] [img]

Prints, per C-ABI entry point and integer-argument signature (shape), the number of calls
in one step, total and average milliseconds (CUDA events, warm), sorted by total time, plus
the time of the step not covered by our kernels (torch glue ops).
"""
import os
import sys
from collections import defaultdict

os.environ["LNX_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200 as L
from linnaeus_b200 import _lib
from linnaeus_b200.engine import TrainStep
from linnaeus_b200.optim import FlatAdamW

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variant = sys.argv[2] if len(sys.argv) > 2 else "sm"
S = int(sys.argv[3]) if len(sys.argv) > 3 else 224
dev = torch.device("cuda", 0)
cfg, nc = L.make_synthetic_config(variant, S)
torch.manual_seed(0)
model = L.build_model(cfg, nc).to(dev)
model.set_compute_dtype(torch.bfloat16)
keys = list(nc.keys())
g = torch.Generator().manual_seed(42)
img = torch.randn(B, 3, S, S, generator=g).to(dev)
meta = torch.randn(B, 15, generator=g).to(dev)
tg = {k: torch.randint(0, c, (B,), generator=g).to(dev) for k, c in nc.items()}
opt = FlatAdamW(model.named_parameters(), lr=1e-4, weight_decay=0.05, clip_grad=5.0, grad_scale=1.0)
model.train()
ts = TrainStep(model, opt, keys, nc, kind="ce", config=cfg)
for _ in range(2):
    ts.step(img, meta, tg)
torch.cuda.synchronize()
_lib.profile_log.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ts.step(img, meta, tg)
e1.record()
torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
tot = 0.0
for name, key, a, b in _lib.profile_log:
    ms = a.elapsed_time(b)
    agg[(name, key)][0] += 1
    agg[(name, key)][1] += ms
    tot += ms
print(f"step (eager, events included) {e0.elapsed_time(e1):.2f} ms; sum of kernel brackets {tot:.2f} ms; calls {len(_lib.profile_log)}")
byname = defaultdict(float)
for (name, key), (n, ms) in agg.items():
    byname[name] += ms
print("--- by entry point")
for name, ms in sorted(byname.items(), key=lambda kv: -kv[1]):
    print(f"{name:24s} {ms:8.3f} ms  {100 * ms / tot:5.1f}%")
print("--- by entry point and integer args (gemm: dt,lda,aT,ldb,bT,odt,M,N,K,act,rpg,acc,simt)")
for (name, key), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{name:22s} n={n:3d} tot={ms:8.3f} avg={ms / n:7.3f}  {key}")
