"""Per-entry-point timing of one mFormerV1 inference forward (LNX_PROFILE event brackets)."""
import os, sys
from collections import defaultdict
os.environ["LNX_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import linnaeus_b200 as L
from linnaeus_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
cfg, nc = L.make_synthetic_config("sm", 224)
torch.manual_seed(0)
model = L.build_model(cfg, nc).to(dev).eval().set_compute_dtype(torch.bfloat16)
img = torch.randn(B, 3, 224, 224, device=dev)
meta = torch.randn(B, 15, device=dev)
with torch.no_grad():
    for _ in range(2):
        model(img, meta)
    torch.cuda.synchronize()
    _lib.profile_log.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model(img, meta)
    e1.record()
torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
tot = 0.0
for name, key, a, b in _lib.profile_log:
    ms = a.elapsed_time(b)
    agg[(name, key)][0] += 1
    agg[(name, key)][1] += ms
    tot += ms
print(f"forward (eager) {e0.elapsed_time(e1):.2f} ms; sum of kernel brackets {tot:.2f} ms; calls {len(_lib.profile_log)}")
byname = defaultdict(float)
for (name, key), (n, ms) in agg.items():
    byname[name] += ms
for name, ms in sorted(byname.items(), key=lambda kv: -kv[1]):
    print(f"{name:24s} {ms:8.3f} ms  {100 * ms / tot:5.1f}%")
for (name, key), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{name:22s} n={n:3d} tot={ms:8.3f} avg={ms / n:7.3f}  {key}")
