"""Weight-gradient GEMM (lnx_wgrad) at the train step's shapes: time at M and M / 2 separates the reduction main loop from the fixed
tail (pipeline fill + the fp32 atomic flush of the split-K partials)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200.functional as F

dev = "cuda"
shapes = {"s0 pw1": (256 * 3136, 384, 96), "s0 pw2": (256 * 3136, 96, 384), "s1 pw1": (200704, 768, 192), "s1 pw2": (200704, 192, 768),
          "s2 fc1": (50176, 1536, 384), "s2 fc2": (50176, 384, 1536), "s2 qkv": (50176, 1152, 384), "s2 proj": (50176, 384, 384),
          "s3 fc1": (12544, 3072, 768), "s3 fc2": (12544, 768, 3072), "s3 qkv": (12544, 2304, 768), "s3 proj": (12544, 768, 768)}


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, (M, N, K) in shapes.items():
    dy = torch.randn(M, N, device=dev).bfloat16()
    x = torch.randn(M, K, device=dev).bfloat16()
    dw = torch.zeros(N, K, device=dev)
    db = torch.zeros(N, device=dev)
    t1 = timeit(lambda: F.wgrad(dy, x, out=dw, db_out=db))
    h = M // 2
    t2 = timeit(lambda: F.wgrad(dy[:h], x[:h], out=dw, db_out=db))
    tail = 2 * t2 - t1
    print(f"{name:8s} M={M:7d} N={N:5d} K={K:5d}: {t1 * 1e3:6.1f} us ({2 * M * N * K / t1 / 1e9:5.0f} TF/s, {(M * (N + K) * 2) / t1 / 1e9:5.2f} TB/s)  half M {t2 * 1e3:6.1f} us"
          f"  -> fixed part ~{tail * 1e3:5.1f} us", flush=True)
