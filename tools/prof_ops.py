"""Micro-benchmarks of the HBM/FMA-bound kernels at the bench shapes (CUDA events, warm)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200.functional as F
from linnaeus_b200._lib import call, dt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda"


def bench(name, fn, nbytes, n=5, flops=None):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    extra = f"  {flops / ms / 1e9:6.1f} TFMA/s" if flops else ""
    print(f"{name:34s} {ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s{extra}")


for (H, C) in ((56, 96), (28, 192)):
    x = torch.randn(B, H, H, C, device=dev).bfloat16()
    dy = torch.randn_like(x)
    y = torch.empty_like(x)
    w = torch.randn(49, C, device=dev)
    bias = torch.randn(C, device=dev)
    dw = torch.zeros(49, C, device=dev)
    db = torch.zeros(C, device=dev)
    nb = x.numel() * 2
    fma = x.numel() * 49
    bench(f"dwconv7 fwd  {H}x{H}x{C}", lambda: call("lnx_dwconv7_fwd", x.data_ptr(), w.data_ptr(), bias.data_ptr(), None, y.data_ptr(), B, H, H, C, dt(x)), 2 * nb, flops=fma)
    bench(f"dwconv7 wgrad {H}x{H}x{C}", lambda: call("lnx_dwconv7_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), B, H, H, C, dt(x)), 2 * nb, flops=fma)

for (rows, C) in ((B * 3136, 96), (B * 784, 192), (B * 200, 384), (B * 53, 768)):
    x = torch.randn(rows, C, device=dev).bfloat16()
    g = torch.randn_like(x)
    res = torch.randn_like(x)
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    w = torch.randn(C, device=dev)
    b = torch.randn(C, device=dev)
    mean = torch.empty(rows, device=dev)
    rstd = torch.empty(rows, device=dev)
    dw = torch.zeros(C, device=dev)
    db = torch.zeros(C, device=dev)
    nb = rows * C * 2
    bench(f"layernorm fwd {rows}x{C}", lambda: call("lnx_layernorm_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), None, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, C, 1e-6, dt(x)), 2 * nb)
    bench(f"layernorm fwd+res {rows}x{C}", lambda: call("lnx_layernorm_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), res.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, C, 1e-6, dt(x)), 3 * nb)
    bench(f"layernorm bwd {rows}x{C}", lambda: call("lnx_layernorm_bwd", g.data_ptr(), x.data_ptr(), w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), None, dx.data_ptr(), dw.data_ptr(), db.data_ptr(), rows, C, dt(x)), 3 * nb)
