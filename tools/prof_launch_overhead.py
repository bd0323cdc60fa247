"""Fixed cost per kernel inside a CUDA graph: N back-to-back dependent launches of a kernel on a tiny problem, replayed.
Separates the launch + prologue + tail of the tcgen05 kernels (tensor-map prefetch, barrier init, TMEM alloc / dealloc) from a trivial kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200.functional as F
from linnaeus_b200._lib import call, dt

dev = "cuda"
N = 200


def graph_time(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 / N * 1e3


x = torch.randn(128, 64, device=dev).bfloat16()
w = torch.randn(128, 64, device=dev).bfloat16()
o = torch.empty(128, 128, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(128, 64, device=dev)
t = torch.randn(4096, device=dev)
t2 = torch.empty(4096, device=dev, dtype=torch.bfloat16)
ln_w, ln_b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
y = torch.empty_like(x)
mean, rstd = torch.empty(128, device=dev), torch.empty(128, device=dev)
print(f"torch elementwise (t.mul_):            {graph_time(lambda: t.mul_(1.0001)):.2f} us per launch")
print(f"lnx_layernorm_fwd 128x64 (SIMT):       {graph_time(lambda: call('lnx_layernorm_fwd', x.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), None, y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), 128, 64, 1e-5, dt(x))):.2f} us per launch")
print(f"lnx_gemm 128x128x64 (tcgen05, tc2):    {graph_time(lambda: F.gemm(x, w, 128, 128, 64, out=o)):.2f} us per launch")
print(f"lnx_wgrad 128x64 K=128 (tcgen05):      {graph_time(lambda: F.wgrad(o[:, :128], x, out=dw)):.2f} us per launch")
