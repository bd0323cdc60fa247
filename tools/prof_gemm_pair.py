"""lnx_gemm_pair (CTA pairs, tcgen05.mma.cta_group::2) against lnx_gemm (one CTA per tile) and cuBLAS on the transformer-stage shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200.functional as F
from linnaeus_b200._lib import call

dev = "cuda"
shapes = [("tiny", 256, 128, 64), ("small", 1000, 256, 192), ("s2 qkv", 50176, 1152, 384), ("s2 proj", 50176, 384, 384), ("s2 fc1", 50176, 1536, 384),
          ("s2 fc2", 50176, 384, 1536), ("s3 qkv", 12544, 2304, 768), ("s3 proj", 12544, 768, 768), ("s3 fc1", 12544, 3072, 768), ("s3 fc2", 12544, 768, 3072)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]


def timeit(fn, iters=20):
    for _ in range(3):
        fn(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, M, N, K in shapes:
    nbuf = max(2, int(300e6 // (M * (N + K) * 2)) + 1)
    xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(nbuf)]
    w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    call("lnx_gemm_pair", xs[0].data_ptr(), w.data_ptr(), bias.data_ptr(), outs[0].data_ptr(), M, N, K)
    torch.cuda.synchronize()
    ref = xs[0].float() @ w.float().t() + bias
    err = float((outs[0].float() - ref).abs().max() / ref.abs().max())
    t_p = timeit(lambda i: call("lnx_gemm_pair", xs[i % nbuf].data_ptr(), w.data_ptr(), bias.data_ptr(), outs[i % nbuf].data_ptr(), M, N, K))
    t_l = timeit(lambda i: F.gemm(xs[i % nbuf], w, M, N, K, out=outs[i % nbuf], bias=bias))
    t_c = timeit(lambda i: torch.addmm(bias.bfloat16(), xs[i % nbuf], w.t(), out=outs[i % nbuf]))
    fl = 2.0 * M * N * K
    print(f"{name:8s} M={M:6d} N={N:4d} K={K:4d}: pair {t_p * 1e3:6.1f} us ({fl / t_p / 1e9:5.0f} TF/s) | 1-CTA {t_l * 1e3:6.1f} us ({fl / t_l / 1e9:5.0f}) | cuBLAS {t_c * 1e3:6.1f} us"
          f"  max rel err {err:.1e}", flush=True)
