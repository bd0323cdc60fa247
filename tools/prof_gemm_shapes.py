"""The train step's data GEMMs (y = x W^T forward, dx = dy W backward) through lnx_gemm against torch.matmul (cuBLAS) on the same
bf16 operands: per-shape time, TFLOP/s and algorithmic TB/s.  Inputs are cycled through buffers larger than L2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200.functional as F

dev = "cuda"
shapes = [("s1 fc1", 200704, 768, 192), ("s1 fc2", 200704, 192, 768), ("s2 qkv", 50176, 1152, 384), ("s2 proj", 50176, 384, 384),
          ("s2 fc1", 50176, 1536, 384), ("s2 fc2", 50176, 384, 1536), ("s3 qkv", 12544, 2304, 768), ("s3 proj", 12544, 768, 768),
          ("s3 fc1", 12544, 3072, 768), ("s3 fc2", 12544, 768, 3072), ("ds 1", 200704, 192, 384), ("ds 2", 50176, 384, 768), ("ds 3", 12544, 768, 1536)]


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
    outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    t_l = timeit(lambda i: F.gemm(xs[i % nbuf], w, M, N, K, out=outs[i % nbuf]))
    t_c = timeit(lambda i: torch.matmul(xs[i % nbuf], w.t(), out=outs[i % nbuf]))
    # backward data GEMM: dx[M, K] = dy[M, N] W[N, K]  (B read untransposed)
    t_b = timeit(lambda i: F.gemm(outs[i % nbuf], w, M, K, N, b_trans=True, ldb=K, out=xs[i % nbuf]))
    t_cb = timeit(lambda i: torch.matmul(outs[i % nbuf], w, out=xs[i % nbuf]))
    fl = 2.0 * M * N * K
    by = (M * (N + K) + N * K) * 2
    print(f"{name:8s} M={M:6d} N={N:4d} K={K:4d}: fwd lnx {t_l * 1e3:6.1f} us ({fl / t_l / 1e9:5.0f} TF/s {by / t_l / 1e9:4.2f} TB/s) cublas {t_c * 1e3:6.1f} us"
          f" | dx lnx {t_b * 1e3:6.1f} us ({fl / t_b / 1e9:5.0f} TF/s) cublas {t_cb * 1e3:6.1f} us", flush=True)
