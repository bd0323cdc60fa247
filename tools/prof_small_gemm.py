"""Micro-benchmark of the metadata-head first Linear (K = 2 / 3 / 10, fp32 input) forward + backward."""
import sys
import torch
sys.path.insert(0, ".")
from linnaeus_b200 import functional as F

torch.manual_seed(0)
dev = "cuda"
meta = torch.randn(256, 15, device=dev)
for K, N in ((2, 384), (3, 384), (10, 384), (2, 768), (10, 768)):
    w = torch.randn(N, K, device=dev, requires_grad=True)
    b = torch.randn(N, device=dev, requires_grad=True)
    wc = F.compute_copy(w, torch.bfloat16)
    m = meta[:, :K]

    def fwd():
        return F.linear(m, w, b, weight_c=wc, act="relu", x_ld=15)

    y = fwd()
    g = torch.randn_like(y)
    for _ in range(3):
        fwd().backward(g)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e[0].record()
    for _ in range(50):
        y = fwd()
    e[1].record()
    for _ in range(50):
        y.backward(g, retain_graph=True)
    e[2].record()
    torch.cuda.synchronize()
    print(f"K={K} N={N}: fwd {e[0].elapsed_time(e[1]) / 50 * 1e3:.1f} us  bwd {e[1].elapsed_time(e[2]) / 50 * 1e3:.1f} us")
