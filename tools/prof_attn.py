"""Attention forward / backward at the bench shapes (CUDA events, warm)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from linnaeus_b200._lib import call

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for heads, N in ((6, 200), (12, 53)):
    hd = 64
    q = (torch.randn(B, heads, N, hd, device=dev) * 0.35).bfloat16()
    k = torch.randn(B, heads, N, hd, device=dev).bfloat16()
    v = torch.randn(B, heads, N, hd, device=dev).bfloat16()
    do = torch.randn(B, N, heads * hd, device=dev).bfloat16()
    out = torch.empty(B, N, heads * hd, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, heads, N, device=dev)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ws = torch.empty(B * heads * N * (hd + 1) + 4, device=dev)
    def fwd():
        call("lnx_attn_fwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), B, heads, N, hd, 1, 0)
    def bwd():
        call("lnx_attn_bwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(),
             dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), B, heads, N, hd, 1, 0)
    for name, fn, nmm in (("fwd", fwd, 2), ("bwd", bwd, 5)):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = nmm * 2 * N * N * hd * B * heads
        by = (4 if name == "fwd" else 8) * B * heads * N * hd * 2
        print(f"attn {name} B={B} h={heads} N={N}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  {by / ms / 1e6:.0f} GB/s")
