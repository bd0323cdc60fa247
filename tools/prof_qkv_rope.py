"""Fused qkv projection + cos-RoPE + attention (qkv_rope_attention) against the split path it replaces (Linear -> rope_qk_fwd ->
attention), forward and forward + backward, at the train step's stage-2 / stage-3 shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200.functional as F
from linnaeus_b200.flat import weights_changed  # noqa: F401

dev = "cuda"


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, B, heads, H, W, n_extra, K in (("stage 2", 256, 6, 14, 14, 4, 384), ("stage 3", 256, 12, 7, 7, 4, 768)):
    D = heads * 64
    N = H * W + n_extra
    x = torch.randn(B, N, K, device=dev).bfloat16().requires_grad_(True)
    w = (torch.randn(3 * D, K, device=dev) * K ** -0.5).requires_grad_(True)
    wc = w.detach().bfloat16()
    b = (0.1 * torch.randn(3 * D, device=dev)).requires_grad_(True)
    fr = (0.3 * torch.randn(2, heads, 32, device=dev)).requires_grad_(True)
    for p in (w, b, fr):
        p.grad = torch.zeros_like(p)
    g = torch.randn(B, N, D, device=dev).bfloat16()

    def fused():
        return F.qkv_rope_attention(x, w, b, wc, fr, H, W, heads, n_extra)

    def split():
        return F.rope_attention(F.linear(x, w, b, weight_c=wc), fr, H, W, heads, n_extra)

    res = {}
    for label, fn in (("fused", fused), ("split", split)):
        with torch.no_grad():
            t_f = timeit(fn)

        def fb():
            x.grad = None
            fn().backward(g)

        t_fb = timeit(fb)
        res[label] = (t_f, t_fb)
    print(f"{name}: forward fused {res['fused'][0] * 1e3:.1f} us | split {res['split'][0] * 1e3:.1f} us;  forward + backward fused "
          f"{res['fused'][1] * 1e3:.1f} us | split {res['split'][1] * 1e3:.1f} us", flush=True)

# ---- the pieces, stage 2
from linnaeus_b200.functional import call, dt, ptr  # noqa: E402

B, heads, H, W, n_extra, K = 256, 6, 14, 14, 4, 384
D, N, hd = heads * 64, H * W + n_extra, 64
x2 = torch.randn(B * N, K, device=dev).bfloat16()
wc = (torch.randn(3 * D, K, device=dev) * K ** -0.5).bfloat16()
bias = 0.1 * torch.randn(3 * D, device=dev)
cos = torch.rand(H * W, heads, 32, device=dev)
frq = 0.3 * torch.randn(2, heads, 32, device=dev)
qkv = torch.empty(B, N, 3 * D, device=dev, dtype=torch.bfloat16)
t_plain = timeit(lambda: F.gemm(x2, wc, B * N, 3 * D, K, bias=bias, out=qkv.view(B * N, 3 * D)))
t_rope = timeit(lambda: call("lnx_qkv_rope_gemm", x2.data_ptr(), wc.data_ptr(), ptr(bias), frq.data_ptr(), qkv.data_ptr(), B * N, D, K, N, n_extra, W, 0.125))
qkv_h = torch.empty(3, B, heads, N, hd, device=dev, dtype=torch.bfloat16)
t_rk = timeit(lambda: call("lnx_rope_qk_fwd", qkv.data_ptr(), cos.data_ptr(), qkv_h[0].data_ptr(), qkv_h[1].data_ptr(), qkv_h[2].data_ptr(), B, N, heads, hd, n_extra, 0.125, dt(qkv)))
out = torch.empty(B, N, D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, heads, N, device=dev)
t_a3 = timeit(lambda: call("lnx_attn_fwd", qkv_h[0].data_ptr(), qkv_h[1].data_ptr(), qkv_h[2].data_ptr(), out.data_ptr(), lse.data_ptr(), B, heads, N, hd, dt(qkv), 0))
t_a4 = timeit(lambda: call("lnx_attn_qkv_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, heads, N, hd, dt(qkv)))
dout = torch.randn(B, N, D, device=dev).bfloat16()
dq = torch.empty_like(qkv_h)
delta = torch.empty(B * heads * N * (hd + 1) + 4, device=dev)
t_b3 = timeit(lambda: call("lnx_attn_bwd", qkv_h[0].data_ptr(), qkv_h[1].data_ptr(), qkv_h[2].data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dq[0].data_ptr(), dq[1].data_ptr(), dq[2].data_ptr(), delta.data_ptr(), B, heads, N, hd, dt(qkv), 0))
t_b4 = timeit(lambda: call("lnx_attn_qkv_bwd", qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dq[0].data_ptr(), dq[1].data_ptr(), dq[2].data_ptr(), B, heads, N, hd, dt(qkv)))
print(f"qkv GEMM plain {t_plain * 1e3:.1f} us | with cos epilogue {t_rope * 1e3:.1f} us | rope_qk_fwd {t_rk * 1e3:.1f} us")
print(f"attention fwd head-major {t_a3 * 1e3:.1f} us | from qkv (4-D maps) {t_a4 * 1e3:.1f} us;  bwd head-major {t_b3 * 1e3:.1f} us | from qkv {t_b4 * 1e3:.1f} us")
