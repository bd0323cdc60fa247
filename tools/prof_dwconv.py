"""Depthwise 7x7 kernels at the bench shapes: forward / data gradient and weight gradient (CUDA events, L2 flushed).
LNX_DWCONV_KERNEL=2|3|4 selects the forward kernel, LNX_DWCONV_WGRAD=1|2 the weight-gradient kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as TF

from linnaeus_b200 import _lib

DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for B, H, C in ((256, 56, 96), (256, 28, 192), (32, 96, 256)):
    x = torch.randn(B, H, H, C, device=DEV).to(torch.bfloat16)
    g = torch.randn(B, H, H, C, device=DEV).to(torch.bfloat16)
    w = 0.2 * torch.randn(C, 1, 7, 7, device=DEV)
    w49 = w.reshape(C, 49).t().contiguous()
    bias = 0.1 * torch.randn(C, device=DEV)
    y = torch.empty_like(x)
    _lib.call("lnx_dwconv7_fwd", x.data_ptr(), w49.data_ptr(), 0, bias.data_ptr(), None, y.data_ptr(), B, H, H, C, 1)
    ref = TF.conv2d(x.float().permute(0, 3, 1, 2), w, bias, padding=3, groups=C).permute(0, 2, 3, 1)
    err = float((y.float() - ref).abs().max() / ref.abs().max())
    t_f = timeit(lambda: _lib.call("lnx_dwconv7_fwd", x.data_ptr(), w49.data_ptr(), 0, bias.data_ptr(), None, y.data_ptr(), B, H, H, C, 1))
    t_d = timeit(lambda: _lib.call("lnx_dwconv7_fwd", x.data_ptr(), w49.data_ptr(), 0, None, g.data_ptr(), y.data_ptr(), B, H, H, C, 1))
    dw = torch.zeros(49, C, device=DEV)
    db = torch.zeros(C, device=DEV)
    _lib.call("lnx_dwconv7_wgrad", x.data_ptr(), g.data_ptr(), dw.data_ptr(), 0, db.data_ptr(), B, H, H, C, 1)
    xr = x.float().permute(0, 3, 1, 2).contiguous()
    wr = w.clone().requires_grad_(True)
    TF.conv2d(xr, wr, None, padding=3, groups=C).backward(g.float().permute(0, 3, 1, 2))
    werr = float((dw.t().reshape(C, 1, 7, 7) - wr.grad).norm() / wr.grad.norm())
    t_w = timeit(lambda: _lib.call("lnx_dwconv7_wgrad", x.data_ptr(), g.data_ptr(), dw.data_ptr(), 0, db.data_ptr(), B, H, H, C, 1))
    fl = 2.0 * 49 * B * H * H * C
    print(f"{B}x{H}x{H}x{C}: fwd {t_f:.4f} ms ({fl / t_f / 1e9:.1f} TFLOP/s, err {err:.1e}) | dgrad+skip {t_d:.4f} ms | wgrad {t_w:.4f} ms "
          f"({fl / t_w / 1e9:.1f} TFLOP/s, relL2 {werr:.1e})", flush=True)
