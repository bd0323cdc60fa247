"""Launch one kernel a few times for ncu.  python tools/prof_one.py dwconv_fwd|dwconv_wgrad|mlp_bwd|attn_bwd"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200.functional as F
from linnaeus_b200 import _lib

DEV = "cuda"
what = sys.argv[1]
B, H, C = 256, 56, 96
x = torch.randn(B, H, H, C, device=DEV).to(torch.bfloat16)
g = torch.randn(B, H, H, C, device=DEV).to(torch.bfloat16)
w49 = 0.2 * torch.randn(49, C, device=DEV)
bias = 0.1 * torch.randn(C, device=DEV)
y = torch.empty_like(x)
dw = torch.zeros(49, C, device=DEV)
db = torch.zeros(C, device=DEV)
for _ in range(4):
    if what == "dwconv_fwd":
        _lib.call("lnx_dwconv7_fwd", x.data_ptr(), w49.data_ptr(), 0, bias.data_ptr(), None, y.data_ptr(), B, H, H, C, 1)
    elif what == "dwconv_wgrad":
        _lib.call("lnx_dwconv7_wgrad", x.data_ptr(), g.data_ptr(), dw.data_ptr(), 0, db.data_ptr(), B, H, H, C, 1)
    elif what == "mlp_bwd":
        M = B * H * H
        w1 = (torch.randn(4 * C, C, device=DEV) * C ** -0.5).to(torch.bfloat16)
        w2 = (torch.randn(C, 4 * C, device=DEV) * (4 * C) ** -0.5).to(torch.bfloat16)
        b1 = torch.randn(4 * C, device=DEV) * 0.1
        F.mlp_fused_bwd(x.view(M, C), g.view(M, C), w1, b1, w2)
torch.cuda.synchronize()
