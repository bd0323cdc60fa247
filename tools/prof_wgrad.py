import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import linnaeus_b200.functional as F
dev = "cuda"
Mr, No, Ki = 256 * 3136, 384, 96
dyw = torch.randn(Mr, No, device=dev).bfloat16()
xw = torch.randn(Mr, Ki, device=dev).bfloat16()
dww = torch.zeros(No, Ki, device=dev)
for _ in range(3):
    F.gemm(dyw, xw, No, Ki, Mr, a_trans=True, b_trans=True, lda=No, ldb=Ki, out=dww, accumulate=True)
    F.wgrad(dyw, xw, out=dww)
torch.cuda.synchronize()
