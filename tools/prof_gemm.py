"""Micro-driver for ncu: a handful of launches of the dominant GEMM shapes (no training loop)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import linnaeus_b200.functional as F

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = "cuda"
M, K, N = B * 3136, 96, 384
a = torch.randn(M, K, device=dev).bfloat16()
w = (torch.randn(N, K, device=dev) / 10).bfloat16()
b = torch.randn(N, device=dev)
o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
aux = torch.empty_like(o)
dy = torch.randn(M, K, device=dev).bfloat16()
w2 = (torch.randn(K, N, device=dev) / 10).bfloat16()   # pwconv2 weight [C, 4C]
dpre = torch.empty_like(o)
cs = torch.zeros(N, device=dev)
res = torch.randn(M, K, device=dev).bfloat16()
gam = torch.rand(K, device=dev)
y = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
for it in range(3):
    ev[0].record()
    F.gemm(a, w, M, N, K, out=o, bias=b, act=1, aux_out=aux)                       # pwconv1 fwd
    ev[1].record()
    F.gemm(o, w2, M, K, N, out=y, bias=gam, residual=res, col_scale=gam)             # pwconv2 fwd
    ev[2].record()
    F.gemm(dy, w2, M, N, K, b_trans=True, ldb=N, out=dpre, act=1, act_grad_in=aux, colsum_out=cs)  # dPre
    ev[3].record()
    F.gemm(dpre, w, M, K, N, b_trans=True, ldb=K, out=y)                            # dX
    ev[4].record()
    F.wgrad(dpre, a)                                                                # dW1
    ev[5].record()
    F.wgrad(dy, o)                                                                  # dW2
    ev[6].record()
torch.cuda.synchronize()
names = ["pw1_fwd", "pw2_fwd", "dpre", "dx", "dw1", "dw2"]
bytes_ = [(M*K + 2*M*N)*2, (M*N + 2*M*K)*2, (M*K + 2*M*N)*2, (M*N + M*K)*2, (M*N + M*K)*2, (M*N + M*K)*2]
for i, n in enumerate(names):
    ms = ev[i].elapsed_time(ev[i+1])
    print(f"{n:8s} {ms:7.3f} ms  {bytes_[i]/ms/1e6:7.0f} GB/s")
