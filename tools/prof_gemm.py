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

# ---- transformer-stage shapes (tensor-bound): report TFLOP/s
def bench(name, fn, flops, n=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:28s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s")

for (M2, D) in ((B * 200, 384), (B * 53, 768)):
    x = torch.randn(M2, D, device=dev).bfloat16()
    wq = (torch.randn(3 * D, D, device=dev) / 20).bfloat16()
    w1 = (torch.randn(4 * D, D, device=dev) / 20).bfloat16()
    w2_ = (torch.randn(D, 4 * D, device=dev) / 20).bfloat16()
    bq = torch.randn(3 * D, device=dev); b1 = torch.randn(4 * D, device=dev); b2 = torch.randn(D, device=dev)
    oq = torch.empty(M2, 3 * D, device=dev, dtype=torch.bfloat16)
    h = torch.empty(M2, 4 * D, device=dev, dtype=torch.bfloat16); pre = torch.empty_like(h); dpre2 = torch.empty_like(h)
    y2 = torch.empty(M2, D, device=dev, dtype=torch.bfloat16)
    cs2 = torch.zeros(4 * D, device=dev)
    bench(f"qkv  {M2}x{3*D}x{D}", lambda: F.gemm(x, wq, M2, 3 * D, D, out=oq, bias=bq), 2 * M2 * 3 * D * D)
    bench(f"fc1  {M2}x{4*D}x{D} gelu+aux", lambda: F.gemm(x, w1, M2, 4 * D, D, out=h, bias=b1, act=1, aux_out=pre), 2 * M2 * 4 * D * D)
    bench(f"fc2  {M2}x{D}x{4*D} +res", lambda: F.gemm(h, w2_, M2, D, 4 * D, out=y2, bias=b2, residual=x), 2 * M2 * 4 * D * D)
    bench(f"dpre {M2}x{4*D}x{D} act'", lambda: F.gemm(y2, w2_, M2, 4 * D, D, b_trans=True, ldb=4 * D, out=dpre2, act=1, act_grad_in=pre, colsum_out=cs2), 2 * M2 * 4 * D * D)
    bench(f"dx   {M2}x{D}x{4*D}", lambda: F.gemm(dpre2, w1, M2, D, 4 * D, b_trans=True, ldb=D, out=y2), 2 * M2 * 4 * D * D)
    bench(f"dw1  {4*D}x{D}x{M2}", lambda: F.wgrad(dpre2, x), 2 * M2 * 4 * D * D)

# ---- dPre variants at the stage-0 shape (which part of the epilogue costs what)
bench("dpre gelu' + colsum", lambda: F.gemm(dy, w2, M, N, K, b_trans=True, ldb=N, out=dpre, act=1, act_grad_in=aux, colsum_out=cs), 2 * M * N * K)
bench("dpre gelu' no colsum", lambda: F.gemm(dy, w2, M, N, K, b_trans=True, ldb=N, out=dpre, act=1, act_grad_in=aux), 2 * M * N * K)
bench("dpre relu' + colsum", lambda: F.gemm(dy, w2, M, N, K, b_trans=True, ldb=N, out=dpre, act=2, act_grad_in=aux, colsum_out=cs), 2 * M * N * K)
bench("dpre relu' no colsum", lambda: F.gemm(dy, w2, M, N, K, b_trans=True, ldb=N, out=dpre, act=2, act_grad_in=aux), 2 * M * N * K)
bench("plain b_trans N=384", lambda: F.gemm(dy, w2, M, N, K, b_trans=True, ldb=N, out=dpre), 2 * M * N * K)
w2t = w2.t().contiguous()
bench("plain K-major N=384", lambda: F.gemm(dy, w2t, M, N, K, out=dpre), 2 * M * N * K)
bench("residual K-major N=384", lambda: F.gemm(dy, w2t, M, N, K, out=dpre, residual=aux), 2 * M * N * K)

# ---- weight-gradient kernels: generic split-K GEMM (v1) vs lnx_wgrad with / without the fused bias gradient
for (Mr, No, Ki) in ((B * 3136, 384, 96), (B * 3136, 96, 384), (B * 200, 1536, 384), (B * 200, 384, 1536), (B * 53, 3072, 768)):
    dyw = torch.randn(Mr, No, device=dev).bfloat16()
    xw = torch.randn(Mr, Ki, device=dev).bfloat16()
    dww = torch.zeros(No, Ki, device=dev)
    dbw = torch.zeros(No, device=dev)
    fl = 2 * Mr * No * Ki
    bench(f"wgrad v1   {No}x{Ki} K={Mr}", lambda: F.gemm(dyw, xw, No, Ki, Mr, a_trans=True, b_trans=True, lda=No, ldb=Ki, out=dww, accumulate=True), fl)
    bench(f"lnx_wgrad  {No}x{Ki} K={Mr}", lambda: F.wgrad(dyw, xw, out=dww), fl)
    bench(f"lnx_wgrad+db {No}x{Ki} K={Mr}", lambda: F.wgrad(dyw, xw, out=dww, db_out=dbw), fl)
