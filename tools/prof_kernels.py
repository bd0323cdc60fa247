"""Every hot kernel of the training step at the bench shapes (B = 256, mFormerV1_sm), 3 launches each, in a fixed order.
Used for the ncu --set full capture of the round (tools/gpu_profiles.sh) and for a quick timing table."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import linnaeus_b200.functional as F
from linnaeus_b200 import _lib
from linnaeus_b200._lib import call, dt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
rows = []


def bench(name, fn, nbytes=None, flops=None):
    for _ in range(1):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    gbs = nbytes / ms / 1e6 if nbytes else 0.0
    tfs = flops / ms / 1e9 if flops else 0.0
    rows.append((name, ms, gbs, tfs))
    print(f"{name:44s} {ms:7.3f} ms  {gbs:7.0f} GB/s  {tfs:7.1f} TFLOP/s", flush=True)


bf = torch.bfloat16
# ---- ConvNeXt stage 0 pointwise pair (HBM bound)
M, C, Hd = B * 3136, 96, 384
xln = torch.randn(M, C, device=dev).to(bf)
w1 = (torch.randn(Hd, C, device=dev) / 10).to(bf)
w2 = (torch.randn(C, Hd, device=dev) / 10).to(bf)
b1 = torch.randn(Hd, device=dev)
gam = torch.rand(C, device=dev)
h = torch.empty(M, Hd, device=dev, dtype=bf)
dg = torch.empty_like(h)
dpre = torch.empty_like(h)
y = torch.empty(M, C, device=dev, dtype=bf)
dy = torch.randn(M, C, device=dev).to(bf)
res = torch.randn(M, C, device=dev).to(bf)
dw1 = torch.zeros(Hd, C, device=dev)
db1 = torch.zeros(Hd, device=dev)
dw2 = torch.zeros(C, Hd, device=dev)
db2 = torch.zeros(C, device=dev)
e = 2
bench("gemm pw1 fwd  (GELU, saves gelu')  802816x384x96", lambda: F.gemm(xln, w1, M, Hd, C, out=h, bias=b1, act=3, aux_out=dg), (M * C + 2 * M * Hd) * e, 2 * M * C * Hd)
bench("gemm pw2 fwd  (gamma, +residual)   802816x96x384", lambda: F.gemm(h, w2, M, C, Hd, out=y, bias=gam, residual=res, col_scale=gam), (M * Hd + 2 * M * C) * e, 2 * M * C * Hd)
bench("gemm dPre     (x saved gelu')      802816x384x96", lambda: F.gemm(dy, w2, M, Hd, C, b_trans=True, ldb=Hd, out=dpre, act=4, act_grad_in=dg), (M * C + 2 * M * Hd) * e, 2 * M * C * Hd)
bench("gemm dX                            802816x96x384", lambda: F.gemm(dpre, w1, M, C, Hd, b_trans=True, ldb=C, out=y), (M * Hd + M * C) * e, 2 * M * C * Hd)
bench("wgrad dW1+db1                      384x96 K=802816", lambda: F.wgrad(dpre, xln, out=dw1, db_out=db1), (M * Hd + M * C) * e, 2 * M * C * Hd)
bench("wgrad dW2+db2                      96x384 K=802816", lambda: F.wgrad(dy, h, out=dw2, db_out=db2), (M * Hd + M * C) * e, 2 * M * C * Hd)
# the single-kernel pointwise pair (what the train step launches at C = 96) -- same tensors
w2e = (w2.float() * gam[:, None]).to(bf)
bench("mlp fused fwd (pw1+GELU+pw2+gamma+res) 802816x96", lambda: F.mlp_fused_fwd(xln, w1, b1, w2, gam, gamma=gam, residual=res, out=y), 3 * M * C * e, 4 * M * C * Hd)
bench("mlp fused bwd (recompute; h, dPre, dX) 802816x96", lambda: call("lnx_mlp_fused_bwd", xln.data_ptr(), dy.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2e.data_ptr(), h.data_ptr(), dpre.data_ptr(), y.data_ptr(), M, C, Hd), (3 * M * C + 2 * M * Hd) * e, 6 * M * C * Hd)
bench("mlp fused bwd, dX only                 802816x96", lambda: call("lnx_mlp_fused_bwd", xln.data_ptr(), dy.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2e.data_ptr(), None, None, y.data_ptr(), M, C, Hd), 3 * M * C * e, 6 * M * C * Hd)
bench("mlp fused wgrad (dW1, db1, dW2 on chip) 802816x96", lambda: call("lnx_mlp_fused_wgrad", xln.data_ptr(), dy.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2e.data_ptr(), dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(), M, C, Hd), 2 * M * C * e, 8 * M * C * Hd)
del h, dg, dpre

# ---- transformer stage 2 GEMMs (tensor bound)
M2, D = B * 200, 384
x = torch.randn(M2, D, device=dev).to(bf)
wq = (torch.randn(3 * D, D, device=dev) / 20).to(bf)
wf1 = (torch.randn(4 * D, D, device=dev) / 20).to(bf)
wf2 = (torch.randn(D, 4 * D, device=dev) / 20).to(bf)
bq, bf1, bf2 = torch.randn(3 * D, device=dev), torch.randn(4 * D, device=dev), torch.randn(D, device=dev)
oq = torch.empty(M2, 3 * D, device=dev, dtype=bf)
hh = torch.empty(M2, 4 * D, device=dev, dtype=bf)
dgg = torch.empty_like(hh)
dpp = torch.empty_like(hh)
y2 = torch.empty(M2, D, device=dev, dtype=bf)
dwf = torch.zeros(4 * D, D, device=dev)
dbf = torch.zeros(4 * D, device=dev)
bench("gemm qkv                           51200x1152x384", lambda: F.gemm(x, wq, M2, 3 * D, D, out=oq, bias=bq), None, 2 * M2 * 3 * D * D)
frq = 0.3 * torch.randn(2, 6, 32, device=dev)
bench("gemm qkv + cos-RoPE epilogue       51200x1152x384", lambda: call("lnx_qkv_rope_gemm", x.data_ptr(), wq.data_ptr(), bq.data_ptr(), frq.data_ptr(), oq.data_ptr(), M2, D, D, 200, 4, 14, 0.125), None, 2 * M2 * 3 * D * D)
bench("gemm fc1 (GELU, saves gelu')       51200x1536x384", lambda: F.gemm(x, wf1, M2, 4 * D, D, out=hh, bias=bf1, act=3, aux_out=dgg), None, 2 * M2 * 4 * D * D)
bench("gemm fc2 (+residual)               51200x384x1536", lambda: F.gemm(hh, wf2, M2, D, 4 * D, out=y2, bias=bf2, residual=x), None, 2 * M2 * 4 * D * D)
bench("gemm dPre fc1                      51200x1536x384", lambda: F.gemm(y2, wf2, M2, 4 * D, D, b_trans=True, ldb=4 * D, out=dpp, act=4, act_grad_in=dgg), None, 2 * M2 * 4 * D * D)
bench("wgrad fc1                          1536x384 K=51200", lambda: F.wgrad(dpp, x, out=dwf, db_out=dbf), None, 2 * M2 * 4 * D * D)

# ---- depthwise 7x7, LayerNorm (stage 0)
H = 56
xi = torch.randn(B, H, H, C, device=dev).to(bf)
gi = torch.randn_like(xi)
yo = torch.empty_like(xi)
# stage 1 (28 x 28 x 192)
xi1 = torch.randn(B, 28, 28, 192, device=dev).to(bf)
yo1 = torch.empty_like(xi1)
w491 = torch.randn(49, 192, device=dev)
bc1 = torch.randn(192, device=dev)
dw491 = torch.zeros(49, 192, device=dev)
dbc1 = torch.zeros(192, device=dev)
w49 = torch.randn(49, C, device=dev)
bc = torch.randn(C, device=dev)
dw49 = torch.zeros(49, C, device=dev)
dbc = torch.zeros(C, device=dev)
nb = xi.numel() * 2
bench("dwconv7 fwd (tensor pipe)           256x56x56x96", lambda: call("lnx_dwconv7_fwd", xi.data_ptr(), w49.data_ptr(), 0, bc.data_ptr(), None, yo.data_ptr(), B, H, H, C, dt(xi)), 2 * nb, xi.numel() * 49 * 2)
bench("dwconv7 dgrad + skip (tensor pipe)  256x56x56x96", lambda: call("lnx_dwconv7_fwd", gi.data_ptr(), w49.data_ptr(), 0, None, xi.data_ptr(), yo.data_ptr(), B, H, H, C, dt(xi)), 3 * nb, xi.numel() * 49 * 2)
bench("dwconv7 wgrad (tensor pipe)         256x56x56x96", lambda: call("lnx_dwconv7_wgrad", xi.data_ptr(), gi.data_ptr(), dw49.data_ptr(), 0, dbc.data_ptr(), B, H, H, C, dt(xi)), 2 * nb, xi.numel() * 49 * 2)
_lib.load().lnx_dwconv7_set_impl(0)
bench("dwconv7 fwd (FFMA2 kernels)         256x56x56x96", lambda: call("lnx_dwconv7_fwd", xi.data_ptr(), w49.data_ptr(), 0, bc.data_ptr(), None, yo.data_ptr(), B, H, H, C, dt(xi)), 2 * nb, xi.numel() * 49 * 2)
bench("dwconv7 wgrad (FFMA2 kernels)       256x56x56x96", lambda: call("lnx_dwconv7_wgrad", xi.data_ptr(), gi.data_ptr(), dw49.data_ptr(), 0, dbc.data_ptr(), B, H, H, C, dt(xi)), 2 * nb, xi.numel() * 49 * 2)
_lib.load().lnx_dwconv7_set_impl(-1)
bench("dwconv7 fwd                        256x28x28x192", lambda: call("lnx_dwconv7_fwd", xi1.data_ptr(), w491.data_ptr(), 0, bc1.data_ptr(), None, yo1.data_ptr(), B, 28, 28, 192, dt(xi1)), 2 * xi1.numel() * 2, xi1.numel() * 49 * 2)
bench("dwconv7 wgrad                      256x28x28x192", lambda: call("lnx_dwconv7_wgrad", xi1.data_ptr(), yo1.data_ptr(), dw491.data_ptr(), 0, dbc1.data_ptr(), B, 28, 28, 192, dt(xi1)), 2 * xi1.numel() * 2, xi1.numel() * 49 * 2)
x2 = xi.view(-1, C)
g2 = gi.view(-1, C)
y2d = yo.view(-1, C)
lw, lb = torch.randn(C, device=dev), torch.randn(C, device=dev)
mean = torch.empty(M, device=dev)
rstd = torch.empty(M, device=dev)
dlw, dlb = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
bench("layernorm fwd                      802816x96", lambda: call("lnx_layernorm_fwd", x2.data_ptr(), lw.data_ptr(), lb.data_ptr(), None, y2d.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, C, 1e-6, dt(x2)), 2 * nb)
bench("layernorm bwd                      802816x96", lambda: call("lnx_layernorm_bwd", g2.data_ptr(), x2.data_ptr(), lw.data_ptr(), mean.data_ptr(), rstd.data_ptr(), None, y2d.data_ptr(), dlw.data_ptr(), dlb.data_ptr(), M, C, dt(x2)), 3 * nb)

# ---- attention + rope (stage 2)
heads, N, hd = 6, 200, 64
q = (torch.randn(B, heads, N, hd, device=dev) * 0.35).to(bf)
k = torch.randn(B, heads, N, hd, device=dev).to(bf)
v = torch.randn(B, heads, N, hd, device=dev).to(bf)
do = torch.randn(B, N, heads * hd, device=dev).to(bf)
out = torch.empty(B, N, heads * hd, device=dev, dtype=bf)
lse = torch.empty(B, heads, N, device=dev)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
ws = torch.empty(B * heads * N * (hd + 1) + 4, device=dev)
ab = B * heads * N * hd * 2
bench("attention fwd                      256x6x200x64", lambda: call("lnx_attn_fwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), B, heads, N, hd, 1, 0), 4 * ab, 2 * 2 * N * N * hd * B * heads)
bench("attention bwd                      256x6x200x64", lambda: call("lnx_attn_bwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), ws.data_ptr(), B, heads, N, hd, 1, 0), 8 * ab, 5 * 2 * N * N * hd * B * heads)
qkv = torch.randn(B, N, 3 * heads * hd, device=dev).to(bf)
bench("attention fwd from qkv (4-D maps)  256x6x200x64", lambda: call("lnx_attn_qkv_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, heads, N, hd, 1), 4 * ab, 2 * 2 * N * N * hd * B * heads)
bench("attention bwd from qkv             256x6x200x64", lambda: call("lnx_attn_qkv_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, heads, N, hd, 1), 8 * ab, 5 * 2 * N * N * hd * B * heads)
cos = torch.rand(196, heads, hd // 2, device=dev)
sin = torch.rand(196, heads, hd // 2, device=dev)
dqkv = torch.empty_like(qkv)
dth = torch.zeros(196, heads, hd // 2, device=dev)
bench("rope fwd (split + cos scale)       256x200x1152", lambda: call("lnx_rope_qk_fwd", qkv.data_ptr(), cos.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), B, N, heads, hd, 4, 0.125, 1), 2 * qkv.numel() * 2)
bench("rope bwd (scaled q / k input)      256x200x1152", lambda: call("lnx_rope_qk_bwd_scaled", dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), dqkv.data_ptr(), dth.data_ptr(), B, N, heads, hd, 4, 0.125, 1), (2 + 2 / 3) * qkv.numel() * 2)
bench("rope bwd                           256x200x1152", lambda: call("lnx_rope_qk_bwd", dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), dqkv.data_ptr(), dth.data_ptr(), B, N, heads, hd, 4, 0.125, 1), (2 + 2 / 3) * qkv.numel() * 2)

# ---- optimizer
n = 31_900_000
p_, g_, m_, v_ = (torch.randn(n, device=dev) for _ in range(4))
v_.abs_()
step_t = torch.ones(1, device=dev)
bench("adamw (flat, 31.9 M params)", lambda: None, None) if False else None
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        f.write("| kernel (B = 256 shapes) | ms | GB/s (algorithmic) | TFLOP/s |\n|---|---|---|---|\n")
        for name, ms, gbs, tfs in rows:
            f.write(f"| {name} | {ms:.3f} | {gbs:.0f} | {tfs:.1f} |\n")
