"""Fused ConvNeXt pointwise pair: parity against fp32 torch on the same bf16 inputs, then timing at the bench shapes
next to the two-GEMM path.  python tools/prof_mlp_fused.py [--no-time]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as TF

import linnaeus_b200.functional as F

DEV = "cuda"


def ref(x, w1, b1, w2, b2, gamma, rs, rpg, res):
    h = TF.gelu(x.float() @ w1.float().t() + (b1 if b1 is not None else 0.0)).to(torch.bfloat16).float()
    y = h @ w2.float().t() + (b2 if b2 is not None else 0.0)
    if gamma is not None:
        y = y * gamma
    if rs is not None:
        y = y * rs.repeat_interleave(rpg)[: y.shape[0], None]
    if res is not None:
        y = y + res.float()
    return y


def make(M, C, seed=0, gamma_scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    H = 4 * C
    x = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(H, C, device=DEV, generator=g) * C ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, H, device=DEV, generator=g) * H ** -0.5).to(torch.bfloat16)
    b1 = torch.randn(H, device=DEV, generator=g) * 0.5
    b2 = torch.randn(C, device=DEV, generator=g) * 0.5
    gamma = (torch.rand(C, device=DEV, generator=g) + 0.5) * gamma_scale
    res = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    return x, w1, b1, w2, b2, gamma, res


def check(M, C, use_gamma=True, use_res=True, use_rs=False, use_bias=True):
    x, w1, b1, w2, b2, gamma, res = make(M, C)
    rpg = 49
    rs = None
    if use_rs:
        n = (M + rpg - 1) // rpg
        rs = torch.floor(0.8 + torch.rand(n, device=DEV)) / 0.8
    args = (x, w1, b1 if use_bias else None, w2, b2 if use_bias else None, gamma if use_gamma else None, rs, rpg, res if use_res else None)
    y = F.mlp_fused_fwd(x, w1, args[2], w2, args[4], gamma=args[5], row_scale=rs, rows_per_group=rpg if use_rs else 0, residual=args[8])
    torch.cuda.synchronize()
    r = ref(*args)
    err = float((y.float() - r).abs().max() / r.abs().max())
    print(f"M={M:7d} C={C:3d} gamma={int(use_gamma)} res={int(use_res)} rs={int(use_rs)} bias={int(use_bias)}  rel err {err:.3e}", flush=True)
    return err


def gelu_grad(pre):
    return 0.5 * (1 + torch.erf(pre * 0.7071067811865476)) + pre * torch.exp(-0.5 * pre * pre) * 0.3989422804014327


def check_bwd(M, C=96):
    x, w1, b1, w2, b2, gamma, res = make(M, C)
    g = torch.Generator(device=DEV).manual_seed(7)
    dy = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    w2e = (w2.float() * gamma[:, None]).to(torch.bfloat16)
    h, dpre, dx = F.mlp_fused_bwd(x, dy, w1, b1, w2e)
    torch.cuda.synchronize()
    pre = x.float() @ w1.float().t() + b1
    h_ref = TF.gelu(pre)
    dpre_ref = (dy.float() @ w2e.float()) * gelu_grad(pre)
    dx_ref = dpre_ref.to(torch.bfloat16).float() @ w1.float()
    errs = [float((a.float() - r).abs().max() / r.abs().max()) for a, r in ((h, h_ref), (dpre, dpre_ref), (dx, dx_ref))]
    print(f"bwd M={M:7d} C={C}: rel err h {errs[0]:.3e} dpre {errs[1]:.3e} dx {errs[2]:.3e}", flush=True)
    return max(errs)


def check_wgrad(M, C=96):
    x, w1, b1, w2, b2, gamma, res = make(M, C)
    g = torch.Generator(device=DEV).manual_seed(11)
    dy = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    w2e = (w2.float() * gamma[:, None]).to(torch.bfloat16)
    H = 4 * C
    dw1 = torch.zeros(H, C, device=DEV)
    db1 = torch.zeros(H, device=DEV)
    dw2 = torch.zeros(C, H, device=DEV)
    db2 = torch.zeros(C, device=DEV)
    F.mlp_fused_wgrad(x, dy, w1, b1, w2e, dw1, db1, dw2, db2)
    _, _, dx = F.mlp_fused_bwd(x, dy, w1, b1, w2e, store_hidden=False)
    torch.cuda.synchronize()
    pre = x.float() @ w1.float().t() + b1
    h = TF.gelu(pre).to(torch.bfloat16).float()
    dpre = ((dy.float() @ w2e.float()) * gelu_grad(pre)).to(torch.bfloat16).float()
    refs = (dpre.t() @ x.float(), dpre.sum(0), dy.float().t() @ h, dy.float().sum(0), dpre @ w1.float())
    errs = [float((a - r).norm() / r.norm()) for a, r in zip((dw1, db1, dw2, db2, dx.float()), refs)]
    print(f"wgrad M={M:7d}: relL2 dW1 {errs[0]:.2e} db1 {errs[1]:.2e} dW2raw {errs[2]:.2e} db2raw {errs[3]:.2e} | dX-only {errs[4]:.2e}", flush=True)
    return max(errs)


def timeit(fn, iters=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def bench_only(C, M, iters=5):
    x, w1, b1, w2, b2, gamma, res = make(M, C)
    out = torch.empty_like(x)
    for _ in range(iters):
        F.mlp_fused_fwd(x, w1, b1, w2, b2, gamma=gamma, residual=res, out=out)
    torch.cuda.synchronize()


def time_fwd(C):
    M = 802816 if C == 96 else 200704
    x, w1, b1, w2, b2, gamma, res = make(M, C)
    out = torch.empty_like(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    import os
    if os.environ.get("NORES"):
        res = None
    t = timeit(lambda: F.mlp_fused_fwd(x, w1, b1, w2, b2, gamma=gamma, residual=res, out=out), flush=flush)
    print(f"C={C} M={M} fused fwd {t:.4f} ms", flush=True)


def main():
    if "--time-fwd" in sys.argv:
        time_fwd(int(sys.argv[sys.argv.index("--time-fwd") + 1]))
        return
    if "--bench-only" in sys.argv:
        C = int(sys.argv[sys.argv.index("--bench-only") + 1])
        bench_only(C, 802816 if C == 96 else 200704)
        return
    worst = 0.0
    for C in (96, 192):
        for M in (128, 300, 4096 + 17, 50000):
            worst = max(worst, check(M, C))
        worst = max(worst, check(1000, C, use_gamma=False, use_res=False, use_bias=False))
        worst = max(worst, check(3000, C, use_rs=True))
    print("worst rel err", worst)
    if "--no-time" in sys.argv:
        return
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for C, M in ((96, 802816), (192, 200704)):
        x, w1, b1, w2, b2, gamma, res = make(M, C)
        worst = max(worst, check(M, C))
        out = torch.empty_like(x)
        t_f = timeit(lambda: F.mlp_fused_fwd(x, w1, b1, w2, b2, gamma=gamma, residual=res, out=out), flush=flush)

        def two():
            h = F.gemm(x, w1, M, 4 * C, C, bias=b1, act=F.ACT_GELU)
            F.gemm(h, w2, M, C, 4 * C, bias=b2, residual=res, col_scale=gamma, out=out)

        t_2 = timeit(two, flush=flush)
        bytes_alg = 3 * M * C * 2
        flops = 4.0 * M * C * 4 * C
        print(f"C={C} M={M}: fused {t_f:.4f} ms ({bytes_alg / t_f / 1e6:.0f} GB/s algorithmic, {flops / t_f / 1e9:.0f} TFLOP/s) | two GEMMs {t_2:.4f} ms",
              flush=True)
    print("worst rel err (incl. bench shapes)", worst)
    wb = 0.0
    for M in (128, 300, 4113, 50000, 802816):
        wb = max(wb, check_bwd(M))
    print("worst bwd rel err", wb)
    ww = 0.0
    for M in (128, 300, 4113, 50000, 802816):
        ww = max(ww, check_wgrad(M))
    print("worst wgrad relL2", ww)
    M, C = 802816, 96
    x, w1, b1, w2, b2, gamma, res = make(M, C)
    dy = torch.randn(M, C, device=DEV).to(torch.bfloat16)
    w2e = (w2.float() * gamma[:, None]).to(torch.bfloat16)
    t_b = timeit(lambda: F.mlp_fused_bwd(x, dy, w1, b1, w2e), flush=flush)
    pre_dg = torch.empty(M, 4 * C, dtype=torch.bfloat16, device=DEV)

    def two_b():
        dpre = F.gemm(dy, w2e, M, 4 * C, C, b_trans=True, ldb=4 * C, act=F.ACT_MUL, act_grad_in=pre_dg)
        F.gemm(dpre, w1, M, C, 4 * C, b_trans=True, ldb=C)

    t_2b = timeit(two_b, flush=flush)
    dw1 = torch.zeros(4 * C, C, device=DEV); db1 = torch.zeros(4 * C, device=DEV); dw2 = torch.zeros(C, 4 * C, device=DEV); db2 = torch.zeros(C, device=DEV)
    t_w = timeit(lambda: F.mlp_fused_wgrad(x, dy, w1, b1, w2e, dw1, db1, dw2, db2), flush=flush)
    t_dx = timeit(lambda: F.mlp_fused_bwd(x, dy, w1, b1, w2e, store_hidden=False), flush=flush)
    print(f"on-chip weight gradients {t_w:.4f} ms | dX-only data kernel {t_dx:.4f} ms | sum {t_w + t_dx:.4f} ms (vs data kernel with h / dPre stores + 2 lnx_wgrad)", flush=True)
    print(f"bwd data path C={C} M={M}: fused (recompute; writes h, dpre, dx) {t_b:.4f} ms ({11 * M * C * 2 / t_b / 1e6:.0f} GB/s) | dPre + dX GEMMs {t_2b:.4f} ms",
          flush=True)


if __name__ == "__main__":
    main()
