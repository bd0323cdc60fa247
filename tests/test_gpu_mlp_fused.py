"""GPU: the fused ConvNeXt pointwise pair (lnx_mlp_fused_fwd / lnx_mlp_fused_bwd) against fp32 torch on the same bf16 inputs -
R/models/blocks/convnext.py:79-86 and its autograd - from ragged small M up to the bench shape (M = 256 * 56 * 56 = 802 816),
and the autograd wiring (_Mlp2) against the two-GEMM path."""
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 8e-3  # max-abs error relative to the max-abs reference value: one bf16 rounding of a K = 384 dot product is ~2-4e-3


def _make(M, C, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    H = 4 * C
    x = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(H, C, device=DEV, generator=g) * C ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, H, device=DEV, generator=g) * H ** -0.5).to(torch.bfloat16)
    b1 = torch.randn(H, device=DEV, generator=g) * 0.5
    b2 = torch.randn(C, device=DEV, generator=g) * 0.5
    gamma = torch.rand(C, device=DEV, generator=g) + 0.5
    res = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    dy = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    return x, w1, b1, w2, b2, gamma, res, dy


def _rel(a, r):
    return float((a.float() - r).abs().max() / r.abs().max())


def _gelu_grad(pre):
    return 0.5 * (1 + torch.erf(pre * 0.7071067811865476)) + pre * torch.exp(-0.5 * pre * pre) * 0.3989422804014327


@pytest.mark.parametrize("C", [96, 192])
@pytest.mark.parametrize("M", [1, 128, 300, 4113, 50000])
def test_fused_forward_matches_fp32_torch(M, C):
    import linnaeus_b200.functional as F

    x, w1, b1, w2, b2, gamma, res, _ = _make(M, C)
    y = F.mlp_fused_fwd(x, w1, b1, w2, b2, gamma=gamma, residual=res)
    h = TF.gelu(x.float() @ w1.float().t() + b1).to(torch.bfloat16).float()
    ref = (h @ w2.float().t() + b2) * gamma + res.float()
    assert _rel(y, ref) < TOL


@pytest.mark.parametrize("C", [96, 192])
def test_fused_forward_options(C):
    import linnaeus_b200.functional as F

    M, rpg = 3000, 49
    x, w1, b1, w2, b2, gamma, res, _ = _make(M, C, seed=3)
    rs = torch.floor(0.8 + torch.rand((M + rpg - 1) // rpg, device=DEV)) / 0.8
    h = TF.gelu(x.float() @ w1.float().t() + b1).to(torch.bfloat16).float()
    y = F.mlp_fused_fwd(x, w1, b1, w2, b2, gamma=gamma, row_scale=rs, rows_per_group=rpg, residual=res)
    ref = (h @ w2.float().t() + b2) * gamma * rs.repeat_interleave(rpg)[:M, None] + res.float()
    assert _rel(y, ref) < TOL
    y = F.mlp_fused_fwd(x, w1, None, w2, None)  # no bias, no layer scale, no residual
    ref = TF.gelu(x.float() @ w1.float().t()).to(torch.bfloat16).float() @ w2.float().t()
    assert _rel(y, ref) < TOL


@pytest.mark.parametrize("M", [1, 128, 300, 4113, 50000])
def test_fused_backward_matches_fp32_torch(M):
    import linnaeus_b200.functional as F

    C = 96
    x, w1, b1, w2, b2, gamma, res, dy = _make(M, C, seed=1)
    w2e = (w2.float() * gamma[:, None]).to(torch.bfloat16)
    h, dpre, dx = F.mlp_fused_bwd(x, dy, w1, b1, w2e)
    pre = x.float() @ w1.float().t() + b1
    dpre_ref = (dy.float() @ w2e.float()) * _gelu_grad(pre)
    assert _rel(h, TF.gelu(pre)) < TOL
    assert _rel(dpre, dpre_ref) < TOL
    assert _rel(dx, dpre_ref.to(torch.bfloat16).float() @ w1.float()) < TOL


@pytest.mark.parametrize("M", [1, 128, 300, 4113, 50000, 256 * 56 * 56])
def test_on_chip_weight_gradients_match_fp32_torch(M):
    """lnx_mlp_fused_wgrad recomputes h / dPre per row tile and accumulates dW1, db1, dW2 in tensor memory (nothing 4C wide
    in HBM); lnx_mlp_fused_bwd without the hidden stores is its dX companion.  Both accumulate / compare in relative L2:
    M-long fp32 sums of bf16-rounded products."""
    import linnaeus_b200.functional as F

    C, H = 96, 384
    x, w1, b1, w2, b2, gamma, res, dy = _make(M, C, seed=4)
    w2e = (w2.float() * gamma[:, None]).to(torch.bfloat16)
    dw1 = torch.zeros(H, C, device=DEV)
    db1 = torch.zeros(H, device=DEV)
    dw2 = torch.zeros(C, H, device=DEV)
    db2 = torch.zeros(C, device=DEV)
    F.mlp_fused_wgrad(x, dy, w1, b1, w2e, dw1, db1, dw2, db2)
    h0, dpre0, dx = F.mlp_fused_bwd(x, dy, w1, b1, w2e, store_hidden=False)
    assert h0 is None and dpre0 is None
    pre = x.float() @ w1.float().t() + b1
    h = TF.gelu(pre).to(torch.bfloat16).float()
    dpre = ((dy.float() @ w2e.float()) * _gelu_grad(pre)).to(torch.bfloat16).float()
    del pre
    for got, ref in ((dw1, dpre.t() @ x.float()), (db1, dpre.sum(0)), (dw2, dy.float().t() @ h), (db2, dy.float().sum(0)),
                     (dx.float(), dpre @ w1.float())):
        assert float((got - ref).norm() / ref.norm()) < 5e-3
    # accumulation: a second call adds the same amounts again
    F.mlp_fused_wgrad(x, dy, w1, b1, w2e, dw1, db1, dw2, None)
    assert float((dw1 - 2 * (dpre.t() @ x.float())).norm() / dw1.norm()) < 5e-3


def test_fused_kernels_at_the_bench_shape():
    """M = 802 816 rows: 6272 row tiles over 148 persistent CTAs (42 or 43 tiles each), the shape the train step launches."""
    import linnaeus_b200.functional as F

    M, C = 256 * 56 * 56, 96
    x, w1, b1, w2, b2, gamma, res, dy = _make(M, C, seed=2)
    y = F.mlp_fused_fwd(x, w1, b1, w2, b2, gamma=gamma, residual=res)
    pre = x.float() @ w1.float().t() + b1
    hq = TF.gelu(pre).to(torch.bfloat16).float()
    assert _rel(y, (hq @ w2.float().t() + b2) * gamma + res.float()) < TOL
    del hq, y
    w2e = (w2.float() * gamma[:, None]).to(torch.bfloat16)
    h, dpre, dx = F.mlp_fused_bwd(x, dy, w1, b1, w2e)
    assert _rel(h, TF.gelu(pre)) < TOL
    dpre_ref = (dy.float() @ w2e.float()) * _gelu_grad(pre)
    del pre
    assert _rel(dpre, dpre_ref) < TOL
    assert _rel(dx, dpre_ref.to(torch.bfloat16).float() @ w1.float()) < TOL


@pytest.mark.parametrize("drop", [False, True])
def test_mlp2_autograd_fused_equals_two_gemm_path(drop):
    """Same inputs through _Mlp2 with the fused kernels and with the two-GEMM path: outputs and every gradient agree to bf16 noise."""
    import linnaeus_b200.functional as F

    M, C, rpg = 6 * 196, 96, 196
    x, w1, b1, w2, b2, gamma, res, dy = _make(M, C, seed=5)
    rs = (torch.floor(0.7 + torch.rand(M // rpg, device=DEV)) / 0.7) if drop else None
    outs = []
    for fused in (True, False):
        F.FUSED_MLP = fused
        try:
            ps = [torch.nn.Parameter(t.float().clone()) for t in (w1, b1, w2, b2, gamma)]
            xi = x.clone().requires_grad_(True)
            ri = res.clone().requires_grad_(True)
            y = F.mlp2(xi, ps[0], ps[1], ps[2], ps[3], act="gelu", residual=ri, col_scale=ps[4], row_scale=rs, rows_per_group=rpg)
            y.backward(dy)
            outs.append([y.detach()] + [xi.grad, ri.grad] + [p.grad for p in ps])
        finally:
            F.FUSED_MLP = True
    names = ["y", "dx", "dres", "dw1", "db1", "dw2", "db2", "dgamma"]
    for n, a, b in zip(names, outs[0], outs[1]):
        assert _rel(a, b.float()) < 1.5e-2, n
