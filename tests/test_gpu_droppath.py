"""GPU: DropPath (stochastic depth) = per-sample mask applied in the GEMM epilogue that closes a residual branch,
against the reference formula x + branch * floor(keep + U) / keep (drop_path.py:11-36), forward and backward."""
import math

import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mlp2_and_linear_row_scale(dtype):
    import linnaeus_b200.functional as F

    torch.manual_seed(0)
    Bn, T, K, Hd = 6, 50, 64, 256
    M = Bn * T
    x = torch.randn(M, K, device=DEV).to(dtype).requires_grad_(True)
    res = torch.randn(M, K, device=DEV).to(dtype).requires_grad_(True)
    w1 = (torch.randn(Hd, K, device=DEV) / math.sqrt(K)).requires_grad_(True)
    b1 = (0.1 * torch.randn(Hd, device=DEV)).requires_grad_(True)
    w2 = (torch.randn(K, Hd, device=DEV) / math.sqrt(Hd)).requires_grad_(True)
    b2 = (0.1 * torch.randn(K, device=DEV)).requires_grad_(True)
    gamma = (torch.rand(K, device=DEV) + 0.5).requires_grad_(True)
    mask = torch.tensor([0.0, 1.25, 1.25, 0.0, 1.25, 1.25], device=DEV)
    y = F.mlp2(x, w1, b1, w2, b2, act="gelu", residual=res, col_scale=gamma, row_scale=mask, rows_per_group=T)
    g = torch.randn_like(y)
    y.backward(g)
    P = [t.detach().float().clone().requires_grad_(True) for t in (x, res, w1, b1, w2, b2, gamma)]
    xr, rr, w1r, b1r, w2r, b2r, gr = P
    br = TF.linear(TF.gelu(TF.linear(xr, w1r, b1r)), w2r, b2r) * gr
    yr = rr + br * mask.repeat_interleave(T)[:, None]
    yr.backward(g.float())
    t = 2e-2 if dtype == torch.bfloat16 else 1e-4
    assert rel_err(y, yr) < t
    for a, b_ in zip((x, res, w1, b1, w2, b2, gamma), P):
        assert rel_err(a.grad, b_.grad) < t
    # dropped samples: output == residual exactly, zero gradient into the branch input
    yv = y.detach().view(Bn, T, K)
    assert torch.equal(yv[0], res.detach().view(Bn, T, K)[0])
    assert float(x.grad.view(Bn, T, K)[3].abs().max()) == 0.0

    x2 = torch.randn(M, K, device=DEV).to(dtype).requires_grad_(True)
    w = (torch.randn(K, K, device=DEV) / math.sqrt(K)).requires_grad_(True)
    y2 = F.linear(x2, w, b2, residual=res.detach(), row_scale=mask, rows_per_group=T)
    y2.backward(g)
    xq, wq = x2.detach().float().requires_grad_(True), w.detach().clone().requires_grad_(True)
    (res.detach().float() + TF.linear(xq, wq, b2.detach()) * mask.repeat_interleave(T)[:, None]).backward(g.float())
    assert rel_err(x2.grad, xq.grad) < t and rel_err(w.grad, wq.grad) < t


def test_model_drop_path_train_vs_eval():
    import linnaeus_b200 as L
    from oracle import mformer_oracle as O

    cfg, nc = L.make_synthetic_config("sm", 64, dims=(32, 64, 128, 256), heads=(2, 4), rope_depths=(2, 1), conv_depths=(1, 1, 1, 1), drop_path=0.5)
    a = O.arch_from_config(cfg, nc)
    P = O.synth_state_dict(O.param_shapes(a), 0)
    x, meta, _ = O.synth_batch(a, 8, 0)
    m = L.build_model(cfg, nc)
    m.load_state_dict(P)
    m = m.to(DEV)
    assert m.stages[3][-1].drop_prob == pytest.approx(0.5)
    m.eval()
    with torch.no_grad():
        e1 = m(x.to(DEV), meta.to(DEV))["taxa_L10"]
        ref = O.forward(P, a, x, meta)["taxa_L10"]
    assert rel_err(e1.cpu(), ref) < 1e-4  # eval: DropPath is the identity
    m.train()
    torch.manual_seed(1)
    t1 = m(x.to(DEV), meta.to(DEV))["taxa_L10"]
    torch.manual_seed(2)
    t2 = m(x.to(DEV), meta.to(DEV))["taxa_L10"]
    assert rel_err(t1, e1) > 1e-3 and rel_err(t1, t2) > 1e-3  # stochastic in training
    t1.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
